#!/bin/bash
# Throughput of the other BASELINE configs (parity-test cases, not the headline bench line).
run() { python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(' '.join(sys.argv[1:]), '| img/s', round(d['value']), '| ms/step', round(d['ms_per_step'],4), '| filter', round(d.get('stages_ms',{}).get('filter_compact',0),4), 'nms', round(d.get('stages_ms',{}).get('select_sort_nms',0),4), '| graph', d.get('cuda_graph'), '| filter alone', round(d['roofline']['launch_ms_alone'],4), 'frac_alone', round(d['roofline']['frac_alone'],3), '| M', round(d['survivors_per_image']))" "$@"; }
run --family yolov5 --batch 64
run --family yolov7 --batch 64
run --family yolox --batch 256
run --family yolov8 --batch 64
run --family retinanet --batch 64
run --family fcos --batch 256
run --family yolov5 --img 1280 --batch 16 --dist dense
run --family yolov5 --img 1280 --batch 16 --dist crowd
run --family yolov7 --batch 64 --dist sparse
run --family retinanet --batch 64 --dist sparse
