(python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
run() { echo "== $*"; python bench.py "$@" --steps 200 --warmup 10 --no-extras --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],5), round(d['roofline']['frac'],4), round(d['roofline']['launch_ms_alone'],5), d.get('stages_ms'))"; }
run --config c2 --dist sparse; run --config c2 --family yolov7 --dist sparse; run --config c3 --dist sparse
export YSB_LIBRARY=$PWD/yoloseries_b200/_lib/libysb_postproc_k2t.so; python profiles/tools/k2_timing.py yolox 256 2>&1 | tail -4; python profiles/tools/k2_timing.py yolov5 64 2>&1 | tail -6
