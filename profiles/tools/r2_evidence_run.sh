set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; tail -3 gpurun_out/r2b_gpu_tests.log
python bench.py --steps 200 --warmup 10 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; head -c 600 gpurun_out/r2b_bench_n1.json
bash profiles/tools/family_sweep.sh > gpurun_out/r2b_family_sweep.txt 2>&1; cat gpurun_out/r2b_family_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter_planes -s 3 -c 2 -o gpurun_out/r2b_prof_filter -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_select_nms -s 3 -c 2 -o gpurun_out/r2b_prof_nms -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_ncu_n.log 2>&1
python profiles/tools/decode_bench.py > gpurun_out/r2b_decode_bench.txt 2>&1; cat gpurun_out/r2b_decode_bench.txt
