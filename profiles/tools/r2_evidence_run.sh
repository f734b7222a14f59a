# Round-2 (final state) evidence on one B200 (run under gpurun): tests, the driver-shaped bench line, the reference arm, the
# other BASELINE configs, the family sweep, the ncu launch list and ncu --set full captures of the tuned kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
python bench.py --steps 200 --warmup 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; head -c 300 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; head -c 400 gpurun_out/r2_bench_reference_n1.json
for c in c3 c4 c5; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; head -c 200 gpurun_out/r2_bench_$c.json; echo; done
bash profiles/tools/family_sweep.sh > gpurun_out/r2_family_sweep.txt 2>&1; grep "^--family" gpurun_out/r2_family_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter_planes -s 3 -c 2 -o gpurun_out/prof_filter_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_select_nms -s 3 -c 2 -o gpurun_out/prof_nms_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_select_nms -s 3 -c 2 -o gpurun_out/prof_nms_c3_r2 -f python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_n3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter_rows -s 3 -c 2 -o gpurun_out/prof_rows_r2 -f python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_r.log 2>&1
ls -la gpurun_out | head -40
