# usage: bash profiles/tools/r2_multi_run.sh N   (under gpurun --gpus N)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
tail -2 gpurun_out/r2b_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "gather", d.get("gather"), "filter_alone", d.get("filter_alone_ms_per_rank"), "clocks", d["clocks"])
print("c2_strong", d.get("c2_strong"))
PY
