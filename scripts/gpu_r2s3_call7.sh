# Round 2, session 3, call 7 (two GPUs): team tests incl. the work-root driver with a team, then ONE config-5-shaped group (200 genomes, k = 31) on both GPUs
set -x
timeout 500 python -m pytest tests/test_gpu_team.py -x -q > gpurun_out/s3c7_team.log 2>&1; echo "team rc=$?"; tail -4 gpurun_out/s3c7_team.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" > gpurun_out/s3c7_$name.json 2> gpurun_out/s3c7_$name.err; echo "$name rc=$?"; }
run c5x1_team2 KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=1 timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 2
run c5x3_team2 KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=3 timeout 600 $TR --master-port 29524 bench.py --gpus 2 --steps 3 --warmup 2
python - <<'PY'
import json
for f in ("c5x1_team2", "c5x3_team2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c7_{f}.json") if l.startswith("{")][-1])
        k = d["kernels"]
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, d["parity_in_run"],
              d["config"]["parallelism"], {n: (v["launches"], round(v["ms"] / v["launches"], 3), v["alg_GBps"]) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -h FATAL gpurun_out/s3c7_*.err | head -3
