# Round 2, session 3, call 5 (two GPUs): team tests after the second restructuring (local partition, regions packed densely into the owners buffers), then config-4-shaped groups sharded in a team of 2
# against whole groups dealt to the 2 ranks
set -x
timeout 400 python -m pytest tests/test_gpu_team.py -x -q > gpurun_out/s3c5_team.log 2>&1; echo "team rc=$?"; tail -15 gpurun_out/s3c5_team.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export KHB_BENCH_CONFIG=4 KHB_BENCH_GROUPS_TOTAL=4
KHB_BENCH_TEAM=2 timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/s3c5_team2.json 2> gpurun_out/s3c5_team2.err; echo "team2 rc=$?"
python - <<'PY'
import json
for f in ("team2",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c5_{f}.json") if l.startswith("{")][-1])
        k = d["kernels"]
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, d["parity_in_run"],
              d["config"]["parallelism"], {n: (v["launches"], round(v["ms"] / v["launches"], 3), v["alg_GBps"]) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/s3c5_team2.err
