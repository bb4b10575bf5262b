# Round 2, session 3, call 6 (two GPUs): team tests, then shapes with fewer groups than GPUs can share evenly:
#   config-4-shaped groups (100 genomes, k = 47), 3 groups: sharded in a team of 2 against whole groups dealt 2 / 1
#   ONE config-5-shaped group (200 genomes, k = 31): both GPUs on the one group against one GPU
set -x
timeout 400 python -m pytest tests/test_gpu_team.py -x -q > gpurun_out/s3c6_team.log 2>&1; echo "team rc=$?"; tail -4 gpurun_out/s3c6_team.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" > gpurun_out/s3c6_$name.json 2> gpurun_out/s3c6_$name.err; echo "$name rc=$?"; }
run c4x3_team2 KHB_BENCH_CONFIG=4 KHB_BENCH_GROUPS_TOTAL=3 timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2
run c4x3_whole KHB_BENCH_CONFIG=4 KHB_BENCH_GROUPS_TOTAL=3 KHB_BENCH_TEAM=1 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 2
run c5x1_team2 KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=1 timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 2
run c5x1_one KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=1 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline
python - <<'PY'
import json
for f in ("c4x3_team2", "c4x3_whole", "c5x1_team2", "c5x1_one"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c6_{f}.json") if l.startswith("{")][-1])
        k = d["kernels"]
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, d["parity_in_run"],
              d["config"]["parallelism"], {n: (v["launches"], round(v["ms"] / v["launches"], 3), v["alg_GBps"]) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
