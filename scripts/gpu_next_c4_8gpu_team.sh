# NOT RUN YET (no GPU budget left in round 2): config 4 (20 groups x 100 genomes, k = 47) on 8 B200s, teams of 2 (5 groups per team, every group
# sharded) against whole groups dealt 3/3/3/3/2/2/2/2, and config 5 (100 groups: 13 / 12 whole, 25 per team of 2).  Launch with
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_next_c4_8gpu_team.sh'
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" > gpurun_out/next_$name.json 2> gpurun_out/next_$name.err; echo "$name rc=$?"; }
run c4x8_team2 KHB_BENCH_CONFIG=4 KHB_BENCH_E2E=0 timeout 600 $TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 2
run c4x8_whole KHB_BENCH_CONFIG=4 KHB_BENCH_E2E=0 KHB_BENCH_TEAM=1 timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 2
run c5x8_team2 KHB_BENCH_CONFIG=5 KHB_BENCH_E2E=0 KHB_BENCH_TEAM=2 timeout 900 $TR --master-port 29553 bench.py --gpus 8 --steps 2 --warmup 1
python - <<'PY'
import json
for f in ("c4x8_team2", "c4x8_whole", "c5x8_team2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/next_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), d["parity_in_run"], d["config"]["parallelism"],
              {n: (v["launches"], round(v["ms"] / v["launches"], 3)) for n, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
