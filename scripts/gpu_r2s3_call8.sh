# Round 2, session 3, call 8 (two GPUs): the across-group stage sorts the pushed regions where they lie (khb_peer_across, gathered first radix pass):
# peer + team tests, then config 2 per GPU on 2 GPUs with and without the import copy
set -x
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_team.py -x -q > gpurun_out/s3c8_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s3c8_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" > gpurun_out/s3c8_$name.json 2> gpurun_out/s3c8_$name.err; echo "$name rc=$?"; }
run c2x2_gather KHB_BENCH_E2E=0 timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3
run c2x2_import KHB_BENCH_E2E=0 KHB_PEER_GATHER=0 timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3
python - <<'PY'
import json
for f in ("c2x2_gather", "c2x2_import"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c8_{f}.json") if l.startswith("{")][-1])
        k = d["kernels"]
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), d["parity_in_run"], d["config"]["exchange"][:40],
              {n: (v["launches"], round(v["ms"] / v["launches"], 3)) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -h FATAL gpurun_out/s3c8_*.err | head -3
