# Round 2, session 3, call 1 (one GPU): the team tests (two / three processes on cuda:0), the peer tests (owner-sorted stores of the fused push),
# the bins tests (planner refactor), then a short config-2 bench to see that the single-GPU path did not move
set -x
timeout 700 python -m pytest tests/test_gpu_team.py -x -q > gpurun_out/s3c1_team.log 2>&1; echo "team rc=$?"; tail -30 gpurun_out/s3c1_team.log
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_bins.py -x -q > gpurun_out/s3c1_peer_bins.log 2>&1; echo "peer+bins rc=$?"; tail -5 gpurun_out/s3c1_peer_bins.log
KHB_BENCH_E2E=0 KHB_BENCH_GROUPS=4 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s3c1_c2.json 2> gpurun_out/s3c1_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/s3c1_c2.json") if l.startswith("{")][-1])
    k = d["kernels"]
    print("c2", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), {n: round(v["ms"] / v["launches"], 3) for n, v in k.items()}, d["parity_in_run"])
except Exception as e:
    print("unreadable", e)
PY
