# Round 2, call 4: 128-bit keys through the minimizer bins + planner: parity, then config-2/4/5-shaped timing
set -x
timeout 1200 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c4_pytest.log
export KHB_BENCH_E2E=0 KHB_BINS_VERBOSE=1
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c4_$name.json 2> gpurun_out/r2c4_$name.err; echo "$name rc=$?"
  grep "^\[bins\]" gpurun_out/r2c4_$name.err | tail -2
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2c4_{f}.json") if l.startswith("{")][-1])
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["parity_in_run"], d["config"].get("bins_counters"), {k: (v["launches"], v["ms"]) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "unreadable", e)
PY
}
run c2 KHB_BENCH_GROUPS=3
run c4_k47 KHB_BENCH_CONFIG=4 KHB_BENCH_GROUPS_TOTAL=4 KHB_BENCH_K=47
run c4_k63 KHB_BENCH_CONFIG=4 KHB_BENCH_GROUPS_TOTAL=4 KHB_BENCH_K=63
