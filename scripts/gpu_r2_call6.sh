# Round 2, call 6: across-group stage bin by bin: tests, then timing on config 2 (and the sort for comparison)
set -x
timeout 1200 python -m pytest tests/test_gpu_bins.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c6_pytest.log
export KHB_BENCH_E2E=0 KHB_BENCH_GROUPS=10
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2c6_$name.json 2> gpurun_out/r2c6_$name.err; echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2c6_{f}.json") if l.startswith("{")][-1])
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["parity_in_run"], d["config"].get("bins_counters"), {k: (v["launches"], v["ms"]) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "unreadable", e)
PY
}
run bins X=1
run sort KHB_ACROSS_MODE=sort
run s13 KHB_ACROSS_SLOTS_LOG2=13 KHB_BINS_VERBOSE=1
run s11 KHB_ACROSS_SLOTS_LOG2=11
