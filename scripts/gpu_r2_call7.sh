set -x
timeout 1200 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c7_pytest.log
export KHB_BENCH_E2E=0
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2c7_$name.json 2> gpurun_out/r2c7_$name.err; echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2c7_{f}.json") if l.startswith("{")][-1])
    print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["parity_in_run"], d["config"].get("bins_counters"), {k: (v["launches"], round(v["ms"] / v["launches"], 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "unreadable", e)
PY
}
run c2 KHB_BENCH_GROUPS=4
run k21 KHB_BENCH_GROUPS=3 KHB_BENCH_K=21
run k47 KHB_BENCH_GROUPS=3 KHB_BENCH_K=47
