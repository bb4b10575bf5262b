set -x
timeout 1200 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_pytest.log
export KHB_BENCH_E2E=0
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/quick_$name.json 2> gpurun_out/quick_$name.err
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/quick_{f}.json") if l.startswith("{")][-1])
    k = d["kernels"]
    print(f, round(d["value"], 2), "part", round(k["bin_partition"]["ms"] / k["bin_partition"]["launches"], 3), "count", round(k["bin_count"]["ms"] / k["bin_count"]["launches"], 3), d["config"].get("bins_counters"), d["parity_in_run"])
except Exception as e:
    print(f, "unreadable", e)
PY
}
run c2 KHB_BENCH_GROUPS=3
run k47 KHB_BENCH_GROUPS=3 KHB_BENCH_K=47
run c5 KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=2
