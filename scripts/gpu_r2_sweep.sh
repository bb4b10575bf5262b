# parameter sweep of the minimizer-bin group stage on 2 config-2 groups (and 2 config-5-shaped groups)
set -x
timeout 900 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/sweep_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/sweep_pytest.log
export KHB_BENCH_GROUPS=2 KHB_BENCH_E2E=0
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/sweep_{n}.json") if l.startswith("{")][-1])
    k = d["kernels"]
    print(n, "value", round(d["value"], 2), "part ms/launch", round(k["bin_partition"]["ms"] / k["bin_partition"]["launches"], 3),
          "count ms/launch", round(k["bin_count"]["ms"] / k["bin_count"]["launches"], 3), d["config"].get("bins_counters"), "parity", d["parity_in_run"])
except Exception as e:
    print(n, "unreadable", e)
PY
}
run default X=1
run d256 KHB_BINS_DCAP=256
run w3000 KHB_BINS_WPB=3000
export KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=2
unset KHB_BENCH_GROUPS
run c5_default X=1
run c5_w6000 KHB_BINS_WPB=6000
