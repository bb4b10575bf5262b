set -x
export KHB_BENCH_E2E=0 KHB_BENCH_GROUPS=3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sw2_$name.json 2> gpurun_out/sw2_$name.err
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/sw2_{f}.json") if l.startswith("{")][-1])
    k = d["kernels"]
    print(f, round(d["value"], 2), "part", round(k["bin_partition"]["ms"] / k["bin_partition"]["launches"], 3), "count", round(k["bin_count"]["ms"] / k["bin_count"]["launches"], 3), d["config"].get("bins_counters"), d["parity_in_run"])
except Exception as e:
    print(f, "unreadable", e)
PY
}
run default X=1
run w2000 KHB_BINS_WPB=2000
run w2500 KHB_BINS_WPB=2500
run w3500 KHB_BINS_WPB=3500
run w2000_d256 KHB_BINS_WPB=2000 KHB_BINS_DCAP=256
run w2500_s10 KHB_BINS_WPB=2500 KHB_BINS_SLOTS_LOG2=10
