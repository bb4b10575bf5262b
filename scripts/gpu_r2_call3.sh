# Round 2, call 3: the whole GPU suite with the minimizer-bin group stage as the default, then the default bench (config 2) and config 3.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c3_pytest.log
python bench.py > gpurun_out/r2c3_bench_c2.json 2> gpurun_out/r2c3_bench_c2.err; echo "bench rc=$?"
KHB_BENCH_CONFIG=3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c3_bench_c3.json 2> gpurun_out/r2c3_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c3_bench_c2.json", "gpurun_out/r2c3_bench_c3.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d["parity_in_run"], json.dumps(d["kernels"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
