# Round 2, call 5: bins tests (incl. the partition against its numpy statement), default bench
set -x
timeout 1200 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c5_pytest.log
python bench.py > gpurun_out/r2c5_bench_c2.json 2> gpurun_out/r2c5_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c5_bench_c2.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_in_run"], d["roofline"]["kernel"][:40], d["roofline"]["frac"], json.dumps(d["kernels"]))
PY
