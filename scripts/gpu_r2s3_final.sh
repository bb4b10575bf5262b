# Round 2, session 3, final single-GPU validation: the whole GPU suite, the default bench with the CPU arm, smoke
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s3f_pytest.log
python bench.py > gpurun_out/s3f_bench_c2.json 2> gpurun_out/s3f_bench_c2.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s3f_smoke.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/s3f_bench_c2.json") if l.startswith("{")][-1])
print("c2", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 2), d["parity_in_run"], {n: round(v["ms"] / v["launches"], 3) for n, v in d["kernels"].items()})
PY
