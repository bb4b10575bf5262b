# Round 2, call 2: first run of the minimizer-bin group stage (bins.cu): parity tests, then config-2 timing in both modes.
set -x
timeout 900 python -m pytest tests/test_gpu_bins.py -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2c2_pytest.log
export KHB_BENCH_GROUPS=3 KHB_BENCH_E2E=0
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c2_bins.json 2> gpurun_out/r2c2_bins.err; echo "bins rc=$?"
KHB_GROUP_MODE=single-sort timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c2_sort.json 2> gpurun_out/r2c2_sort.err; echo "sort rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c2_bins.json", "gpurun_out/r2c2_sort.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], json.dumps(d["kernels"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c2_bins.err
# config-5-shaped groups (200 genomes, 4 chunks of genome bits) and config-4-shaped ones at k=31 for the table variants
KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=2 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c2_c5shape.json 2> gpurun_out/r2c2_c5shape.err; echo "c5 rc=$?"
KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=2 KHB_GROUP_MODE=single-sort timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c2_c5shape_sort.json 2> gpurun_out/r2c2_c5shape_sort.err; echo "c5 sort rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c2_c5shape.json", "gpurun_out/r2c2_c5shape_sort.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], json.dumps(d["kernels"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
