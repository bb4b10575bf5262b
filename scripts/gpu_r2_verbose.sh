set -x
export KHB_BENCH_E2E=0 KHB_BINS_VERBOSE=1
KHB_BENCH_GROUPS=2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "^\[bins\]" | tail -4
KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "^\[bins\]" | tail -4
