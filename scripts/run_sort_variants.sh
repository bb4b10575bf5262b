for v in ${VARIANTS:-5 20}; do KHB_SORT_VARIANT=$v python scripts/bench_sort.py 100000000 31 1 2>&1 | tail -3; done
