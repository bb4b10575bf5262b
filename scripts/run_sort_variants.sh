for v in ${VARIANTS:-1 5 6 7 2 3}; do KHB_SORT_VARIANT=$v python scripts/bench_sort.py 100000000 31 1 2>&1 | tail -3; done
