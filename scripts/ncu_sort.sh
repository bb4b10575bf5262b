export KHB_SORT_VARIANT=${1:-1}
CMD="python scripts/bench_sort.py 60000000 31 1"
$CMD > gpurun_out/plain_sort.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 10 -c 2 -o gpurun_out/prof_sort_v${KHB_SORT_VARIANT} -f $CMD > gpurun_out/ncu_sort.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_sort.log
