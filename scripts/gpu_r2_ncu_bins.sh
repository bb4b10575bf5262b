# ncu --set full of the kernels of the minimizer-bin group stage: one config-2 group, then one config-5-shaped group (200 genomes)
set -x
export KHB_BENCH_GROUPS=1 KHB_BENCH_E2E=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_bins_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'mb_partition|mb_count' -s 2 -c 2 -f -o gpurun_out/ncu_bins $CMD > gpurun_out/ncu_bins.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_bins.log
unset KHB_BENCH_GROUPS
export KHB_BENCH_CONFIG=5 KHB_BENCH_GROUPS_TOTAL=1
$CMD > gpurun_out/ncu_bins5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'mb_count' -s 1 -c 1 -f -o gpurun_out/ncu_bins5 $CMD > gpurun_out/ncu_bins5.log 2>&1
echo "ncu5 rc=$?"; tail -1 gpurun_out/ncu_bins5.log
