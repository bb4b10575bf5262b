set -x
export KHB_BENCH_GROUPS=1 KHB_BENCH_E2E=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'mb_partition|mb_count' -s 2 -c 2 -f -o gpurun_out/ncu_bins $CMD > gpurun_out/ncu_bins.log 2>&1
echo "ncu rc=$?"
