set -x
export KHB_BENCH_GROUPS=10 KHB_BENCH_E2E=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'mb_across' -s 1 -c 1 -f -o gpurun_out/ncu_across $CMD > gpurun_out/ncu_across.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_across.log | cut -c1-200
