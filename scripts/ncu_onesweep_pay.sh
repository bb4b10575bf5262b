export KHB_BENCH_GROUPS=2 KHB_BENCH_GENOMES=25
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 1 -c 2 -f -o gpurun_out/prof_onesweep_pay $CMD > gpurun_out/ncu3.log 2>&1
echo "full rc=$?"
