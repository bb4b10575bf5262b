export KHB_BENCH_GROUPS=2 KHB_BENCH_GENOMES=10
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 20 -c 3 -o gpurun_out/prof_onesweep $CMD > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu2.log
