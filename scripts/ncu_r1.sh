# Round-1 evidence: launch list of bench.py + full capture of the dominant kernel (same command, run plain first).
export KHB_BENCH_GROUPS=2 KHB_BENCH_GENOMES=25
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 9 -c 2 -f -o gpurun_out/prof_onesweep_final $CMD > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
