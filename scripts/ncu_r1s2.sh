# Session-2 evidence: launch list of bench.py + ncu --set full of every kernel of one group + the across stage
# (same command, run plain first).  Outputs under gpurun_out/.
export KHB_BENCH_GROUPS=2 KHB_BENCH_GENOMES=50
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_s2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s2.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
export KHB_BENCH_GROUPS=1
$CMD > gpurun_out/plain_s2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'fasta_|extract64|radix_hist|onesweep|pairs_kernel|mixed_|runs_kernel' -c 22 -f -o gpurun_out/prof_all_s2 $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_f.log
