# Session-3 evidence: launch list of bench.py (default single-sort path, after the exchanger / peer / hash / exp-6 additions).
export KHB_BENCH_GROUPS=2 KHB_BENCH_GENOMES=50
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_s3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s3.csv $CMD > gpurun_out/ncu_l3.log 2>&1
echo "launch list rc=$?"
