# ncu --set full of the two hash group-stage kernels on one config-2 group (50 genomes x 5 Mbp, k=31).
# usage (GPU box): bash scripts/ncu_hash.sh <tag>
export KHB_BENCH_GROUPS=1 KHB_BENCH_GENOMES=${GENOMES:-50}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_hash_$1.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_hash_$1.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:hash_insert -s 1 -c 1 -o gpurun_out/prof_hash_insert_$1 -f $CMD > gpurun_out/ncu_hi.log 2>&1
echo "insert rc=$?"; tail -2 gpurun_out/ncu_hi.log
ncu --set full --clock-control none --import-source on -k regex:hash_count -s 1 -c 1 -o gpurun_out/prof_hash_count_$1 -f $CMD > gpurun_out/ncu_hc.log 2>&1
echo "count rc=$?"; tail -2 gpurun_out/ncu_hc.log
