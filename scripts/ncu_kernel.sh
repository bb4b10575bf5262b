# usage: bash scripts/ncu_kernel.sh <kernel-regex> <skip> <count> <tag>
export KHB_BENCH_GROUPS=1 KHB_BENCH_GENOMES=${GENOMES:-20}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/prof_$4 -f $CMD > gpurun_out/ncu_k.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_k.log
