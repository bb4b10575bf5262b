# Round 2, call 1 (one B200): the whole GPU suite incl. the full-size oracle comparisons, the default bench (config 2) with
# in-run parity, driver-visible config-3 / config-4 records, and a fresh ncu capture of the dominant kernel's DRAM traffic.
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | sed -n 2p
python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c1_pytest.log
python bench.py > gpurun_out/r2c1_bench_c2.json 2> gpurun_out/r2c1_bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c1_ref_c2.json 2> gpurun_out/r2c1_ref_c2.err; echo "ref rc=$?"
KHB_BENCH_CONFIG=3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c1_bench_c3.json 2> gpurun_out/r2c1_bench_c3.err; echo "c3 rc=$?"
KHB_BENCH_CONFIG=4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c1_bench_c4_k47.json 2> gpurun_out/r2c1_bench_c4_k47.err; echo "c4/47 rc=$?"
KHB_BENCH_CONFIG=4 KHB_BENCH_K=63 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c1_bench_c4_k63.json 2> gpurun_out/r2c1_bench_c4_k63.err; echo "c4/63 rc=$?"
export KHB_BENCH_GROUPS=1 KHB_BENCH_E2E=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r2c1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 2 -f -o gpurun_out/r2c1_onesweep $CMD > gpurun_out/r2c1_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2c1_ncu.log
KHB_BENCH_K=47 $CMD > gpurun_out/r2c1_plain128.log 2>&1 && KHB_BENCH_K=47 ncu --set full --clock-control none --import-source on -k regex:'onesweep|pairs_kernel' -s 5 -c 3 -f -o gpurun_out/r2c1_key128 $CMD > gpurun_out/r2c1_ncu128.log 2>&1
echo "ncu128 rc=$?"
