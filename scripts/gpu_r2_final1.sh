# Round 2, final single-GPU validation: the whole GPU suite, the default bench, configs 3 and 4, the CPU arm, ncu launch list + full captures
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref_c2.json 2> gpurun_out/r2f_ref_c2.err; echo "ref rc=$?"
KHB_BENCH_CONFIG=3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err; echo "c3 rc=$?"
KHB_BENCH_CONFIG=4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_c4_k47.json 2> gpurun_out/r2f_bench_c4_k47.err; echo "c4/47 rc=$?"
KHB_BENCH_CONFIG=4 KHB_BENCH_K=63 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_c4_k63.json 2> gpurun_out/r2f_bench_c4_k63.err; echo "c4/63 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
export KHB_BENCH_GROUPS=2 KHB_BENCH_E2E=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_launches.log 2>&1; echo "launch list rc=$?"
export KHB_BENCH_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:'mb_partition|mb_count' -s 2 -c 2 -f -o gpurun_out/r2f_bins $CMD > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
KHB_BENCH_K=47 ncu --set full --clock-control none --import-source on -k regex:'mb_count' -s 1 -c 1 -f -o gpurun_out/r2f_bins128 $CMD > gpurun_out/r2f_ncu128.log 2>&1; echo "ncu128 rc=$?"
