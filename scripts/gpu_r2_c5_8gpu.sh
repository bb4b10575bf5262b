# Round 2: the north-star target.  Config 5 (100 groups x 200 genomes, ~100 Gbases, k = 31) on the 8 B200s of one box: groups dealt over the ranks,
# hash-range exchange of the groups' distinct k-mers over peer memory for the across-group stage, CSVs written by rank 0, oracle on two sampled groups.
set -x
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c; nproc; free -g | sed -n 2p
export KHB_BENCH_CONFIG=5 KHB_BENCH_CSV_DIR=gpurun_out/r2_c5_csv
mkdir -p $KHB_BENCH_CSV_DIR
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 \
  > gpurun_out/r2_c5_8gpu.json 2> gpurun_out/r2_c5_8gpu.err; echo "c5 rc=$?"
tail -c 3000 gpurun_out/r2_c5_8gpu.json; tail -5 gpurun_out/r2_c5_8gpu.err; ls -la $KHB_BENCH_CSV_DIR
unset KHB_BENCH_CSV_DIR
# config 4 (20 groups x 100 genomes, k = 47) on 4 GPUs: strong scaling row
KHB_BENCH_CONFIG=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 1 \
  > gpurun_out/r2_c4_4gpu.json 2> gpurun_out/r2_c4_4gpu.err; echo "c4 rc=$?"
tail -c 1500 gpurun_out/r2_c4_4gpu.json
