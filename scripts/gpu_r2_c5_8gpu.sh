# Round 2: the north-star target.  Config 5 (100 groups x 200 genomes, ~100 Gbases, k = 31) on the 8 B200s of one box: groups dealt over the ranks,
# every bin's distinct k-mers stored straight into their hash-range owner's peer region, CSVs written by rank 0, oracle on two sampled groups.
# Then config 2 on 8 GPUs (weak scaling: 10 groups per GPU), with and without the fused push.
set -x
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c; nproc; free -g | sed -n 2p
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
export KHB_BENCH_CONFIG=5 KHB_BENCH_CSV_DIR=gpurun_out/r2_c5_csv
mkdir -p $KHB_BENCH_CSV_DIR
timeout 1500 $TR --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_c5_8gpu.json 2> gpurun_out/r2_c5_8gpu.err; echo "c5 rc=$?"
tail -c 600 gpurun_out/r2_c5_8gpu.json; ls -la $KHB_BENCH_CSV_DIR
unset KHB_BENCH_CSV_DIR KHB_BENCH_CONFIG
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_c2_8gpu.json 2> gpurun_out/r2_c2_8gpu.err; echo "c2x8 rc=$?"
KHB_PEER_FUSE=0 KHB_BENCH_E2E=0 timeout 600 $TR --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_c2_8gpu_nofuse.json 2> gpurun_out/r2_c2_8gpu_nofuse.err; echo "c2x8 nofuse rc=$?"
python - <<'PY'
import json
for f in ("r2_c5_8gpu", "r2_c2_8gpu", "r2_c2_8gpu_nofuse"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"], 1), d["parity_in_run"], d["config"]["exchange"][:60], {k: (v["launches"], round(v["ms"], 1)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
