# Round 2, session 3, call 12 (two GPUs): config 2 per GPU on 2 GPUs, the end-of-bin routing with one reservation round trip per bin: owner-sorted stores off (default) and on
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" > gpurun_out/s3c12_$name.json 2> gpurun_out/s3c12_$name.err; echo "$name rc=$?"; }
run c2x2_plain KHB_BENCH_E2E=0 timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3
run c2x2_sorted KHB_BENCH_E2E=0 KHB_PEER_SORTED=1 timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3
python - <<'PY'
import json
for f in ("c2x2_plain", "c2x2_sorted"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c12_{f}.json") if l.startswith("{")][-1])
        k = d["kernels"]
        print(f, round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), d["parity_in_run"], d["config"]["exchange"][:40],
              {n: (v["launches"], round(v["ms"] / v["launches"], 3)) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -h FATAL gpurun_out/s3c12_*.err | head -3
