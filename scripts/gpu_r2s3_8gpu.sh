# Round 2, session 3: config 2 per GPU on 8 B200s (weak scaling) with the final code: routed end-of-bin pass, pushed regions sorted where they lie
set -x
KHB_BENCH_E2E=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s3_c2_8gpu.json 2> gpurun_out/s3_c2_8gpu.err; echo "c2x8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/s3_c2_8gpu.json") if l.startswith("{")][-1])
    k = d["kernels"]
    print("c2x8", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), d["parity_in_run"], d["config"]["exchange"][:60], {n: (v["launches"], round(v["ms"] / v["launches"], 3)) for n, v in k.items()})
except Exception as e:
    print("unreadable", e)
PY
grep -h "FATAL\|Error" gpurun_out/s3_c2_8gpu.err | head -5
