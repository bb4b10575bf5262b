# usage: bash scripts/bench_kernels.sh tag [ENV=VAL ...]   -- short bench run, prints value + per-kernel ms
tag=$1; shift
env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bk_$tag.log 2>&1
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
line = open(f"gpurun_out/bk_{tag}.log").read().strip().splitlines()[-1]
try:
    j = json.loads(line)
    print(tag, "value", round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "ms/step", round(j["ms_per_step"], 1),
          {k: round(v["ms"] / j["steps"], 2) for k, v in j["kernels"].items() if v["launches"]})
except Exception as e:
    print(tag, "FAILED", e, line[-500:])
PY
