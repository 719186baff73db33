#!/bin/bash
# quick GPU check: tests + bench (no ncu).  Usage: scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider ${2:+-k "$2"} 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench.json"))
r=l["roofline"]
print("value",round(l["value"],2),"e2e",round(l["e2e"]["value"],2),"ms/step",round(l["ms_per_step"],3),"launches/step",l["gpu_launches"]/l["steps"])
print("gather GB/s",round(r["achieved"],1),"frac",round(r["frac"],3),"gather ms/step",round(r["kpconv_gather_ms_per_step"],3),"step stats",l.get("step_ms_stats"))
print({k:round(v) for k,v in r["per_layer_GBps"].items()})
print(l["pose_vs_synthetic_gt"], l["clocks"])
PY
tail -3 gpurun_out/${TAG}_bench.err
