#!/bin/bash
TAG=${1:-r2C}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_full.log
grep -E "^FAILED" gpurun_out/${TAG}_pytest_full.log | head
grep -E "^\[parity\]|^\[grad parity\]|^\[precision\]|^\[train\]|passed|failed" gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_tests_gpu.log
for conf in default expandable; do
  if [ $conf = expandable ]; then export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True; fi
  timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/${TAG}_bench_$conf.json 2> gpurun_out/${TAG}_bench_$conf.err
  python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench_$conf.json")); s=l["step_ms_stats"]
print("$conf value",round(l["value"],1),"median",round(s["median"],3),"max",round(s["max"],1),"outliers",s["outlier_steps"],"e2e",round(l["e2e"]["value"],1),"gather frac",round(l["roofline"]["frac"],3))
PY
  tail -2 gpurun_out/${TAG}_bench_$conf.err
done
