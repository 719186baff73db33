#!/bin/bash
TAG=${1:-s4i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
timeout 200 python scripts/bench_gather.py > gpurun_out/${TAG}_gather.log 2>&1; tail -15 gpurun_out/${TAG}_gather.log | awk '{print $2,$3,$9,$10,$11}' | tr '\n' ';'; echo
for mode in "" "--no-pipeline"; do
timeout 600 python bench.py --no-cpu-baseline $mode > gpurun_out/${TAG}_bench$mode.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench$mode.json"))
r=l["roofline"]
print("$mode value",round(l["value"],2),"e2e",round(l["e2e"]["value"],2),"ms/step",round(l["ms_per_step"],3),"launches/step",l["gpu_launches"]/l["steps"])
print("gather GB/s",round(r["achieved"],1),"frac",round(r["frac"],3),"gather ms/step",round(r["kpconv_gather_ms_per_step"],3),"wgemm ms/step",round(r["kpconv_weight_gemm_ms_per_step"],3))
print(l["pose_check"], l["clocks"])
PY
tail -3 gpurun_out/${TAG}_bench.err
done
timeout 300 python scripts/timeline.py $TAG 4 > gpurun_out/${TAG}_timeline.log 2>&1 || tail -5 gpurun_out/${TAG}_timeline.log
head -14 gpurun_out/${TAG}_timeline.md
