#!/bin/bash
TAG=${1:-s4g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
for w in 4 8 16 32; do RDM_GATHER_WPS=$w timeout 200 python scripts/bench_gather.py > gpurun_out/${TAG}_gather_wps$w.log 2>&1; echo "WPS=$w"; tail -15 gpurun_out/${TAG}_gather_wps$w.log | awk '{print $2,$3,$9,$10,$11}' | tr '\n' ';'; echo; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench.json"))
r=l["roofline"]
print("value",round(l["value"],2),"e2e",round(l["e2e"]["value"],2),"ms/step",round(l["ms_per_step"],3),"launches/step",l["gpu_launches"]/l["steps"])
print("gather GB/s",round(r["achieved"],1),"frac",round(r["frac"],3),"gather ms/step",round(r["kpconv_gather_ms_per_step"],3),"wgemm ms/step",round(r["kpconv_weight_gemm_ms_per_step"],3))
print(l["pose_check"], l["clocks"])
PY
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/timeline.py $TAG 4 > gpurun_out/${TAG}_timeline.log 2>&1 || tail -5 gpurun_out/${TAG}_timeline.log
head -14 gpurun_out/${TAG}_timeline.md
