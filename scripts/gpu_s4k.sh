#!/bin/bash
TAG=${1:-s4k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -12
for pdl in 1 0; do
RDM_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_pdl$pdl.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench_pdl$pdl.json"))
r=l["roofline"]
print("PDL=$pdl value",round(l["value"],2),"e2e",round(l["e2e"]["value"],2),"ms/step",round(l["ms_per_step"],3),"launches/step",l["gpu_launches"]/l["steps"])
print("gather GB/s",round(r["achieved"],1),"frac",round(r["frac"],3),"gather ms/step",round(r["kpconv_gather_ms_per_step"],3),"wgemm ms/step",round(r["kpconv_weight_gemm_ms_per_step"],3), l["pose_check"])
PY
tail -3 gpurun_out/${TAG}_bench.err
done
timeout 300 python scripts/timeline.py $TAG 4 > gpurun_out/${TAG}_timeline.log 2>&1 || tail -5 gpurun_out/${TAG}_timeline.log
head -8 gpurun_out/${TAG}_timeline.md; grep -n "gap histogram" -A9 gpurun_out/${TAG}_timeline.md
