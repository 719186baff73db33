#!/bin/bash
TAG=${1:-s4d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
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
head -12 gpurun_out/${TAG}_timeline.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tf_attend|sinkhorn128" -c 3 -o gpurun_out/${TAG}_attn \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_attn.log 2>&1
ls -la gpurun_out | tail -4
