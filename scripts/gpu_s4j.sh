#!/bin/bash
TAG=${1:-s4j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
timeout 300 python scripts/bench_gemm.py > gpurun_out/${TAG}_gemm.log 2>&1; cat gpurun_out/${TAG}_gemm.log | tail -22
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench.json"))
r=l["roofline"]
print("value",round(l["value"],2),"e2e",round(l["e2e"]["value"],2),"ms/step",round(l["ms_per_step"],3),"launches/step",l["gpu_launches"]/l["steps"])
print("gather GB/s",round(r["achieved"],1),"frac",round(r["frac"],3),"gather ms/step",round(r["kpconv_gather_ms_per_step"],3),"wgemm ms/step",round(r["kpconv_weight_gemm_ms_per_step"],3))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32x3" -c 30 -o gpurun_out/${TAG}_gemm \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_ncu_gemm.log 2>&1
ls -la gpurun_out | tail -3
