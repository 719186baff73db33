#!/bin/bash
# training tests, precision mode (config 3) tests, GEMM micro-benchmark in both precision modes, bench in tf32 mode, train bench at N=1
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_precision_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
grep -E "\[precision\]|\[train\]|passed|failed|FAILED|Error" gpurun_out/${TAG}_tests.log | tail -60
RDM_LINEAR_USE_REGISTRY=1 timeout 300 python scripts/bench_gemm.py > gpurun_out/${TAG}_gemm_fp32.txt 2>&1; tail -4 gpurun_out/${TAG}_gemm_fp32.txt
RDM_PRECISION=tf32 timeout 300 python scripts/bench_gemm.py > gpurun_out/${TAG}_gemm_tf32.txt 2>&1; cat gpurun_out/${TAG}_gemm_tf32.txt
timeout 900 python bench.py --precision tf32 --no-cpu-baseline --no-reference-gpu > gpurun_out/${TAG}_bench_tf32.json 2> gpurun_out/${TAG}_bench_tf32.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench_tf32.json"))
print("tf32 value",round(l["value"],1),"ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"e2e",round(l["e2e"]["value"],1),"gather frac",round(l["roofline"]["frac"],3), l.get("roofline_gemm"), l.get("pose_vs_synthetic_gt"))
PY
tail -3 gpurun_out/${TAG}_bench_tf32.err
timeout 900 python bench.py --train --steps 20 --warmup 3 > gpurun_out/${TAG}_train_n1.json 2> gpurun_out/${TAG}_train_n1.err
cat gpurun_out/${TAG}_train_n1.json; tail -5 gpurun_out/${TAG}_train_n1.err
