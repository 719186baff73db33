#!/bin/bash
# Round-2 evidence call: training tests, full GPU suite, bench, ncu launch list, full captures of the gather and the tcgen05 GEMM.
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_train.log 2>&1
grep -E "grad parity|\[train\]|passed|failed|FAILED" gpurun_out/${TAG}_train.log | tail -60
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_train_gpu.py > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kpconv_gather -c 14 -o gpurun_out/${TAG}_kpconv \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 60 -o gpurun_out/${TAG}_gemm \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_gemm.log 2>&1
ls -la gpurun_out | grep ${TAG}
