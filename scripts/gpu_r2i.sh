#!/bin/bash
# Round-2 evidence call: training tests, bench, ncu launch list, full captures (exported to CSV on the box: the .ncu-rep files
# exceed the 64 MiB return limit) of the KPConv gather and the tcgen05 GEMM at the benched pair size.
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_train.log 2>&1
grep -E "grad parity\] (sink|Trans)|\[train\]|passed|failed|FAILED|Error" gpurun_out/${TAG}_train.log | tail -40
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kpconv_gather -c 14 -o /tmp/${TAG}_kpconv \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_kpconv.ncu-rep --page raw --csv > gpurun_out/${TAG}_kpconv_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 40 -o /tmp/${TAG}_gemm \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline --no-reference-gpu > gpurun_out/${TAG}_ncu_gemm.log 2>&1
ncu -i /tmp/${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_raw.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_launches.csv
ls -la gpurun_out | grep ${TAG}
