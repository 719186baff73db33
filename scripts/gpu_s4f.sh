#!/bin/bash
TAG=${1:-s4f}
mkdir -p gpurun_out
./scripts/ubench/ffma_bench > gpurun_out/${TAG}_ffma.log 2>&1; cat gpurun_out/${TAG}_ffma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kpconv_gather -c 14 -o gpurun_out/${TAG}_kpconv \
    python scripts/bench_gather.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -3
