#!/bin/bash
# timeline + (optional) ncu full capture of the gather kernels.  Usage: scripts/gpu_tl.sh <tag> [ncu]
TAG=${1:-tl}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -8
timeout 300 python scripts/timeline.py $TAG 4 > gpurun_out/${TAG}_timeline.log 2>&1 || tail -5 gpurun_out/${TAG}_timeline.log
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kpconv_gather -c 14 -o gpurun_out/${TAG}_kpconv \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -6
