#!/bin/bash
# Runs on the GPU box (under gpurun): tests, smoke, bench, ncu launch list + full capture of the KPConv gather kernel.
# Usage: scripts/gpu_profile.sh <tag>     outputs -> gpurun_out/<tag>_*
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
# launch list of the same command (short run): per-launch device time, cold-cache + serialised -> compare SHARES
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --pairs 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full capture of the KPConv gather kernel: the 14 launches of one step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kpconv_gather -c 14 -o gpurun_out/${TAG}_kpconv \
    python bench.py --steps 1 --warmup 0 --pairs 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
