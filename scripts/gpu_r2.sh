#!/bin/bash
# Round-2 GPU check: whole suite (achieved parity errors in the log), parity diag, gather micro-benchmark sparse vs dense, bench both arms.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rA 2>&1 | grep -E "\[parity\]|passed|failed|error|Error|FAILED|assert|variant" | head -150 > gpurun_out/${TAG}_pytest.log
tail -70 gpurun_out/${TAG}_pytest.log
timeout 400 python scripts/diag_parity.py > gpurun_out/${TAG}_diag.log 2>&1; tail -40 gpurun_out/${TAG}_diag.log
timeout 200 python scripts/bench_gather.py 2>&1 | tail -17
RDM_GATHER_MODE=dense timeout 200 python scripts/bench_gather.py 2>&1 | tail -17
RDM_GATHER_MODE=sparsew timeout 200 python scripts/bench_gather.py 2>&1 | tail -17
timeout 600 python bench.py --no-reference-gpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 4500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "$2" = "ref" ]; then timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 1500 gpurun_out/${TAG}_ref.json; tail -5 gpurun_out/${TAG}_ref.err; fi
