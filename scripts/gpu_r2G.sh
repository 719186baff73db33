#!/bin/bash
TAG=${1:-r2G}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_full.log 2>&1
tail -2 gpurun_out/${TAG}_pytest_full.log; grep -E "^FAILED" gpurun_out/${TAG}_pytest_full.log | head
grep -E "^\[parity\]|^\[grad parity\]|^\[precision\]|^\[train\]|passed|failed" gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --precision tf32 --no-cpu-baseline --no-reference-gpu > gpurun_out/${TAG}_bench_tf32.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 --no-reference-gpu > gpurun_out/${TAG}_ref.json 2>/dev/null
python - <<PY
import json
for f in ("bench","bench_tf32"):
    l=json.load(open("gpurun_out/${TAG}_%s.json"%f)); s=l["step_ms_stats"]
    print(f,"value",round(l["value"],1),"median",round(s["median"],3),"max",round(s["max"],1),"e2e",round(l["e2e"]["value"],1),l["e2e"]["result_interval_ms"],"gather frac",round(l["roofline"]["frac"],3),l.get("vs_reference_gpu_eager"),l.get("parity_vs_cpu"))
print(open("gpurun_out/${TAG}_ref.json").read()[:300])
PY
