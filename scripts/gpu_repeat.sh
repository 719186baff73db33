#!/bin/bash
# Repeatability of the bench line (GPU box): N short runs, prints value / e2e / per-step min-median-max.  Usage: scripts/gpu_repeat.sh <tag> [N] [ENV=VAL ...]
TAG=${1:-rep}; N=${2:-3}; shift; shift
mkdir -p gpurun_out
for r in $(seq 1 $N); do
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_$r.json 2> gpurun_out/${TAG}_$r.err
  python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_$r.json"))
print("run $r value",round(l["value"],1),"ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"e2e",round(l["e2e"]["value"],1),"gather frac",round(l["roofline"]["frac"],3),l["clocks"])
PY
  tail -2 gpurun_out/${TAG}_$r.err
done
