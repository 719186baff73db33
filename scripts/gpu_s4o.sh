#!/bin/bash
TAG=${1:-s4o}
mkdir -p gpurun_out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_$label.json 2> gpurun_out/${TAG}_$label.err
  python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_$label.json"))
print("$label value",round(l["value"],1),"ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"e2e",round(l["e2e"]["value"],1),l["clocks"])
PY
  tail -2 gpurun_out/${TAG}_$label.err
}
for r in 1 2 3 4; do run nvml_$r BENCH_CLOCKS=nvml; done
run noprof_1 BENCH_NO_PROF=1; run noprof_2 BENCH_NO_PROF=1
