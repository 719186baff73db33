#!/bin/bash
TAG=${1:-s4m}
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/${TAG}_$label.json 2> gpurun_out/${TAG}_$label.err
  python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_$label.json"))
print("$label ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"wall",round(l["wall_ms_per_step_incl_flush"],3),"e2e ms",round(l["e2e"]["ms_per_step"],3))
PY
}
for r in 1 2 3; do run pdl1_$r RDM_PDL=1; done
run pdl0_1 RDM_PDL=0
