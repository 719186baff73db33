#!/bin/bash
TAG=${1:-s4l}
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 12 --warmup 4 > gpurun_out/${TAG}_$label.json 2> gpurun_out/${TAG}_$label.err
  python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_$label.json"))
print("$label ms/step",round(l["ms_per_step"],3),"wall",round(l["wall_ms_per_step_incl_flush"],3),"e2e ms",round(l["e2e"]["ms_per_step"],3))
PY
}
run pdl1 RDM_PDL=1
run pdl1_noprof RDM_PDL=1 BENCH_NO_PROF=1
run pdl1_noflush RDM_PDL=1 BENCH_NO_FLUSH=1
run pdl1_noboth RDM_PDL=1 BENCH_NO_FLUSH=1 BENCH_NO_PROF=1
run pdl0 RDM_PDL=0
run pdl0_noprof RDM_PDL=0 BENCH_NO_PROF=1
run pdl0_noflush RDM_PDL=0 BENCH_NO_FLUSH=1
