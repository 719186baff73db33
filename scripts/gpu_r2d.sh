#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|assert|coarse pairs|achieved" | head -40
timeout 300 python scripts/diag_parity.py --synthetic 4k > gpurun_out/${TAG}_diag4k.log 2>&1; tail -32 gpurun_out/${TAG}_diag4k.log
for mode in auto dense sparse; do
RDM_GATHER_MODE=$mode BG_ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:kpconv_gather --csv --log-file gpurun_out/${TAG}_gather_${mode}.csv python scripts/bench_gather.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_gather_${mode}.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID"); gi=hdr.index("Grid Size")
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{"k":r[ki][:44],"g":r[gi]})[r[mi]]=r[vi]
print("mode ${mode}")
tot=0
for k,v in d.items():
    t=float(v["gpu__time_duration.sum"].replace(",",""))/1e3; tot+=t
    print("%-44s grid %-14s %6.1f us act %s el %s inst %s warps %s issue %s"%(v["k"],v["g"],t,v.get("sm__cycles_active.avg"),v.get("sm__cycles_elapsed.max"),v.get("smsp__inst_executed.sum"),v.get("sm__warps_active.avg.pct_of_peak_sustained_active"),v.get("smsp__issue_active.avg.pct_of_peak_sustained_active")))
print("total us",tot)
PY
done
for ps in 1 0; do RDM_GEMM_PRESPLIT=$ps timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_ps$ps.json 2> gpurun_out/${TAG}_bench_ps$ps.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench_ps$ps.json"))
print("presplit $ps value",round(l["value"],1),"ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"e2e",round(l["e2e"]["value"],1),"gather frac",round(l["roofline"]["frac"],3))
PY
tail -2 gpurun_out/${TAG}_bench_ps$ps.err; done
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -22
RDM_LINEAR_USE_REGISTRY=1 timeout 300 python scripts/bench_gemm.py 2>&1 | tail -22
