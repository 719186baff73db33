#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|assert|voxel_downsample" | head -30
for hf in 1 0; do
RDM_GATHER_HEAVY_FIRST=$hf BG_ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:kpconv_gather --csv --log-file gpurun_out/${TAG}_gather_hf$hf.csv python scripts/bench_gather.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_gather_hf$hf.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID"); gi=hdr.index("Grid Size")
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{"k":r[ki][:40],"g":r[gi]})[r[mi]]=r[vi]
tot=0; out=[]
for k,v in d.items():
    t=float(v["gpu__time_duration.sum"].replace(",",""))/1e3; tot+=t; out.append("%.1f(%s/%s)"%(t,v.get("sm__cycles_active.avg","").split(".")[0],v.get("sm__cycles_elapsed.max")))
print("heavy_first=$hf total us %.1f"%tot, " ".join(out))
PY
done
for hf in 1 0; do RDM_GATHER_HEAVY_FIRST=$hf timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_hf$hf.json 2> gpurun_out/${TAG}_bench_hf$hf.err
python - <<PY
import json
l=json.load(open("gpurun_out/${TAG}_bench_hf$hf.json"))
print("heavy_first $hf value",round(l["value"],1),"ms/step",round(l["ms_per_step"],3),l.get("step_ms_stats"),"e2e",round(l["e2e"]["value"],1),"gather frac",round(l["roofline"]["frac"],3))
PY
tail -2 gpurun_out/${TAG}_bench_hf$hf.err; done
