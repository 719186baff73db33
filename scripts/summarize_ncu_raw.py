"""Condenses an `ncu -i X.ncu-rep --page raw --csv` export (one row per launch, ~2400 columns) into the per-launch table kept under
profiles/. Usage: python scripts/summarize_ncu_raw.py gpurun_out/<tag>_{kpconv,gemm}_raw.csv > profiles/<name>.csv"""
import csv
import re
import sys

COLS = [
    ("kernel", "Kernel Name"), ("grid", "Grid Size"), ("block", "Block Size"),
    ("duration_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
    ("dram_read_B", "dram__bytes_read.sum"), ("dram_write_B", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tf32_ops_pct", "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("tc_smem_wavefronts_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("lsu_smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("tmem_pipe_pct", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("sm_active_cycles", "sm__cycles_active.avg"), ("elapsed_cycles", "sm__cycles_elapsed.max"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]


def col(name):
    if name in hdr:
        return hdr.index(name)
    cand = [i for i, h in enumerate(hdr) if h.endswith(name)]
    return cand[0] if cand else None


idx = [(out, col(src)) for out, src in COLS]
idx = [(o, i) for o, i in idx if i is not None]
w = csv.writer(sys.stdout)
w.writerow([o + (f"[{units[i]}]" if units[i] and o not in ("kernel", "grid", "block") else "") for o, i in idx])
for r in body:
    out = []
    for o, i in idx:
        v = r[i]
        if o == "kernel":
            v = re.sub(r"\(.*", "", v).replace("void ", "").replace("<unnamed>::", "")
        elif o == "duration_us" and units[i] == "ns":
            v = "%.2f" % (float(v.replace(",", "")) / 1e3)
        out.append(v)
    w.writerow(out)
