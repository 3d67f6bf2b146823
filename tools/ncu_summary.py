"""Compact per-kernel summary of an `ncu --page raw --csv` export: python tools/ncu_summary.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("dur_us", "gpu__time_duration.sum"), ("rd_MB", "dram__bytes_read.sum"), ("wr_MB", "dram__bytes_write.sum"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("L1hit", "l1tex__t_sector_hit_rate.pct"),
        ("L2hit", "lts__t_sector_hit_rate.pct"), ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("fp64%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("lsu%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("inst", "smsp__inst_executed.sum"), ("occ_lim_reg", "launch__occupancy_limit_registers"), ("occ_lim_smem", "launch__occupancy_limit_shared_mem")]
stall = [(h.split("issue_stalled_")[1].split("_per_warp_active")[0], i) for i, h in enumerate(hdr)
         if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]


def conv(v, u):
    try:
        x = float(v)
    except ValueError:
        return v
    if u == "Gbyte":
        x *= 1000
    if u == "ms":
        x *= 1000
    if u == "Kbyte":
        x /= 1000
    return f"{x:.4g}"


for r in data:
    name = r[hdr.index("Kernel Name")][:40]
    out = [name]
    for lab, key in want:
        if key in hdr:
            i = hdr.index(key)
            out.append(f"{lab}={conv(r[i], units[i])}")
    st = []
    for lab, i in stall:
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v >= 0.4:
            st.append((v, lab.replace("_per_issue_active.ratio", "")))
    out.append("stalls: " + " ".join(f"{l}={v:.1f}" for v, l in sorted(st, reverse=True)[:6]))
    print(" ".join(out))
