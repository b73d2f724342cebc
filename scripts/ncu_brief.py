"""Prints a short per-kernel digest of an ncu --set full report (developer tool): ncu_brief.py <rep.ncu-rep> [kernel substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__mio_inst_issued.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if not stalls:
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith(".pct")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if filt not in name:
        continue
    print("====", name.split("(")[0])
    for w in WANT:
        if w in hdr:
            print(f"  {w:72s} {r[hdr.index(w)]}")
    st = []
    for h in stalls:
        try:
            st.append((float(r[hdr.index(h)]), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:6]:
        print(f"  stall {h:66s} {v:.3f}")
