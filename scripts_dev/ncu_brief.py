"""Brief summary of an .ncu-rep: key metrics + top stall lines of the SASS page."""
import csv, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, units = rd[0], rd[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for row in rd[2:]:
    print("-" * 60)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-80s %s %s" % (w, row[i][:90], units[i]))
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(row[i] or 0) > 0.15:
            print("   stall %-60s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), row[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if his:
    hi = his[0]
    h2 = rows[hi]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    data = rows[hi + 1:end]
    iS, iSrc, iE = h2.index("Warp Stall Sampling (All Samples)"), h2.index("Source"), h2.index("Instructions Executed")
    sc = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
    print("SASS lines %d total samples %d" % (len(data), tot))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS]) if data[i][iS].isdigit() else 0)[:ntop]
    for i in sorted(idx):
        r = data[i]
        st = sorted(((h2[c][6:], int(r[c])) for c in sc if r[c].isdigit() and int(r[c]) > 0), key=lambda kv: -kv[1])[:2]
        print("%5d %6s %8s  %-60s %s" % (i, r[iS], r[iE], r[iSrc].strip()[:60], st))
