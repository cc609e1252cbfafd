"""Summarise gpurun_out/launches.csv (per-launch durations) and a full ncu report into profiles/."""
import collections, csv, json, re, subprocess, sys, os
tag = sys.argv[1]
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)
with open("gpurun_out/launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(list)
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot[re.sub(r"\(.*", "", r["Kernel Name"])[:64]].append(v)
S = sum(sum(v) for v in tot.values())
txt = ["ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ : python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline",
       "(own kernels of the whole process: 2-iteration warm-up solver + 3-iteration run + 23 isolated SpMM launches; cold-cache, serialised - compare shares)",
       "total %.2f ms over %d launches" % (S / 1e3, sum(len(v) for v in tot.values())), ""]
for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
    vs = sorted(v)
    txt.append("%-64s n=%4d total %9.1f us  median %8.1f us  max %8.1f us  share %5.1f%%" % (k, len(v), sum(v), vs[len(vs) // 2], vs[-1], 100 * sum(v) / S))
open(os.path.join(out_dir, tag + "_launches.txt"), "w").write("\n".join(txt) + "\n")
print("\n".join(txt))
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_dia.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, units = rd[0], rd[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
res = []
for row in rd[2:]:
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = (row[i] + " " + units[i]).strip()
    res.append(d)
json.dump(res, open(os.path.join(out_dir, tag + "_spmm_dia_ncu_full.json"), "w"), indent=1)
for d in res[:1]:
    for k, v in d.items():
        print("%-70s %s" % (k, v))
