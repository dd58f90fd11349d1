"""Turn the raw ncu outputs of a gpurun call into the committed summaries under profiles/.
usage: python tools/make_profile_summary.py <tag> <launches.csv> <report.ncu-rep> <workload>"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, workload = sys.argv[1:5]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -> per-kernel shares
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("b200::<unnamed>::", "")
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, "%s_launches_%s.md" % (tag, workload)), "w") as f:
    f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), %s\n\n" % workload)
    f.write("Cold-cache, serialised launches: compare SHARES, not absolutes. Raw list: `%s_launches_%s.csv`.\n\n" % (tag, workload))
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if t / tot < 0.0005: continue
        f.write("| `%s` | %d | %.2f | %.1f %% |\n" % (k[:80], n, t / 1e6, 100 * t / tot))
subprocess.check_call(["cp", launches, os.path.join(out, "%s_launches_%s.csv" % (tag, workload))])

# ---- full capture -> key metrics
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
traffic = {}
mode = "a" if len(sys.argv) > 5 and sys.argv[5] == "append" else "w"
with open(os.path.join(out, "%s_ncu_%s.csv" % (tag, workload)), mode) as f:
    w = csv.writer(f)
    if mode == "w": w.writerow(["kernel", "metric", "value", "unit"])
    for r in rr[2:]:
        kn = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        rd = wr = 0.0
        for m in want:
            if m in h:
                i = h.index(m); w.writerow([kn, m, r[i], u[i]])
                if m.startswith("dram__bytes"):
                    v = float(r[i].replace(",", "")); scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[u[i]]
                    if "read" in m: rd = v * scale
                    else: wr = v * scale
        base = kn.split("<")[0]
        traffic[base] = int(rd + wr) if rd == rd and wr == wr and (rd + wr) > 0 else None
tp = os.path.join(out, "traffic.json")
allt = json.load(open(tp)) if os.path.exists(tp) else {}
allt.setdefault(workload, {})
for k, v in traffic.items():
    if v is not None or k not in allt[workload]:
        allt[workload][k] = v
json.dump(allt, open(tp, "w"), indent=1)
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "22"], capture_output=True, text=True).stdout
open(os.path.join(out, "%s_hot_lines_%s.txt" % (tag, workload)), mode).write(hot)
print("wrote profiles for", tag, workload, traffic)
