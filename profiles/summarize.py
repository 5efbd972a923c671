#!/usr/bin/env python3
"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.
usage: summarize.py <tag>   (reads gpurun_out/<tag>_prof.ncu-rep, <tag>_launches.csv, <tag>_bench.json ...)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
def ncu_summary(rep, suffix):
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
  summ = []; traffic = {}
  nev = 2048
  for vals in rows[2:]:
    ncu_row(dict(zip(hdr, vals)), dict(zip(hdr, units)), summ, traffic, nev)
  json.dump(summ, open(os.path.join(P, tag + "_ncu_full_summary" + suffix + ".json"), "w"), indent=1)
  json.dump(traffic, open(os.path.join(P, tag + "_dram_traffic" + suffix + ".json"), "w"), indent=1)
  return traffic

def ncu_row(d, u, summ, traffic, nev):
  if True:
    k = {"kernel": d["Kernel Name"].split("(")[0].replace("void smc::", "").replace("<0>", "")}
    for key in KEYS:
        if key in d: k[key] = d[key] + (" " + u[key] if u.get(key) else "")
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v.replace(",", "")) for h, v in d.items()
          if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
    k["top_stalls_per_issue"] = dict(sorted(st.items(), key=lambda x: -x[1])[:5])
    summ.append(k)
    def num(x, unit):
        v = float(x.replace(",", "")); m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]; return v * m
    tb = num(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    traffic[k["kernel"].split("<")[0].replace("void ", "").strip()] = {"dram_bytes_per_launch": tb, "events_per_launch": nev, "dram_bytes_per_event": tb / nev}
traffic = ncu_summary(os.path.join(G, tag + "_prof.ncu-rep"), "")
if os.path.exists(os.path.join(G, tag + "_prof_kln.ncu-rep")):
    ncu_summary(os.path.join(G, tag + "_prof_kln.ncu-rep"), "_kln")
# launch list: per-kernel share of the bench step
agg = {}
for r in csv.DictReader(l for l in open(os.path.join(G, tag + "_launches.csv")) if l.startswith('"')):
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    n = r["Kernel Name"].split("(")[0].replace("void smc::", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(r["Metric Unit"], 1e-6)
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, tag + "_launch_shares.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 200 ... python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms_per_step, not absolutes\n")
    for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-28s launches %4d  total %9.3f ms  share %5.1f %%\n" % (n, v[0], v[1], 100 * v[1] / tot))
for f in (tag + "_launches.csv", tag + "_bench.json", tag + "_bench_reference.json", tag + "_bench_kln.json", tag + "_tests.txt"):
    if os.path.exists(os.path.join(G, f)):
        open(os.path.join(P, f), "w").write(open(os.path.join(G, f)).read())
print(open(os.path.join(P, tag + "_launch_shares.txt")).read())
print(json.dumps(traffic, indent=1))
