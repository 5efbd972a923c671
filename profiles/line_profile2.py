#!/usr/bin/env python3
"""Samples / executed warp instructions per CUDA source line straight from an ncu report (needs --import-source on):
usage: line_profile2.py report.ncu-rep kernel-regex [topN]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = None; hdr = None; agg = {}; stall = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fn = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        try: s = int(r[si]); n = int(r[ii])
        except ValueError: continue
        a = agg.setdefault((fn, int(r[0]), r[1].strip()[:110]), [0, 0]); a[0] += s; a[1] += n
        st = stall.setdefault((fn, int(r[0])), {})
        for i in sc:
            try: st[hdr[i]] = st.get(hdr[i], 0) + int(r[i])
            except ValueError: pass
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (ts, ti))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(stall[(k[0], k[1])].items(), key=lambda x: -x[1])[:3]
    print("%5.2f%% smp %5.2f%% inst %s:%d  [%s]  %s" % (100.0 * v[0] / ts, 100.0 * v[1] / ti, k[0], k[1], " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in st), k[2]))
