#!/usr/bin/env python3
"""Join an ncu `--page source --csv` dump with `nvdisasm -g -c` line info: samples and executed
instructions per CUDA source line.  usage: line_profile.py src.csv dis.txt mangled_kernel_name [topN] [kernel-name substring in the csv]"""
import csv, re, sys
src_csv, dis, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
want = sys.argv[5] if len(sys.argv) > 5 else ""
# offset -> (file, line) from nvdisasm
off2line = {}; cur = None; inside = False
for ln in open(dis, errors="ignore"):
    if ln.startswith(mangled + ":") or ln.startswith(".text." + mangled + ":"):
        inside = True; continue
    if inside and ln.startswith("//--------------------- .text.") and mangled not in ln:
        inside = False
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
i = 0; best = None
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        hdr = rows[i + 1]; j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): data.append(rows[j])
            j += 1
        si = hdr.index("# Samples"); tot = sum(int(r[si] or 0) for r in data)
        if want in rows[i][1] and (best is None or tot > best[0]): best = (tot, hdr, data)
        i = j
    else: i += 1
tot, hdr, data = best
ai, si, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = min(int(r[ai], 16) for r in data)
agg = {}
for r in data:
    key = off2line.get(int(r[ai], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0, 0]); a[0] += int(r[si] or 0); a[1] += int(r[ii] or 0)
tinst = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (tot, tinst))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% samples %6.2f%% inst  %s:%d" % (100.0 * v[0] / tot, 100.0 * v[1] / tinst, key[0], key[1]))
