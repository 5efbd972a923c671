#!/usr/bin/env python3
"""Fixture for the reference's 3-D extension (tests/golden/profile3d.npz).  Build container only.
oracle/_ref/ref_profile3d = /root/reference/scripts/generate_3d_profiles/{profile_3d,Regge96}.cpp behind a harness main:
participants of a golden Pb+Pb event -> (eta_s, x, y) Gaussians, random_flag 1 (rapidities drawn) and 3 (widths drawn too);
the harness also writes the rapidities and widths the reference drew (time-seeded), so the comparison is deterministic."""
import os, struct, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden
g = Golden("pbpb2760_glb")
t = [t for t in g.tries() if int(t["hdr"][4]) and 60 < int(t["hdr"][2]) + int(t["hdr"][3]) < 200][0]
rows = [(t["proj"][i, 0], t["proj"][i, 1], 1) for i in t["proj_part"].astype(int)] + [(t["targ"][i, 0], t["targ"][i, 1], 2) for i in t["targ_part"].astype(int)]
NX, NY, NETA, DX, DY, DETA, ECM = 49, 49, 33, 0.5, 0.5, 0.4, 19.6
out = {"grid": np.array([NX, NY, NETA, DX, DY, DETA, ECM])}
d = tempfile.mkdtemp()
with open(os.path.join(d, "part.dat"), "w") as f:
    for x, y, i in rows:
        f.write("%10.3g   %10.3g   %d\n" % (x, y, i))            # the format of ParticipantTable_event_<k>.dat
with open(os.path.join(d, "bin.dat"), "w") as f:
    for c in t["coll"]:
        f.write("%10.3g%10.3g\n" % (c[0], c[1]))
for flag in (0, 1, 3):
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_profile3d"), "part.dat", "bin.dat", str(NX), str(NY), str(NETA), str(DX), str(DY), str(DETA),
                           str(ECM), str(flag), "o.bin"], cwd=d, stdout=subprocess.DEVNULL)
    buf = open(os.path.join(d, "o.bin"), "rb").read()
    (n,) = struct.unpack_from("<q", buf, 0)
    src = np.frombuffer(buf, dtype="<f8", count=7 * n, offset=8).reshape(n, 7).copy()
    rho = np.frombuffer(buf, dtype="<f8", count=NETA * NX * NY, offset=8 + 56 * n).reshape(NETA, NX, NY).copy()
    out["src_%d" % flag] = src; out["rho_%d" % flag] = rho
    if flag == 1:
        out["text_rhob"] = np.frombuffer(open(os.path.join(d, "ref_rhob.dat"), "rb").read(), dtype=np.uint8)      # output_3d_rhob_profile verbatim
p = os.path.join(ROOT, "tests", "golden", "profile3d.npz")
np.savez_compressed(p, **out)
print(p, os.path.getsize(p) // 1024, "KB", len(rows), "participants")
