#!/usr/bin/env python3
"""Collect a statistical reference sample: 8 processes of the UNMODIFIED reference binary (oracle/_ref/superMC_ref.e,
operation 9, 261^2 grid, 12,500 events each, seeds 1000..1007 -- the reference's own 8-process mode) were run by
tests/golden/run_ks_reference.sh <system> under /tmp/ks_<system>_<i>; this script reads their data/sn_ecc_eccp_10.dat and
stores the columns the KS tests need as tests/golden/ks_<system>_ref.npz.  The mean number of tries per accepted event
comes from the oracle port driven by drand48 (bit-identical to the reference's sampler and sweep).

    python tests/golden/make_ks_reference.py pbpb2760 | auau200 | ppb5020 | auau200_kln
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port

SYS = {"pbpb2760": (208, 208, 2760.0, 0.118), "auau200": (197, 197, 200.0, 0.14), "ppb5020": (1, 208, 5020.0, 0.118), "auau200_kln": (197, 197, 200.0, 0.14)}
name = sys.argv[1] if len(sys.argv) > 1 else "pbpb2760"
A, B, ecm, alpha = SYS[name]
pat = "/tmp/ks_run_%d/data/sn_ecc_eccp_10.dat" if (name == "pbpb2760" and os.path.exists("/tmp/ks_run_0")) else "/tmp/ks_" + name + "_%d/data/sn_ecc_eccp_10.dat"
rows = np.concatenate([np.loadtxt(pat % i) for i in range(8)])
cfg = port.make_cfg(ecm=ecm, alpha=alpha)
nA, nB = port.nucleus(A, cfg.width), port.nucleus(B, cfg.width)
st = port.Stream48(seed=4321); tries = acc = 0
while acc < 4000:
    b = np.sqrt(400.0 * st.next())
    p, _ = port.populate(nA, b / 2, 0.0, stream=st); t, _ = port.populate(nB, -b / 2, 0.0, stream=st)
    r = port.collide(cfg, p, t, stream=st); tries += 1
    npart = int((r["ncollA"] > 0).sum() + (r["ncollB"] > 0).sum())
    acc += (r["ncoll"] > 0 and 2 <= npart <= 500)
out = dict(npart=rows[:, 45], ncoll=rows[:, 46], dsdy=rows[:, 47], b=rows[:, 48],
           e2=np.hypot(rows[:, 5], rows[:, 6]), e3=np.hypot(rows[:, 10], rows[:, 11]), e2p=np.hypot(rows[:, 7], rows[:, 8]), r2=rows[:, 9],
           mean_tries=np.array(tries / acc))
for k in out:
    out[k] = out[k].astype(np.float32) if out[k].ndim else out[k]
p = os.path.join(ROOT, "tests", "golden", "ks_%s_ref.npz" % name)
np.savez_compressed(p, **out)
print(p, len(rows), "events", os.path.getsize(p) // 1024, "KB", "mean tries", tries / acc)
