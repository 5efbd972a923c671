#!/usr/bin/env python3
"""Text-format fixtures: the reference's own writers (dumpDensityBlock/4Col, dumpEccentricities at the
stock 8-digit precision, dumpparticipantTable, dumpBinaryTable, dumpSpectatorsTable) run on two events,
next to the full-precision numbers they were given.  Build container only (needs oracle/_ref)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio
REF = os.path.join(ROOT, "oracle", "_ref"); RUN = os.path.join(REF, "run_rand")      # random valence-quark offsets: quarks.data is not trivial
ARGS = ["which_mc_model=5", "sub_model=1", "Aproj=208", "Atarg=208", "ecm=2760", "alpha=0.118", "maxx=8", "maxy=8", "dx=0.5", "dy=0.5",
        "finalFactor=40", "randomSeed=31", "cc_fluctuation_model=6", "dump_grids=1", "dump_extra=1", "dump_text=1", "bmax=10"]
out = {}
for exe, tag in (("ref_dump_stock", "stock"), ("ref_dump", "hi")):
    for f in os.listdir(os.path.join(RUN, "data")):
        os.remove(os.path.join(RUN, "data", f))
    subprocess.check_call([os.path.join(REF, exe), "/tmp/text_%s.bin" % tag, "2"] + ARGS, cwd=RUN, stdout=subprocess.DEVNULL)
    if tag == "stock":
        for f in sorted(os.listdir(os.path.join(RUN, "data"))):
            out["file/" + f] = np.frombuffer(open(os.path.join(RUN, "data", f), "rb").read(), dtype=np.uint8)
    else:
        out["ecc_rows"] = np.loadtxt(os.path.join(RUN, "data", "h_ecc_10.dat")).reshape(-1, 49)
        glob, tries = refio.group_tries(refio.read_records("/tmp/text_hi.bin"))
        out["consts"] = glob["consts"]
        for i, t in enumerate(tries):
            for k in ("hdr", "proj", "targ", "proj_part", "targ_part", "coll", "rho", "spectators", "proj_x", "targ_x"):
                out["t%d/%s" % (i, k)] = t[k]
out["args"] = np.array(ARGS)
p = os.path.join(ROOT, "tests", "golden", "text_formats.npz")
np.savez_compressed(p, **out)
print(p, os.path.getsize(p) // 1024, "KB", [k for k in out if k.startswith("file/")])
