#!/usr/bin/env python3
"""Create the run directory the reference binaries (oracle/_ref/superMC_ref.e, ref_dump) need.

TEST INFRASTRUCTURE ONLY.  The reference reads, relative to its cwd: parameters.dat,
EOS/hotQCD/hrg_hotqcd_eos_binary.dat (src/EOS.cpp:92-122, ctor exits without it), tables/*.dat
(src/Nucleus.cpp:37-48,383-522) and appends to data/.  tables/QuarkPos.txt is a missing blob
upstream (.MISSING_LARGE_BLOBS) and is dereferenced for every nucleon (src/Particle.cpp:38-41), so a
synthetic 250,000-row stand-in is generated here:
  quark=zero   r1=r2=0  -> every nucleon AABB is exactly +-4w   (strict parity default)
  quark=rand   seeded pseudo-random (r1, r2, cos theta12)      -> exercises the AABB-union logic
Everything lands under oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
"""
import os, shutil, sys
import numpy as np

REF = os.environ.get("SMC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def make(run, quark="zero"):
    os.makedirs(os.path.join(run, "data"), exist_ok=True)
    os.makedirs(os.path.join(run, "tables"), exist_ok=True)
    os.makedirs(os.path.join(run, "EOS", "hotQCD"), exist_ok=True)
    shutil.copy(os.path.join(REF, "parameters.dat"), run)
    shutil.copy(os.path.join(REF, "EOS/hotQCD/hrg_hotqcd_eos_binary.dat"), os.path.join(run, "EOS/hotQCD"))
    for f in ("he3_plaintext.dat", "he4_plaintext.dat", "carbon_plaintext.dat"):
        shutil.copy(os.path.join(REF, "tables", f), os.path.join(run, "tables"))
    write_quarkpos(os.path.join(run, "tables", "QuarkPos.txt"), quark)


def quark_table(kind):
    if kind == "zero":
        return np.zeros((250000, 3))
    rng = np.random.default_rng(20240607)
    t = np.empty((250000, 3))
    t[:, 0] = rng.gamma(3.0, 0.25, 250000)          # r1 / R
    t[:, 1] = rng.gamma(3.0, 0.25, 250000)          # r2 / R
    t[:, 2] = rng.uniform(-1.0, 1.0, 250000)        # cos(theta12)
    return np.round(t, 6)


def write_quarkpos(path, kind):
    np.savetxt(path, quark_table(kind), fmt="%.6g")


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "zero"
    run = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "_ref", "run_" + kind)
    make(run, kind)
    print(run)
