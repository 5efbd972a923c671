#!/usr/bin/env python3
"""Sample sequences of the UNMODIFIED reference's NBD::rand (oracle/_ref/nbd_probe, built from src/NBD.cpp,
RandomVariable.cpp, TableFunction.cpp, Table.cpp, arsenal.cpp): srand48(7), then 20000 consecutive draws for each
(p, r) pair, one drand48 stream running through all of them.  Build container only; the fixture is committed."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PAIRS = [(0.5714285714285714, 0.75), (0.02, 0.75), (0.9, 0.75), (0.3, 2.5), (0.8, 5.0), (0.05, 0.01), (0.97, 0.3), (0.6, 1.0),
         (0.0344827586206896, 0.75), (0.25, 0.037)]
N, SEED = 20000, 7
args = []
for p, r in PAIRS:
    args += [repr(p), repr(r)]
out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "nbd_probe"), str(SEED), str(N)] + args).decode().split("\n")
samples = np.zeros((len(PAIRS), N), dtype=np.int32); pos = 0
for i in range(len(PAIRS)):
    pos += 1
    samples[i] = [int(x) for x in out[pos:pos + N]]; pos += N
path = os.path.join(ROOT, "tests", "golden", "nbd_ref.npz")
np.savez_compressed(path, pairs=np.array(PAIRS), samples=samples, seed=np.array(SEED))
print(path, os.path.getsize(path) // 1024, "KB")
