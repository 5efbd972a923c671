#!/usr/bin/env python3
"""Operation-3 fixture (tests/golden/pbpb2760_avg3.npz) from the UNMODIFIED reference.  Build container only.

Two runs with the same seed and physics parameters:
  1. oracle/_ref/superMC_ref.e operation=3 (the reference's own main and generate_profile_average,
     src/MakeDensity.cpp:736-2103) with every output switch on -> its 48 averaged-profile files
     (sd / ed branch, rotated / reaction-plane, TA*TB, rho_binary, TA, TB, spectator densities; orders 2 and 3);
  2. oracle/_ref/ref_dump with dump_rotate=1 -> the same accepted events (nothing after the collision stage draws
     random numbers, so both programs see the same event sequence) as full-precision nucleon records including the
     state quirk Q4 depends on (stale base boxes, quark offsets).
The script checks that the two runs did see the same events: the mean of ref_dump's rotated order-2 densities must
reproduce the reference's sdAvg_order_2 file to its 12 printed digits.  The lattice is 81 x 81 (dx = 0.25 fm) to keep
the fixture small; nothing on the path depends on the lattice size.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
NEV = 5
PHYS = dict(which_mc_model=5, sub_model=1, Aproj=208, Atarg=208, ecm=2760, alpha=0.118, cc_fluctuation_model=6,
            cc_fluctuation_Gamma_theta=0.75, shape_of_nucleons=2, collision_criterion=2, shape_of_entropy=2,
            maxx=10, maxy=10, dx=0.25, dy=0.25, finalFactor=2.5, bmin=3, bmax=9, Npmin=2, Npmax=500, randomSeed=26,
            average_from_order=2, average_to_order=3, ecc_from_order=1, ecc_to_order=9)
OP3 = dict(operation=3, nev=NEV, use_sd=1, use_ed=1, use_block=1, use_4col=1, output_TATB=1, output_rho_binary=1, output_TA=1,
           output_spectator_density=1, generate_reaction_plane_avg_profile=1, cutdSdy=0)


def workdir(kind):
    run = os.path.join(REFDIR, "run_" + kind)
    work = tempfile.mkdtemp(prefix="avg3_")
    for d in ("tables", "EOS"):
        os.symlink(os.path.join(run, d), os.path.join(work, d))
    shutil.copy(os.path.join(run, "parameters.dat"), work); os.mkdir(os.path.join(work, "data"))
    return work


def main():
    w1 = workdir("rand")
    args = ["%s=%s" % kv for kv in {**PHYS, **OP3}.items()]
    subprocess.check_call([os.path.join(REFDIR, "superMC_ref.e")] + args, cwd=w1, stdout=subprocess.DEVNULL)
    files = sorted(f for f in os.listdir(os.path.join(w1, "data")) if f.endswith("_block.dat"))
    out = {"files": np.array(files)}
    for f in files:
        out["avg/" + f[:-len("_block.dat")]] = np.loadtxt(os.path.join(w1, "data", f))
    # the 4-column writer on an averaged profile, verbatim (format fixture: header carries the last event's Npart)
    out["text_4col_sdAvg_order_2"] = np.frombuffer(open(os.path.join(w1, "data", "sdAvg_order_2_4col.dat"), "rb").read(), dtype=np.uint8)
    out["all_files"] = np.array(sorted(os.listdir(os.path.join(w1, "data"))))
    w2 = workdir("rand")
    binf = os.path.join(w2, "ev.bin")
    subprocess.check_call([os.path.join(REFDIR, "ref_dump"), binf, str(NEV)] + ["%s=%s" % kv for kv in PHYS.items()] + ["dump_rotate=1", "dump_extra=1"],
                          cwd=w2, stdout=subprocess.DEVNULL)
    glob, tries = refio.group_tries(refio.read_records(binf))
    assert len(tries) == NEV
    ff = PHYS["finalFactor"]
    mean = sum(t["rot2/rho"] for t in tries) / NEV * ff
    ref = out["avg/sdAvg_order_2"]
    scale = np.maximum(np.abs(ref), 1e-12 * ref.max())
    assert (np.abs(mean - ref) / scale).max() < 1e-10, "the two reference runs did not see the same events"
    par = {**PHYS, **OP3}
    out.update({"consts": glob["consts"], "params_keys": np.array(list(par.keys())), "params_vals": np.array([float(v) for v in par.values()]),
                "quark_kind": np.array("rand"), "ntries": np.array(NEV), "ecc_rows": np.zeros((0, 49))})
    for it, t in enumerate(tries):
        for k in ("hdr", "proj", "targ", "proj_part", "targ_part", "coll", "proj_x", "targ_x", "spectators"):
            out["t%d/%s" % (it, k)] = t[k]
    path = os.path.join(ROOT, "tests", "golden", "pbpb2760_avg3.npz")
    np.savez_compressed(path, **out)
    print("pbpb2760_avg3: %d events, %d averaged grids, %.1f KB" % (NEV, len(files), os.path.getsize(path) / 1024))
    shutil.rmtree(w1); shutil.rmtree(w2)


if __name__ == "__main__":
    main()
