#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_dump, built from
/root/reference/src by oracle/ref_build/Makefile).  Run in the build container only; the fixtures are
committed so the GPU box (where /root/reference does not exist) can check against them.

    python tests/golden/make_golden.py            # all systems
    python tests/golden/make_golden.py NAME ...   # some; SMC_GOLDEN_REUSE=1 re-reads /tmp/golden_<name>.bin if present

Records of rejected tries are kept up to the KEEP_TRIES-th accepted event only (collision parity on Ncoll == 0 tries
needs a few of them, not all); every accepted event keeps its full record.

Per system the fixture holds, for every try of the reference's rejection loop: the sorted nucleon rows
(x y z xL xR yL yR ncoll weight), the drand48 state before getBinaryCollision, the collision list
(x y weight additional_weight i j), participant orders, and for accepted events the 49-column
eccentricity row at 17 significant digits, sum(rho), the hot-spot region, and (first event only) the
full TA1/TA2/rho/rho_binary/spectator grids.
"""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")

COMMON = dict(maxx=13, maxy=13, dx=0.1, dy=0.1, finalFactor=1, bmin=0, bmax=20, Npmin=2, Npmax=500,
              shape_of_nucleons=2, collision_criterion=2, shape_of_entropy=2, cc_fluctuation_model=6)
NEV = 32          # accepted events per system (>= 32)
KEEP_TRIES = 8    # rejected tries are stored while fewer than this many events have been accepted
SYSTEMS = {
    # name: (quark table kind, n accepted events, n events with full grids, parameters)
    "pbpb2760_glb": ("zero", NEV, 1, dict(which_mc_model=5, sub_model=1, Aproj=208, Atarg=208, ecm=2760, alpha=0.118,
                                       cc_fluctuation_Gamma_theta=0.75, randomSeed=11)),
    "auau200_glb_quarks": ("rand", NEV, 1, dict(which_mc_model=5, sub_model=1, Aproj=197, Atarg=197, ecm=200, alpha=0.14,
                                             cc_fluctuation_Gamma_theta=0.61, randomSeed=12)),
    "ppb5020_glb_quarks": ("rand", NEV, 1, dict(which_mc_model=5, sub_model=1, Aproj=1, Atarg=208, ecm=5020, alpha=0.118,
                                             cc_fluctuation_Gamma_theta=0.75, randomSeed=13)),
    "pbpb2760_sqrt_disk": ("zero", NEV, 0, dict(which_mc_model=7, sub_model=1, Aproj=208, Atarg=208, ecm=2760,
                                             collision_criterion=1, randomSeed=14)),
    "pbpb2760_uli": ("zero", NEV, 0, dict(which_mc_model=5, sub_model=2, Aproj=208, Atarg=208, ecm=2760, alpha=0.118,
                                       randomSeed=15)),
    "auau200_disk_nucleons": ("zero", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=197, Atarg=197, ecm=200, alpha=0.14,
                                                 shape_of_nucleons=1, shape_of_entropy=1, collision_criterion=1,
                                                 cc_fluctuation_model=0, randomSeed=16)),
    "he3au200_glb": ("rand", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=3, Atarg=197, ecm=200, alpha=0.14,
                                        cc_fluctuation_Gamma_theta=0.61, randomSeed=17)),
    "cc200_glb": ("zero", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=12, Atarg=12, ecm=200, alpha=0.14,
                                     cc_fluctuation_Gamma_theta=0.61, randomSeed=18)),
    "uu193_deformed": ("zero", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=238, Atarg=238, ecm=193, alpha=0.14,
                                          proj_deformed=1, targ_deformed=1, randomSeed=19)),
    "cuau200_glb": ("rand", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=63, Atarg=197, ecm=200, alpha=0.14,
                                       cc_fluctuation_Gamma_theta=0.61, randomSeed=23)),
    "pbpb5020_lambda_width": ("zero", NEV, 0, dict(which_mc_model=5, sub_model=1, Aproj=208, Atarg=208, ecm=5020, alpha=0.118,
                                                 shape_of_nucleons=3, gaussian_lambda=4.14, cc_fluctuation_Gamma_theta=0.75, randomSeed=22)),
    "auau200_kln": ("zero", 16, 1, dict(which_mc_model=1, sub_model=7, Aproj=197, Atarg=197, ecm=200, tmax=24, tmax_subdivision=3,
                                       cc_fluctuation_model=0, randomSeed=21, bmin=8)),
    # MC-KLN Pb+Pb 2.76 TeV, lambda = 0.138 (scripts/generateAvgprofile.py:211-222): the reference's full 211^2 BASES table
    # (MCnucl.cpp:911-960) + minimum-bias events; the second run holds central events (its table must be the same bits)
    "pbpb2760_kln": ("zero", 12, 1, dict(which_mc_model=1, sub_model=7, Aproj=208, Atarg=208, ecm=2760, tmax=71, tmax_subdivision=3,
                                         cc_fluctuation_model=0, randomSeed=31, **{"lambda": 0.138})),
    "pbpb2760_kln_central": ("zero", 4, 1, dict(which_mc_model=1, sub_model=7, Aproj=208, Atarg=208, ecm=2760, tmax=71, tmax_subdivision=3,
                                                cc_fluctuation_model=0, randomSeed=32, bmax=3.5, **{"lambda": 0.138})),
    # three rapidity slices (ny = 3, ymax = 2: y = -2, -2/3, 2/3 -- MCnucl.cpp:932 divides by ny, not ny - 1): one table per slice
    "auau200_kln_ny3": ("zero", 6, 1, dict(which_mc_model=1, sub_model=7, Aproj=197, Atarg=197, ecm=200, tmax=24, tmax_subdivision=3,
                                           cc_fluctuation_model=0, randomSeed=27, bmin=8, ny=3, ymax=2)),
    # valence-quark substructure (Particle::getFluctuatedDensity, GaussianNucleonsCal::testFluctuatedCollision): entropy from three
    # quark Gaussians per wounded nucleon with one Gamma weight each; second system: the quark-overlap hit test as well
    "pbpb2760_quarks": ("rand", 8, 1, dict(which_mc_model=5, sub_model=1, Aproj=208, Atarg=208, ecm=2760, alpha=0.118, shape_of_entropy=3,
                                           cc_fluctuation_Gamma_theta=0.75, randomSeed=28)),
    "auau200_quarkhit": ("rand", 8, 1, dict(which_mc_model=5, sub_model=1, Aproj=197, Atarg=197, ecm=200, alpha=0.14, shape_of_entropy=3,
                                            collision_criterion=3, cc_fluctuation_Gamma_theta=0.61, randomSeed=29)),
    # table-driven nuclei with synthetic configuration files in the reference's formats (tests/table_synth.py; the real files
    # are missing blobs upstream): O+O (Nucleus.cpp:462-478,555-574) and NN-correlated Au (Nucleus.cpp:481-522,623-666)
    "oo200_glb": ("rand", NEV, 1, dict(which_mc_model=5, sub_model=1, Aproj=16, Atarg=16, ecm=200, alpha=0.14,
                                       cc_fluctuation_Gamma_theta=0.61, randomSeed=24)),
    "auau200_nncorr": ("rand", NEV, 1, dict(which_mc_model=5, sub_model=1, Aproj=197, Atarg=197, ecm=200, alpha=0.14,
                                            include_NN_correlation=1, cc_fluctuation_Gamma_theta=0.61, randomSeed=25)),
    "pbpb2760_rotate": ("rand", 6, 2, dict(which_mc_model=5, sub_model=1, Aproj=208, Atarg=208, ecm=2760, alpha=0.118,
                                           cc_fluctuation_Gamma_theta=0.75, randomSeed=20, bmax=12, dump_rotate=1)),
}


def run_system(name):
    kind, nev, ngrid, par = SYSTEMS[name]
    run = os.path.join(REFDIR, "run_" + kind)
    p = dict(COMMON); p.update(par)
    p.update(dump_grids=1, dump_extra=1, dump_tries=1)
    if "lambda" not in p and p.get("which_mc_model") == 1:
        p["lambda"] = 0.218
    args = ["%s=%s" % kv for kv in p.items()]
    binf = "/tmp/golden_%s.bin" % name
    eccf = "/tmp/golden_%s.ecc" % name
    if not (os.environ.get("SMC_GOLDEN_REUSE") and os.path.exists(binf) and os.path.exists(eccf)):
        import shutil, tempfile
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import table_synth
        table_synth.install(os.path.join(run, "tables"), p)          # synthetic O / NN-correlated files when the system needs them
        # a private copy of the run directory's data/ so that several systems can be generated concurrently
        work = tempfile.mkdtemp(prefix="golden_%s_" % name)
        for d in ("tables", "EOS"):
            os.symlink(os.path.join(run, d), os.path.join(work, d))
        shutil.copy(os.path.join(run, "parameters.dat"), work); os.mkdir(os.path.join(work, "data"))
        subprocess.check_call([os.path.join(REFDIR, "ref_dump"), binf, str(nev)] + args, cwd=work, stdout=subprocess.DEVNULL)
        shutil.copy(os.path.join(work, "data", "h_ecc_10.dat"), eccf)
        shutil.rmtree(work)
    rec = refio.read_records(binf)
    glob, tries = refio.group_tries(rec)
    ncol = 53 if (p.get("proj_deformed") or p.get("targ_deformed")) else 49   # deformed rows carry 4 uninitialised extras (quirk Q9)
    ecc = np.loadtxt(eccf).reshape(-1, ncol)[:, :49]
    out = {k: glob[k] for k in glob if k.startswith("kln_")}
    ny = int(p.get("ny", 1))
    if name == "pbpb2760_kln_central":         # same table as the minimum-bias run (deterministic BASES seed): keep one copy
        assert np.array_equal(np.load(os.path.join(ROOT, "tests", "golden", "pbpb2760_kln.npz"))["kln_table"], out.pop("kln_table"))
    out.update({"consts": glob["consts"], "params_keys": np.array(list(p.keys())), "params_vals": np.array([float(v) for v in p.values()]),
           "quark_kind": np.array(kind), "ntries": np.array(len(tries)), "ecc_rows": ecc})
    ia = 0
    tries = [t for t in tries]
    kept = []
    for t in tries:                         # drop the records of rejected tries once KEEP_TRIES events are in
        acc = int(t["hdr"][4])
        if acc or sum(int(q["hdr"][4]) for q in kept) < KEEP_TRIES:
            kept.append(t)
    tries = kept
    out["ntries"] = np.array(len(tries))
    for it, t in enumerate(tries):
        pre = "t%d/" % it
        acc = int(t["hdr"][4])
        for k in ("hdr", "proj", "targ", "proj_part", "targ_part", "coll"):
            out[pre + k] = t[k]
        if par.get("dump_rotate") or int(p.get("shape_of_entropy", 2)) == 3:
            out[pre + "proj_x"] = t["proj_x"]; out[pre + "targ_x"] = t["targ_x"]
        if int(p.get("shape_of_entropy", 2)) == 3:
            out[pre + "proj_qf"] = t["proj_qf"]; out[pre + "targ_qf"] = t["targ_qf"]
        if acc:
            out[pre + "dndy"] = t["dndy"]; out[pre + "region"] = t["region"]; out[pre + "spectators"] = t["spectators"]
            out[pre + "ecc_index"] = np.array(ia * ny)          # ny rows per event, slice-major
            if ia < ngrid:
                for k in ("TA1", "TA2", "rho", "rho_binary", "spec1", "spec2"):
                    out[pre + k] = t[k]
                for k in t:
                    if k.startswith("rho_y"):
                        out[pre + k] = t[k]
                for k in t:
                    if k.startswith("rp") or k.startswith("rot"):
                        if k.endswith("/TA1") or k.endswith("/TA2") or k.endswith("spec1") or k.endswith("spec2") or k.endswith("rho_binary") or k.endswith("_x"):
                            continue        # keep the rotated fixtures small: rho + positions pin the sequence
                        out[pre + k] = t[k]
            ia += 1
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("%-24s tries=%3d accepted=%d  %6.1f KB" % (name, len(tries), ia, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for n in (sys.argv[1:] or SYSTEMS):
        run_system(n)
