"""Build container only (skipped where /root/reference or oracle/_ref is absent, e.g. on the GPU box):
the oracle restatement against the UNMODIFIED reference run live -- whole multi-try drand48 streams, so a
single uniform consumed out of order would break every later event."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
pytestmark = [pytest.mark.ref, pytest.mark.skipif(not (os.path.isdir("/root/reference/src") and os.path.exists(os.path.join(REF, "ref_dump"))
                                                       and os.path.isdir(os.path.join(REF, "run_rand"))),
                                                  reason="needs /root/reference and the oracle/_ref build")]


def _dump(kind, n, args, tmp):
    from oracle import refio
    run = os.path.join(REF, "run_" + kind)
    for f in os.listdir(os.path.join(run, "data")):
        os.remove(os.path.join(run, "data", f))
    out = str(tmp / "d.bin")
    subprocess.check_call([os.path.join(REF, "ref_dump"), out, str(n), "maxx=13", "maxy=13", "finalFactor=1", "dump_tries=1"] + args,
                          cwd=run, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return refio.group_tries(refio.read_records(out))


@pytest.mark.parametrize("kind,A,B,ecm,extra", [("zero", 208, 208, 2760, []), ("rand", 197, 197, 200, []), ("rand", 1, 208, 5020, []),
                                                ("rand", 238, 238, 193, ["proj_deformed=1", "targ_deformed=1"]), ("rand", 2, 197, 200, [])])
def test_sampler_and_sweep_follow_the_reference_stream(kind, A, B, ecm, extra, tmp_path, oracle_lib):
    port = oracle_lib
    seed = 7
    glob, tries = _dump(kind, 4, ["which_mc_model=5", "sub_model=1", "Aproj=%d" % A, "Atarg=%d" % B, "ecm=%g" % ecm, "randomSeed=%d" % seed] + extra, tmp_path)
    cfg = port.make_cfg(ecm=float(ecm))
    qt = np.loadtxt(os.path.join(REF, "run_" + kind, "tables", "QuarkPos.txt"))
    defo = 1 if extra else 0
    nA = port.nucleus(A, cfg.width, quark_table=qt, deformed=defo); nB = port.nucleus(B, cfg.width, quark_table=qt, deformed=defo)
    st = port.Stream48(seed=seed)
    for t in tries:
        b = np.sqrt(400.0 * st.next())
        assert b == t["hdr"][0]
        p = port.populate_deuteron(nA, b / 2.0, 0.0, st) if A == 2 else port.populate(nA, b / 2.0, 0.0, stream=st)[0]
        q, _ = port.populate(nB, -b / 2.0, 0.0, stream=st)
        assert np.array_equal(p, t["proj"][:, :7]) and np.array_equal(q, t["targ"][:, :7])
        r = port.collide(cfg, p, q, stream=st)
        assert r["ncoll"] == int(t["hdr"][1])
        assert np.array_equal(r["ncollA"], t["proj"][:, 7].astype(int)) and np.array_equal(r["ncollB"], t["targ"][:, 7].astype(int))


@pytest.mark.parametrize("A,nncorr", [(16, 0), (197, 1)])
def test_table_nuclei_follow_the_reference_stream(A, nncorr, tmp_path, oracle_lib):
    """O+O (configuration picked with libc rand(), rotated, NOT recentred; Nucleus.cpp:462-478,555-574) and
    NN-correlated Au+Au (recentred, rotation re-drawn, recentred again; Nucleus.cpp:481-522,623-666,236-272) on synthetic
    tables in the reference's file formats (tests/table_synth.py): whole drand48 + rand() streams"""
    import ctypes
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import table_synth
    port = oracle_lib
    seed = 9
    par = {"Aproj": A, "Atarg": A, "include_NN_correlation": nncorr}
    table_synth.install(os.path.join(REF, "run_rand", "tables"), par)
    glob, tries = _dump("rand", 6, ["which_mc_model=5", "sub_model=1", "Aproj=%d" % A, "Atarg=%d" % A, "ecm=200", "randomSeed=%d" % seed,
                                    "include_NN_correlation=%d" % nncorr], tmp_path)
    cfg = port.make_cfg(ecm=200.0)
    qt = np.loadtxt(os.path.join(REF, "run_rand", "tables", "QuarkPos.txt"))
    tab = table_synth.oxygen() if A == 16 else table_synth.au197()
    nA = port.nucleus(A, cfg.width, quark_table=qt); nB = port.nucleus(A, cfg.width, quark_table=qt)
    libc = ctypes.CDLL("libc.so.6"); libc.srand(seed)                     # src/main.cpp:32
    st = port.Stream48(seed=seed)
    for t in tries:
        b = np.sqrt(400.0 * st.next())
        assert b == t["hdr"][0]
        rows = []
        for n, xc in ((nA, b / 2.0), (nB, -b / 2.0)):
            # the reference draws the orientation first, then picks the configuration (Nucleus.cpp:193-194, 569 / 625)
            icfg = libc.rand() % len(tab)
            rows.append(port.populate_table(n, tab[icfg], nncorr, nncorr, xc, 0.0, stream=st))
        assert np.array_equal(rows[0], t["proj"][:, :7]) and np.array_equal(rows[1], t["targ"][:, :7])
        r = port.collide(cfg, rows[0], rows[1], stream=st)
        assert r["ncoll"] == int(t["hdr"][1])
