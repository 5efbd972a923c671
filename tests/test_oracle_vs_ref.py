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
