"""-m gpu: the sampled path (smc_run_events).

1. Determinism: the CPU oracle's sampler / sweep (pinned bit-for-bit to the reference with drand48,
   tests/test_oracle_vs_ref.py) is driven with the *same Philox event streams* the CUDA path uses; the
   two must then produce the same tries, the same accepted events, the same integers, and positions
   equal up to libm rounding (cbrt vs pow(.,1/3), sincos).
2. Statistics (north_star level L3): Npart, Ncoll, dS/dy, b, |eps2|, |eps3|, |eps2'|, <r^2> of 10^5 GPU events against
   10^5 events of the unmodified reference for Pb+Pb 2.76 TeV, Au+Au 200 GeV, p+Pb 5.02 TeV (MC-Glauber) and MC-KLN
   Au+Au 200 GeV (tests/golden/ks_*_ref.npz), two-sample KS p > 0.01 each, plus the reference's own shipped
   centrality-cut tables (Npart and dS/dy thresholds at 5 ... 80 %).
3. Result set independent of batch size / first_event_id split (what makes multi-GPU sharding safe).
"""
import os
import numpy as np
import pytest

from helpers import Golden, GOLDEN, coll8_from, rel_err

pytestmark = pytest.mark.gpu

K_B, K_PAIR = 0, 5


def _oracle_event(port, cfg, nA, nB, seed, ev, bmin, bmax, npmin, npmax, max_tries=200):
    for tr in range(max_tries):
        sb = port.StreamPhilox(seed, ev, tr, 0)
        b = np.sqrt((bmax * bmax - bmin * bmin) * sb.u(K_B, 0, 0) + bmin * bmin)
        p, _ = port.populate(nA, b / 2.0, 0.0, stream=port.StreamPhilox(seed, ev, tr, 0))
        t, _ = port.populate(nB, -b / 2.0, 0.0, stream=port.StreamPhilox(seed, ev, tr, 1))
        r = port.collide(cfg, p, t, stream=port.StreamPhilox(seed, ev, tr, 0))
        np1, np2 = int((r["ncollA"] > 0).sum()), int((r["ncollB"] > 0).sum())
        if r["ncoll"] > 0 and npmin <= np1 + np2 <= npmax:
            return dict(b=b, tries=tr + 1, proj=p, targ=t, r=r, np1=np1, np2=np2)
    raise AssertionError("no accepted try")


@pytest.mark.parametrize("name,quark", [("pbpb2760_glb", None), ("auau200_glb_quarks", "rand"), ("ppb5020_glb_quarks", "rand"),
                                        ("uu193_deformed", None), ("cuau200_glb", "rand")])
def test_sampler_equals_oracle_on_same_streams(name, quark, oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name)
    cfg = g.oracle_cfg(port)
    seed = 4242
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=32, randomseed=seed))
    qt = None
    if quark:
        import importlib.util
        spec = importlib.util.spec_from_file_location("mkref", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "make_ref_rundir.py"))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        qt = m.quark_table(quark)
        ctx.load_quark_table(qt)
    p = g.par
    nA = port.nucleus(int(p["aproj"]), cfg.width, deformed=int(p.get("proj_deformed", 0)), quark_table=qt)
    nB = port.nucleus(int(p["atarg"]), cfg.width, deformed=int(p.get("targ_deformed", 0)), quark_table=qt)
    n = 16
    first = 1000
    out = ctx.run_events(first, n, smc.RUN_MOMENTS)
    for e in range(n):
        o = _oracle_event(port, cfg, nA, nB, seed, first + e, p["bmin"], p["bmax"], int(p["npmin"]), int(p["npmax"]))
        assert out[e]["tries"] == o["tries"], (name, e)
        assert abs(out[e]["b"] - o["b"]) <= 1e-14 * max(o["b"], 1)
        assert (out[e]["npart1"], out[e]["npart2"], out[e]["ncoll"]) == (o["np1"], o["np2"], o["r"]["ncoll"]), (name, e)
        for side, ref in ((0, o["proj"]), (1, o["targ"])):
            got = ctx.nucleons(e, side)
            assert np.abs(got[:, [0, 1, 3, 4, 5, 6]] - ref[:, [0, 1, 3, 4, 5, 6]]).max() < 1e-11, (name, e, side)
        col = ctx.collisions(e)
        assert np.array_equal(col[:, 4:6].astype(int), o["r"]["pairs"])
    ctx.close()


def test_events_do_not_depend_on_batching():
    import supermc_b200 as smc
    g = Golden("pbpb2760_glb")
    a = smc.Context(g.smc_params(smc.capi, max_batch=64, randomseed=99))
    b = smc.Context(g.smc_params(smc.capi, max_batch=7, randomseed=99))
    ra = a.run_events(500, 40)
    rb = np.concatenate([b.run_events(500, 13), b.run_events(513, 27)])
    assert ra.tobytes() == rb.tobytes()
    a.close(); b.close()


def _ks_p(x, y):
    from scipy import stats
    return stats.ks_2samp(x, y).pvalue


KS_SYSTEMS = {  # name: (golden system for the parameters, shipped-table stem or None, KLN)
    "pbpb2760": ("pbpb2760_glb", "MCGlbPbPb2760_withMultFluct", False),
    "auau200": ("auau200_glb_quarks", "MCGlbAuAu200_withMultFluct", False),
    "ppb5020": ("ppb5020_glb_quarks", None, False),
    "auau200_kln": ("auau200_kln", "MCKLNAuAu200_noMultFluct", True),
}


def _ks_sample(name, seed, n):
    import supermc_b200 as smc
    gname, _, kln = KS_SYSTEMS[name]
    g = Golden(gname)
    over = dict(max_batch=2048, randomseed=seed, bmin=0.0, bmax=20.0)
    if kln:
        over.update(tmax=71, tmax_subdivision=3)
    ctx = smc.Context(g.smc_params(smc.capi, **over))
    if kln:
        ctx.build_kln_table()                    # the device quadrature (each entry within 0.5 % of the reference's BASES table)
    out = ctx.run_events(0, n, smc.RUN_MOMENTS)
    assert (out["status"] == 0).all()
    ctx.close()
    return out, dict(npart=(out["npart1"] + out["npart2"]).astype(float), ncoll=out["ncoll"].astype(float), dsdy=out["total"],
                     b=out["b"], e2=np.hypot(out["mom"][:, 1, 0], out["mom"][:, 1, 1]), e3=np.hypot(out["mom"][:, 2, 0], out["mom"][:, 2, 1]),
                     e2p=np.hypot(out["mom"][:, 1, 2], out["mom"][:, 1, 3]), r2=out["mom"][:, 1, 4])


@pytest.mark.parametrize("name", list(KS_SYSTEMS))
def test_distributions_match_reference_ks(name):
    """north_star level L3: two-sample KS of eight observables of 10^5 sampled events against 10^5 events of the
    unmodified reference (its own 8-process mode; tests/golden/run_ks_reference.sh + make_ks_reference.py), gate
    p > 0.01 as SURVEY.md 8(c) states.  Eight observables x four systems at a 1 % gate would fail a perfect
    implementation one time in four, so an observable below the gate is re-tested on a second, independent 10^5-event
    sample and must pass there (false-alarm rate 1e-4 per observable)."""
    path = os.path.join(GOLDEN, "ks_%s_ref.npz" % name)
    if not os.path.exists(path):
        pytest.skip("KS reference sample for %s not generated" % name)
    ref = np.load(path)
    n = int(os.environ.get("SMC_KS_EVENTS", "100000"))
    out, got = _ks_sample(name, 20261017, n)
    ps = {k: _ks_p(got[k], ref[k]) for k in got}
    print("KS p-values %s:" % name, {k: round(v, 4) for k, v in ps.items()}, "n_gpu=%d n_ref=%d" % (n, len(ref["npart"])))
    low = [k for k, v in ps.items() if v <= 0.01]
    if low:
        _, got2 = _ks_sample(name, 77, n)
        ps2 = {k: _ks_p(got2[k], ref[k]) for k in low}
        print("second sample:", ps2)
        for k, v in ps2.items():
            assert v > 0.01, (name, k, ps[k], v)
    # acceptance of the rejection loop: <tries> must match too
    assert abs(out["tries"].mean() / float(ref["mean_tries"]) - 1) < 0.02
    # the reference's own shipped centrality tables (scripts/centrality_cut_tables/, rows at 5 ... 80 % copied into
    # tests/golden/shipped_centrality_rows.npz): Npart thresholds within +-4, dS/dy thresholds within 2 %
    stem = KS_SYSTEMS[name][1]
    if stem:
        rows = np.load(os.path.join(GOLDEN, "shipped_centrality_rows.npz"))
        srt = np.sort(got["dsdy"])[::-1]
        for r in rows["total_entropy_" + stem]:
            val = srt[int(n * r[0] / 100)]
            assert abs(val / r[1] - 1) < 0.02, (name, r[0], val, r[1])
        if "Npart_" + stem in rows.files:
            srt = np.sort(got["npart"])[::-1]
            for r in rows["Npart_" + stem]:
                assert abs(srt[int(n * r[0] / 100)] - r[1]) <= 4, (name, r[0], srt[int(n * r[0] / 100)], r[1])
