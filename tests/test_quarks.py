"""Valence-quark substructure: shape_of_entropy = 3 (Particle::getFluctuatedDensity, reference
src/Particle.cpp:149-163, src/Quark.cpp:14-22 incl. its units quirk d > 5*width on a squared distance) and
collision_criterion = 3 (GaussianNucleonsCal::testFluctuatedCollision, src/GaussianNucleonsCal.cpp:70-97).

CPU: the oracle restatement against fixtures of the unmodified reference (tests/golden/pbpb2760_quarks.npz,
auau200_quarkhit.npz: quark offsets, per-quark Gamma weights, drand48 snapshots, grids, 17-digit moment rows) -- bit-exact.
GPU: the CUDA path from the same positions."""
import numpy as np
import pytest

from helpers import Golden, event_in_from, src8_from, coll8_from, rel_err

QUARK_SYSTEMS = ["pbpb2760_quarks", "auau200_quarkhit"]
QW = 0.3          # quark_width of parameters.dat


def quark_inputs(t):
    """(qP, fP, qT, fT) of the participants, in participant order: offsets (x, y) x 3 and weights x 3"""
    out = []
    for side in ("proj", "targ"):
        idx = t[side + "_part"].astype(int)
        x = t[side + "_x"][idx]
        out += [np.stack([x[:, 4], x[:, 5], x[:, 7], x[:, 8], x[:, 10], x[:, 11]], axis=1), t[side + "_qf"][idx]]
    return out


def all_offsets(t, side):
    x = t[side + "_x"]
    return np.stack([x[:, 4], x[:, 5], x[:, 7], x[:, 8], x[:, 10], x[:, 11]], axis=1)


@pytest.mark.parametrize("name", QUARK_SYSTEMS)
def test_oracle_quark_paths_equal_the_reference(name, oracle_lib):
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    assert cfg.shape_of_entropy == 3
    ngrid = 0
    for it, t in enumerate(g.tries()):
        hdr = t["hdr"]
        st = port.Stream48(state=hdr[5:8])
        if int(g.par["collision_criterion"]) == 3:
            r = port.collide_quarks(cfg, t["proj"][:, :7], all_offsets(t, "proj"), t["targ"][:, :7], all_offsets(t, "targ"), QW, stream=st)
        else:
            r = port.collide(cfg, t["proj"][:, :7], t["targ"][:, :7], stream=st)
        assert r["ncoll"] == int(hdr[1]) and np.array_equal(r["ncollA"], t["proj"][:, 7].astype(int)), (name, it)
        if not int(hdr[4]):
            continue
        assert np.array_equal(r["pairs"], t["coll"][:, 4:6].astype(int))
        p8 = src8_from(t["proj"], t["proj_part"]); t8 = src8_from(t["targ"], t["targ_part"]); c8 = coll8_from(t["coll"])
        qP, fP, qT, fT = quark_inputs(t)
        rho, dndy = port.density_quarks(cfg, p8, qP, fP, t8, qT, fT, c8, QW)
        assert dndy == t["dndy"][0], (name, it)
        if "rho" in t:
            assert np.array_equal(rho, t["rho"]); ngrid += 1
        boxes = np.concatenate([p8[:, 2:6], t8[:, 2:6], np.zeros((len(c8), 4))])
        e = port.eccentricities(cfg, rho * g.par["finalfactor"], boxes)
        assert np.array_equal(e["mom"], g.ecc_rows[int(t["ecc_index"])][:45].reshape(9, 5)), (name, it)
    assert ngrid >= 1


def quark_event_in(t, port, cfg, crit3):
    """golden try -> smc_event_in with the quark state in the extras rows (offsets [4..12], weights [15..17])"""
    ev = event_in_from(t, port, cfg, with_uniforms=not crit3)
    if crit3 and len(t["coll"]) >= 0:
        st = port.Stream48(state=t["hdr"][5:8])
        r = port.collide_quarks(cfg, t["proj"][:, :7], all_offsets(t, "proj"), t["targ"][:, :7], all_offsets(t, "targ"), QW, stream=st, want_u=True)
        ev["pair_uniform"] = np.where(r["u"] < 0, 2.0, r["u"])
    for side in ("proj", "targ"):
        x = np.zeros((len(t[side]), 20)); x[:, :16] = t[side + "_x"]; x[:, 15:18] = t[side + "_qf"]
        ev[side + "_extra"] = x
    return ev


@pytest.mark.gpu
@pytest.mark.parametrize("name", QUARK_SYSTEMS)
def test_gpu_quark_paths_match_reference(name, oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    crit3 = int(g.par["collision_criterion"]) == 3
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=max(32, g.ntries)))
    tries = g.tries()
    out = ctx.run_from_positions([quark_event_in(t, port, cfg, crit3) for t in tries], smc.RUN_MOMENTS | smc.RUN_KEEP_RHO)
    for it, t in enumerate(tries):
        hdr = t["hdr"]; o = out[it]
        assert (o["ncoll"], o["npart1"], o["npart2"]) == (int(hdr[1]), int(hdr[2]), int(hdr[3])), (name, it)
        if not int(hdr[4]):
            continue
        assert np.array_equal(ctx.collisions(it)[:, 4:6].astype(int), t["coll"][:, 4:6].astype(int))
        p8 = src8_from(t["proj"], t["proj_part"]); t8 = src8_from(t["targ"], t["targ_part"]); c8 = coll8_from(t["coll"])
        qP, fP, qT, fT = quark_inputs(t)
        rho, dndy = port.density_quarks(cfg, p8, qP, fP, t8, qT, fT, c8, QW)
        got = ctx.grid(it, smc.GRID_RHO)
        assert np.array_equal(got == 0, rho == 0), (name, it, "zero pattern")
        assert rel_err(got, rho).max() <= 1e-9, (name, it, rel_err(got, rho).max())
        row = g.ecc_rows[int(t["ecc_index"])]
        assert np.abs(o["mom"][:, :4] - row[:45].reshape(9, 5)[:, :4]).max() <= 1e-9, (name, it)
        assert abs(o["total"] - row[47]) <= 1e-11 * row[47]
    ctx.close()


@pytest.mark.gpu
def test_gpu_quark_sampler_equals_oracle(oracle_lib):
    """sampled events with shape_of_entropy = 3 and the quark-overlap hit test: same tries, counts and pair lists as the
    oracle on the same Philox streams (the oracle's sampler is bit-equal to the reference on drand48 streams)"""
    import os
    import importlib.util
    import supermc_b200 as smc
    from helpers import GOLDEN
    port = oracle_lib
    g = Golden("auau200_quarkhit"); cfg = g.oracle_cfg(port)
    spec = importlib.util.spec_from_file_location("mkref", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "make_ref_rundir.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    qt = m.quark_table("rand")
    seed = 31
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=16, randomseed=seed))
    ctx.load_quark_table(qt)
    nA = port.nucleus(197, cfg.width, quark_table=qt); nB = port.nucleus(197, cfg.width, quark_table=qt)
    n = 6
    out = ctx.run_events(200, n, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO)
    for e in range(n):
        ev = 200 + e
        for tr in range(300):
            b = np.sqrt(400.0 * port.StreamPhilox(seed, ev, tr, 0).u(0, 0, 0))
            p, qp = port.populate_q(nA, b / 2, 0.0, port.StreamPhilox(seed, ev, tr, 0))
            t, qtq = port.populate_q(nB, -b / 2, 0.0, port.StreamPhilox(seed, ev, tr, 1))
            r = port.collide_quarks(cfg, p, qp, t, qtq, QW, stream=port.StreamPhilox(seed, ev, tr, 0))
            npart = int((r["ncollA"] > 0).sum() + (r["ncollB"] > 0).sum())
            if r["ncoll"] > 0 and npart >= 2:
                break
        assert out[e]["tries"] == tr + 1 and out[e]["ncoll"] == r["ncoll"], e
        assert np.array_equal(ctx.collisions(e)[:, 4:6].astype(int), r["pairs"])
        assert out[e]["total"] > 0
    ctx.close()
