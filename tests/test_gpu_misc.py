"""-m gpu: centrality sort, dS/dy window re-draw loop, table-driven nuclei on the sampled path."""
import os
import numpy as np
import pytest

from helpers import Golden, ROOT, event_in_from, rel_err

pytestmark = pytest.mark.gpu


def test_centrality_sort_is_argsort_descending():
    import supermc_b200 as smc
    g = Golden("pbpb2760_glb")
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=512))
    ev = ctx.run_events(0, 5000)
    key = ev["total"]
    perm = ctx.centrality_sort(key)
    assert np.array_equal(np.sort(perm), np.arange(len(key)))
    assert np.all(np.diff(key[perm]) <= 0)
    # scripts/centrality_cut_h5.py:78-90: the 0-5 % bin
    nsample = int(len(key) * 5 / 100) - 1
    top = ev[perm[:nsample]]
    assert top["npart1"].min() + top["npart2"].min() > 150 and top["b"].max() < 6.5
    ctx.close()


def test_dsdy_window_redraws_until_inside():
    import supermc_b200 as smc
    g = Golden("pbpb2760_glb")
    lo, hi = 60.0, 120.0
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=128, cutdsdy=1, cutdsdy_lowerbound=lo, cutdsdy_upperbound=hi, randomseed=3))
    ev = ctx.run_events(0, 300)
    assert ((ev["dsdy"] >= lo) & (ev["dsdy"] <= hi)).all() and (ev["status"] == 0).all()
    free = smc.Context(g.smc_params(smc.capi, max_batch=128, randomseed=3)).run_events(0, 300)
    inside = (free["dsdy"] >= lo) & (free["dsdy"] <= hi)
    # events that were inside the window at their first accepted try are untouched; the others consumed more tries
    assert np.array_equal(ev["b"][inside], free["b"][inside]) and (ev["tries"][~inside] > free["tries"][~inside]).all()
    ctx.close()


@pytest.mark.parametrize("name", ["he3au200_glb", "oo200_glb", "auau200_nncorr"])
def test_table_nuclei_sampling_equals_oracle(name, oracle_lib):
    """sampler modes 2 (He3 / O configurations: rotated, not recentred, Nucleus.cpp:555-574) and 3 (NN-correlated Au:
    recentred, rotation re-drawn, recentred, Nucleus.cpp:623-666) on the same Philox streams as the oracle, whose
    restatement equals the reference on whole drand48 + rand() streams (tests/test_oracle_vs_ref.py)"""
    import supermc_b200 as smc
    import table_synth
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    A, B = int(g.par["aproj"]), int(g.par["atarg"])
    nn = int(g.par.get("include_nn_correlation", 0))
    tabs = {3: lambda: np.loadtxt(os.path.join(ROOT, "tests", "golden", "he3_configs_small.txt")), 16: table_synth.oxygen, 197: table_synth.au197}
    seed = 11
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=16, randomseed=seed))
    tab = [None, None]
    for side, a in ((0, A), (1, B)):
        if a in (3, 16) or nn:
            tab[side] = tabs[a](); ctx.load_config_table(side, tab[side])
    nuc = [port.nucleus(A, cfg.width), port.nucleus(B, cfg.width)]
    n = 6
    out = ctx.run_events(50, n)
    for e in range(n):
        ev = 50 + e
        for tr in range(400):
            b = np.sqrt(400.0 * port.StreamPhilox(seed, ev, tr, 0).u(0, 0, 0))
            rows = []
            for side in (0, 1):
                xc = b / 2 if side == 0 else -b / 2
                if tab[side] is None:
                    rows.append(port.populate(nuc[side], xc, 0.0, stream=port.StreamPhilox(seed, ev, tr, side))[0])
                else:
                    icfg = int(port.StreamPhilox(seed, ev, tr, side).u(8, 0, 0) * len(tab[side]))
                    rows.append(port.populate_table(nuc[side], tab[side][icfg], nn, nn, xc, 0.0, stream=port.StreamPhilox(seed, ev, tr, side)))
            r = port.collide(cfg, rows[0], rows[1], stream=port.StreamPhilox(seed, ev, tr, 0))
            npart = int((r["ncollA"] > 0).sum() + (r["ncollB"] > 0).sum())
            if r["ncoll"] > 0 and npart >= 2:
                break
        assert out[e]["tries"] == tr + 1 and out[e]["ncoll"] == r["ncoll"], (name, e)
        for side in (0, 1):
            got = ctx.nucleons(e, side)
            assert np.abs(got[:, [0, 1, 3, 4, 5, 6]] - rows[side][:, [0, 1, 3, 4, 5, 6]]).max() < 1e-11, (name, e, side)
        assert np.array_equal(ctx.collisions(e)[:, 4:6].astype(int), r["pairs"])
    ctx.close()


def test_deuteron_sampling_equals_oracle(oracle_lib):
    """d+Au: Hulthen inverse-CDF sampler (oracle restatement is bit-equal to the reference with drand48)"""
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden("auau200_glb_quarks"); cfg = g.oracle_cfg(port)
    seed = 5
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=16, randomseed=seed, aproj=2))
    nA = port.nucleus(2, cfg.width); nB = port.nucleus(197, cfg.width)
    out = ctx.run_events(10, 6)
    for e in range(6):
        ev = 10 + e
        for tr in range(300):
            b = np.sqrt(400.0 * port.StreamPhilox(seed, ev, tr, 0).u(0, 0, 0))
            p = port.populate_deuteron(nA, b / 2, 0.0, port.StreamPhilox(seed, ev, tr, 0))
            t, _ = port.populate(nB, -b / 2, 0.0, stream=port.StreamPhilox(seed, ev, tr, 1))
            r = port.collide(cfg, p, t, stream=port.StreamPhilox(seed, ev, tr, 0))
            npart = int((r["ncollA"] > 0).sum() + (r["ncollB"] > 0).sum())
            if r["ncoll"] > 0 and npart >= 2:
                break
        assert out[e]["tries"] == tr + 1 and out[e]["ncoll"] == r["ncoll"]
        assert np.abs(ctx.nucleons(e, 0)[:, [0, 1, 3, 4, 5, 6]] - p[:, [0, 1, 3, 4, 5, 6]]).max() < 1e-10
    ctx.close()


@pytest.mark.parametrize("name", ["pbpb2760_glb", "pbpb2760_sqrt_disk", "auau200_kln"])
def test_scan_mode_equals_profile_mode(name, oracle_lib):
    """Moments-only runs skip the zero fill and work on each event's bounding rectangle only (deposit tiles,
    combine, moments); the profile modes start from zeroed lattices.  Both must give the same rows and -- through
    the getter, which blanks what the device never wrote -- the same grids (identical zero pattern, cells to rounding), also when the grid pool
    still holds a different, larger event from the run before."""
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=64))      # the getters address the last batch: keep all tries in one
    if name == "auau200_kln":
        ctx.set_kln_table(g.z["kln_table"], float(g.z["kln_consts"][0]))
    evs = [event_in_from(t, port, cfg) for t in g.tries()]
    acc = [i for i, t in enumerate(g.tries()) if int(t["hdr"][4])]
    assert len(acc) >= 2
    full = ctx.run_from_positions(evs, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO | smc.RUN_THICKNESS)
    ref_rho = {i: ctx.grid(i, smc.GRID_RHO).copy() for i in acc}
    # dirty the pool: the accepted events in reverse order land in the slots of other (different-sized) events
    ctx.run_from_positions([evs[i] for i in acc[::-1]], smc.RUN_MOMENTS)
    scan = ctx.run_from_positions(evs, smc.RUN_MOMENTS)
    for i in acc:
        # the rectangle (hence the tiling, hence the order of the centre-of-mass partial sums) depends on which grids
        # were asked for: rows agree to rounding
        assert np.allclose(scan[i]["mom"], full[i]["mom"], rtol=1e-11, atol=1e-13), (name, i)
        assert abs(scan[i]["total"] / full[i]["total"] - 1) < 1e-13 and abs(scan[i]["dsdy"] / full[i]["dsdy"] - 1) < 1e-13
        assert scan[i]["nonzero_cells"] == full[i]["nonzero_cells"]
        got = ctx.grid(i, smc.GRID_RHO)
        # (the Gaussian recurrence restarts at tile boundaries, so cells agree to rounding, not bit for bit)
        assert np.array_equal(got == 0, ref_rho[i] == 0) and rel_err(got, ref_rho[i]).max() < 1e-12, (name, i)
    ctx.close()
