"""MC-KLN.

CPU: the oracle's deterministic quadrature of the restated integrand against the reference's own BASES
Monte-Carlo tables, built by the unmodified reference: tests/golden/auau200_kln.npz (Au+Au 200 GeV, lambda = 0.218,
70 x 70) and tests/golden/pbpb2760_kln.npz (Pb+Pb 2.76 TeV, lambda = 0.138, the full 211 x 211 table of the BASELINE
configuration).  The reference's stated MC accuracy is 0.1 %, observed differences are -0.02 .. -0.2 %, gate 0.5 %.
GPU: (1) the device table against the oracle quadrature on the same nodes (1e-10) and against EVERY entry of the
reference tables (0.5 %), (2) the 6-point table look-up + moments on the reference's golden KLN events (minimum-bias
and central), with the reference's own table installed."""
import numpy as np
import pytest

from helpers import Golden, KLN_SYSTEM, event_in_from, src8_from, rel_err

ENTRIES = {"auau200_kln": [(1, 1), (2, 17), (5, 9), (11, 3), (20, 20), (3, 30), (33, 12), (38, 38), (39, 1), (69, 69), (60, 7)],
           "pbpb2760_kln": [(1, 1), (2, 170), (5, 9), (110, 3), (20, 20), (30, 130), (133, 12), (138, 138), (209, 1), (210, 210), (77, 201)]}
KLN_SYSTEMS = [("auau200_kln", 70), ("pbpb2760_kln", 211)]


@pytest.mark.parametrize("name,size", KLN_SYSTEMS)
def test_oracle_quadrature_vs_reference_bases_table(name, size, oracle_lib):
    port = oracle_lib
    g = Golden(name)
    T = g.z["kln_table"]; dT, tmax = g.z["kln_consts"]
    assert T.shape == (size, size) and (T[0] == 0).all() and (T[:, 0] == 0).all()         # MCnucl.cpp:937-944
    k = port.kln(g.par["ecm"], g.par["lambda"])
    for i, j in ENTRIES[name]:
        v = port.kln_dndy(k, 0.0, dT * i, dT * j, 400, 200, 64)
        assert abs(v / T[i, j] - 1) < 5e-3, (i, j, v, T[i, j])
    # y = 0: dN/dy(TA,TB) = dN/dy(TB,TA)
    assert abs(port.kln_dndy(k, 0.0, dT * 4, dT * 9, 200, 100, 32) / port.kln_dndy(k, 0.0, dT * 9, dT * 4, 200, 100, 32) - 1) < 1e-12


@pytest.mark.parametrize("name", ["auau200_kln", "pbpb2760_kln", "pbpb2760_kln_central"])
def test_oracle_kln_density_on_reference_events(name, oracle_lib):
    """six-point look-up (MCnucl.cpp:654-687) with the reference's table on the reference's TA1/TA2: bit-exact rho"""
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    T = Golden(name.replace("_central", "")).z["kln_table"]; dT, tmax = g.z["kln_consts"]
    done = 0
    for t in g.tries():
        if "rho" not in t:
            continue
        rho, dndy = port.density_kln(cfg, t["TA1"], t["TA2"], T, float(dT))
        assert np.array_equal(rho, t["rho"]) and dndy == t["dndy"][0]
        done += 1
    assert done >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("name,size", KLN_SYSTEMS)
def test_gpu_table_equals_oracle_quadrature(name, size, oracle_lib, monkeypatch):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name)
    monkeypatch.setenv("SMC_KLN_QUAD", "200,100,32")
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=8))
    T = ctx.build_kln_table()
    dT = ctx.k.kln_dt
    assert T.shape == (size, size) and abs(dT - float(g.z["kln_consts"][0])) < 1e-15
    k = port.kln(g.par["ecm"], g.par["lambda"])
    for i, j in ENTRIES[name]:
        v = port.kln_dndy(k, 0.0, dT * i, dT * j, 200, 100, 32)
        assert abs(T[i, j] / v - 1) < 1e-10, (i, j, T[i, j], v)
    assert np.abs(T - T.T).max() <= 1e-12 * T.max() and (T[0] == 0).all()
    ref = g.z["kln_table"]
    assert np.abs(T[1:, 1:] / ref[1:, 1:] - 1).max() < 5e-3          # EVERY entry of the reference's BASES table, to its MC error
    ctx.close()


@pytest.mark.gpu
def test_gpu_default_quadrature_vs_full_reference_table():
    """the production quadrature (400 x 200 x 64 nodes) against all 210^2 non-trivial entries of the reference's
    Pb+Pb 2.76 TeV lambda = 0.138 table"""
    import supermc_b200 as smc
    g = Golden("pbpb2760_kln")
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=8))
    T = ctx.build_kln_table(); ref = g.z["kln_table"]
    d = T[1:, 1:] / ref[1:, 1:] - 1
    assert np.abs(d).max() < 5e-3 and abs(d.mean()) < 2.5e-3, (np.abs(d).max(), d.mean())
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["auau200_kln", "pbpb2760_kln", "pbpb2760_kln_central"])
def test_gpu_kln_events_match_reference(name, oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=32))
    ctx.set_kln_table(Golden(name.replace("_central", "")).z["kln_table"], float(g.z["kln_consts"][0]))
    tries = g.tries()
    out = ctx.run_from_positions([event_in_from(t, port, cfg) for t in tries], smc.RUN_MOMENTS | smc.RUN_THICKNESS)
    for it, t in enumerate(tries):
        hdr = t["hdr"]
        assert (out[it]["ncoll"], out[it]["npart1"], out[it]["npart2"]) == (int(hdr[1]), int(hdr[2]), int(hdr[3]))
        if not int(hdr[4]):
            continue
        row = g.ecc_rows[int(t["ecc_index"])]
        assert np.abs(out[it]["mom"][:, :4] - row[:45].reshape(9, 5)[:, :4]).max() < 1e-9
        assert abs(out[it]["total"] / row[47] - 1) < 1e-10
        if "rho" in t:
            assert rel_err(ctx.grid(it, smc.GRID_RHO), t["rho"]).max() < 1e-9
            assert rel_err(ctx.grid(it, smc.GRID_TA1), t["TA1"]).max() < 1e-9
    ctx.close()


# ---- rcBK tabulated uGD (sub_model 100): table files are absent upstream, synthetic stand-ins (tests/rcbk_synth.py) ----
def _rcbk_tables(tmp_path):
    import rcbk_synth
    return rcbk_synth.write_files(str(tmp_path / "javier"), 100)


def test_oracle_rcbk_vs_reference_on_synthetic_tables(oracle_lib, tmp_path):
    """rcBKfunc::getFunc (rcBKfunc.h:65-121) and the kT integral with it, against the unmodified reference run on the
    same synthetic tables (tests/golden/rcbk_synth_ref.npz): uGD values to 1e-13, BASES table to its MC error"""
    import os
    from helpers import GOLDEN
    port = oracle_lib
    ref = np.load(os.path.join(GOLDEN, "rcbk_synth_ref.npz"))
    kt, na = _rcbk_tables(tmp_path)
    t = port.rcbk(100, kt, na)
    for qs2, x, kt2, alp, val in ref["ugd"]:
        got = port.rcbk_func(t, qs2, x, kt2, alp)
        assert abs(got - val) <= 1e-13 * max(abs(val), 1e-30), (qs2, x, kt2, got, val)
    T = ref["table"]; dT = float(ref["kln_consts"][0])
    k = port.kln(2760.0, 0.138, model=100)
    for i, j in [(1, 1), (3, 7), (10, 10), (21, 2), (21, 21)]:
        v = port.rcbk_dndy(k, t, 0.0, dT * i, dT * j, 200, 100, 32)
        assert abs(v / T[i, j] - 1) < 5e-3, (i, j, v, T[i, j])


@pytest.mark.gpu
def test_gpu_rcbk_table_equals_oracle(oracle_lib, tmp_path, monkeypatch):
    import supermc_b200 as smc
    port = oracle_lib
    kt, na = _rcbk_tables(tmp_path)
    t = port.rcbk(100, kt, na)
    monkeypatch.setenv("SMC_KLN_QUAD", "100,50,16")
    ctx = smc.Context(smc.capi.default_params(which_mc_model=1, sub_model=100, aproj=208, atarg=208, ecm=2760.0, tmax=8, tmax_subdivision=3,
                                              maxx=13.0, maxy=13.0, cc_fluctuation_model=0, max_batch=8, **{"lambda": 0.138}))
    ctx.load_rcbk_tables(kt, na)
    T = ctx.build_kln_table()
    dT = ctx.k.kln_dt
    k = port.kln(2760.0, 0.138, model=100)
    for i, j in [(1, 1), (3, 7), (10, 10), (21, 2), (21, 21)]:
        v = port.rcbk_dndy(k, t, 0.0, dT * i, dT * j, 100, 50, 16)
        assert abs(T[i, j] / v - 1) < 1e-9, (i, j, T[i, j], v)
    ctx.close()


# ---- several rapidity slices (ny = 3, ymax = 2): one table per slice at y = rapMin + (rapMax - rapMin) / ny * iy ----
def _ny3():
    g = Golden("auau200_kln_ny3")
    tables = np.stack([g.z["kln_table"], g.z["kln_table_y1"], g.z["kln_table_y2"]])
    ys = [-2.0 + 4.0 / 3 * iy for iy in range(3)]              # MCnucl.cpp:932 (divides by ny, not ny - 1)
    return g, tables, ys


def test_oracle_rapidity_slices_vs_reference(oracle_lib):
    port = oracle_lib
    g, tables, ys = _ny3()
    dT = float(g.z["kln_consts"][0])
    k = port.kln(g.par["ecm"], g.par["lambda"])
    for iy, y in enumerate(ys):
        for i, j in [(2, 3), (10, 31), (40, 8), (69, 69)]:
            v = port.kln_dndy(k, y, dT * i, dT * j, 400, 200, 64)
            assert abs(v / tables[iy][i, j] - 1) < 5e-3, (iy, i, j, v, tables[iy][i, j])
    assert np.abs(tables[0] / np.maximum(tables[2], 1e-300) - 1)[1:, 1:].max() > 0.05       # the slices really differ
    cfg = g.oracle_cfg(port)
    t = [t for t in g.tries() if "rho_y1" in t][0]
    for iy, key in enumerate(("rho", "rho_y1", "rho_y2")):
        rho, _ = port.density_kln(cfg, t["TA1"], t["TA2"], tables[iy], dT)
        assert np.array_equal(rho, t[key])
        boxes = np.concatenate([t["proj"][t["proj_part"].astype(int), 3:7], t["targ"][t["targ_part"].astype(int), 3:7], np.zeros((1, 4))])
        e = port.eccentricities(cfg, rho * g.par["finalfactor"], boxes)
        assert np.array_equal(e["mom"], g.ecc_rows[int(t["ecc_index"]) + iy][:45].reshape(9, 5))


@pytest.mark.gpu
def test_gpu_rapidity_slices_match_reference(oracle_lib, monkeypatch):
    import supermc_b200 as smc
    port = oracle_lib
    g, tables, ys = _ny3(); cfg = g.oracle_cfg(port)
    dT = float(g.z["kln_consts"][0])
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=32))
    assert ctx.p.ny == 3
    # (1) the device builds one table per slice, each within BASES' error of the reference's
    monkeypatch.setenv("SMC_KLN_QUAD", "200,100,32")
    T = ctx.build_kln_table()
    assert T.shape == tables.shape
    for iy in range(3):
        assert np.abs(T[iy][1:, 1:] / tables[iy][1:, 1:] - 1).max() < 5e-3, iy
    k = port.kln(g.par["ecm"], g.par["lambda"])
    assert abs(T[1][7, 22] / port.kln_dndy(k, ys[1], dT * 7, dT * 22, 200, 100, 32) - 1) < 1e-10
    # (2) with the reference's tables installed: one row per event and slice, equal to the reference's rows
    ctx.set_kln_table(tables, dT)
    tries = g.tries()
    out = ctx.run_from_positions([event_in_from(t, port, cfg) for t in tries], smc.RUN_MOMENTS | smc.RUN_THICKNESS)
    assert len(out) == 3 * len(tries)
    for it, t in enumerate(tries):
        if not int(t["hdr"][4]):
            continue
        for iy in range(3):
            o = out[3 * it + iy]; row = g.ecc_rows[int(t["ecc_index"]) + iy]
            assert (o["ncoll"], o["npart1"]) == (int(t["hdr"][1]), int(t["hdr"][2]))
            assert np.abs(o["mom"][:, :4] - row[:45].reshape(9, 5)[:, :4]).max() < 1e-9, (it, iy)
            assert abs(o["total"] / row[47] - 1) < 1e-10
        if "rho_y2" in t:      # the grid left on the device is the last slice's
            assert rel_err(ctx.grid(it, smc.GRID_RHO), t["rho_y2"]).max() < 1e-9
    # (3) sampled events: ny rows each, slices differ, event content independent of ny
    ev = ctx.run_events(0, 8)
    one = smc.Context(g.smc_params(smc.capi, max_batch=32, ny=1, ymax=2.0))
    one.set_kln_table(tables[0], dT)
    ev1 = one.run_events(0, 8)
    assert len(ev) == 24 and np.array_equal(ev["b"][0::3], ev1["b"]) and np.array_equal(ev["ncoll"][1::3], ev1["ncoll"])
    assert np.allclose(ev["total"][0::3], ev1["total"], rtol=1e-13) and (np.abs(ev["total"][2::3] / ev["total"][0::3] - 1) > 1e-3).all()
    ctx.close(); one.close()
