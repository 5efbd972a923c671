"""CPU (-m "not gpu"): the oracle restatement against the committed golden vectors that the UNMODIFIED
reference produced (tests/golden/make_golden.py).  This is what pins the oracle: every comparison below
is bit-exact (==) because the oracle follows the reference's expression order and is compiled without
FMA contraction, like the reference build."""
import numpy as np
import pytest

from helpers import Golden, SYSTEMS, src8_from, coll8_from


@pytest.mark.parametrize("name", SYSTEMS)
def test_constants(name, oracle_lib):
    g = Golden(name); cfg = g.oracle_cfg(oracle_lib)
    assert cfg.siginNN == g.consts[0] and cfg.width == g.consts[1] and cfg.sigma_gg == g.consts[2] and cfg.dsq == g.consts[3]
    assert (cfg.Maxx, cfg.Maxy) == (int(g.consts[4]), int(g.consts[5]))


@pytest.mark.parametrize("name", SYSTEMS)
def test_collisions_replay_drand48(name, oracle_lib):
    """the sweep + hit test, consuming a clone of the reference's drand48 stream (state snapshot in the
    fixture), must reproduce Ncoll, Npart1/2, per-nucleon collision counts and the (i,j) list exactly"""
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    for it, t in enumerate(g.tries()):
        hdr = t["hdr"]
        r = port.collide(cfg, t["proj"][:, :7], t["targ"][:, :7], stream=port.Stream48(state=hdr[5:8]))
        assert r["ncoll"] == int(hdr[1]), (name, it)
        assert int((r["ncollA"] > 0).sum()) == int(hdr[2]) and int((r["ncollB"] > 0).sum()) == int(hdr[3])
        assert np.array_equal(r["ncollA"], t["proj"][:, 7].astype(int)) and np.array_equal(r["ncollB"], t["targ"][:, 7].astype(int))
        if r["ncoll"] and int(hdr[4]):
            coll = t["coll"]
            assert np.array_equal(r["pairs"], coll[:, 4:6].astype(int))
            # midpoints (MCnucl.cpp:339-340)
            mx = (t["proj"][r["pairs"][:, 0], 0] + t["targ"][r["pairs"][:, 1], 0]) / 2.0
            assert np.array_equal(mx, coll[:, 0])
            # Uli-Glauber additional weight: integer division 1/ncoll (MCnucl.cpp:345-348)
            if int(g.par["which_mc_model"]) == 5 and int(g.par["sub_model"]) == 2:
                addw = (r["ncollA"][r["pairs"][:, 0]] == 1).astype(float) + (r["ncollB"][r["pairs"][:, 1]] == 1).astype(float)
                assert np.array_equal(addw, coll[:, 3])
            # target participants are listed in first-hit order (Nucleus::markWounded)
            order = np.argsort(np.where(r["firsthitB"] >= 0, r["firsthitB"], 1 << 30), kind="stable")[:int(hdr[3])]
            assert np.array_equal(order, t["targ_part"].astype(int))


@pytest.mark.parametrize("name", SYSTEMS)
def test_grids_and_moments(name, oracle_lib):
    port = oracle_lib
    g = Golden(name); cfg = g.oracle_cfg(port)
    ff = g.par["finalfactor"]
    for it, t in enumerate(g.tries()):
        if not int(t["hdr"][4]):
            continue
        p8 = src8_from(t["proj"], t["proj_part"]); t8 = src8_from(t["targ"], t["targ_part"]); c8 = coll8_from(t["coll"])
        rho, dndy = port.density(cfg, p8, t8, c8)
        assert dndy == t["dndy"][0], (name, it)
        if "rho" in t:
            assert np.array_equal(rho, t["rho"])
            assert np.array_equal(port.thickness(cfg, p8), t["TA1"]) and np.array_equal(port.thickness(cfg, t8), t["TA2"])
            assert np.array_equal(port.unit_gauss(cfg, c8), t["rho_binary"])
            sp = t["spectators"]; s8 = np.zeros((len(sp), 8)); s8[:, :2] = sp[:, :2]
            assert np.array_equal(port.unit_gauss(cfg, s8[sp[:, 2] > 0]), t["spec1"])
            assert np.array_equal(port.unit_gauss(cfg, s8[sp[:, 2] <= 0]), t["spec2"])
        boxes = np.concatenate([p8[:, 2:6], t8[:, 2:6], np.zeros((len(c8), 4))])     # getHotSpots order, quirk Q13
        e = port.eccentricities(cfg, rho * ff, boxes)
        row = g.ecc_rows[int(t["ecc_index"])]
        assert np.array_equal(e["mom"], row[:45].reshape(9, 5)), (name, it, np.abs(e["mom"] - row[:45].reshape(9, 5)).max())
        assert abs(e["total"] * cfg.dx * cfg.dy - row[47]) <= 4e-16 * row[47] and row[48] == t["hdr"][0]
        reg = t["region"]
        assert reg[0] == boxes[:, 0].min() and reg[1] == boxes[:, 1].max()


def test_rotation_sequence(oracle_lib):
    """GlueDensity::calcCMAngle + recenterGrid (the averaged-profile path): centre of mass and
    participant-plane angle of the reference, orders 2 and 3"""
    port = oracle_lib
    g = Golden("pbpb2760_rotate"); cfg = g.oracle_cfg(port)
    t = g.tr(0)
    o = port.cm_angle(cfg, t["rho"], 2)
    assert np.array_equal(o[:3], t["rp2/cm"])
    o3 = port.cm_angle(cfg, t["rot2/rho"], 3)
    assert np.array_equal(o3[:3], t["rp3/cm"])


def test_six_point_and_kln_integrand(oracle_lib):
    port = oracle_lib
    # f(x,y) quadratic reproduces itself (arsenal.cpp:33-54)
    f = lambda x, y: 1.5 * x * x - 0.7 * x * y + 0.3 * y * y + 2 * x - y + 4
    v = port.lib().smc_o_six_point(0.3, 0.6, f(0, 0), f(0, 1), f(0, 2), f(1, 0), f(1, 1), f(2, 0))
    assert abs(v - f(0.3, 0.6)) < 1e-13
    k = port.kln(200.0, 0.218)
    a = port.kln_integrand(k, 0.0, 1.2, 0.7, [0.2, 0.5, 0.25]); b = port.kln_integrand(k, 0.0, 0.7, 1.2, [0.2, 0.5, 0.75])
    assert a > 0 and abs(a - b) < 1e-12 * a      # TA<->TB symmetry at y=0 under phi -> phi + pi


def test_hulthen_inverse_cdf(oracle_lib):
    """deuteron separation: CDF(invCDF(u)) == u to the reference's own accuracy (1e-6 in r)"""
    L = oracle_lib.lib()
    for u in (0.01, 0.2, 0.5, 0.9, 0.999):
        r = L.smc_o_hulthen_inv_cdf(u)
        a, b = .228, 1.18; c = a * b * (a + b) / (a - b) ** 2
        cdf = 2 * c * (2 * np.exp(-r * (a + b)) / (a + b) - .5 * np.exp(-2 * a * r) / a - .5 * np.exp(-2 * b * r) / b + .5 / a + .5 / b - 2 / (a + b))
        assert abs(cdf - u) < 1e-6 and r > 0
