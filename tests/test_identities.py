"""Analytic identities the reference's formulas must satisfy: they pin the oracle and the CUDA
path independently of any golden vector.
  * sigma_in is recovered from sigma_gg: Int d^2b [1 - exp(-sigma_gg Tpp(b))] = sigma_in (GaussianNucleonsCal.cpp:130-163)
  * a deposited nucleon / collision integrates to its weight: the Gaussian is normalised (GaussianNucleonsCal.cpp:120-124),
    cut at 5 w (loses exp(-12.5) = 3.7e-6 of it)
"""
import numpy as np
import pytest


@pytest.mark.parametrize("ecm,shape", [(200.0, 2), (2760.0, 2), (5020.0, 2), (2760.0, 4), (5020.0, 3)])
def test_sigma_in_recovered_from_sigma_gg(ecm, shape, oracle_lib):
    port = oracle_lib
    sig = port.sigma_inel(ecm)
    w, sgg = port.gauss_params(shape, sig)
    # shapes 1, 2, 4: Newton iteration on the integral cut at 5 w (stops at |delta sigma_gg| < 1e-4); shape 3: closed form
    # of the untruncated integral (gamma_E + E1 + ln)
    bmax = 5.0 * w if shape != 3 else 14.0 * w
    b = (np.arange(400000) + 0.5) * (bmax / 400000)
    tpp = np.exp(-b * b / (4 * w * w)) / (4 * np.pi * w * w)
    got = np.sum(2 * np.pi * b * (1 - np.exp(-sgg * tpp))) * (b[1] - b[0]) * 10.0      # fm^2 -> mb
    assert abs(got / sig - 1) < 2e-5, (got, sig)


def test_oracle_deposit_is_normalised(oracle_lib):
    port = oracle_lib
    cfg = port.make_cfg(ecm=2760.0)
    src = np.zeros((1, 8)); src[0, 0], src[0, 1], src[0, 6] = 0.2345, -1.0123, 1.0
    g = port.unit_gauss(cfg, src)
    tot = g.sum() * cfg.dx * cfg.dy
    assert 1 - 1e-5 < tot < 1.0 and abs(tot - (1 - np.exp(-12.5))) < 2e-6


@pytest.mark.gpu
def test_gpu_deposits_are_normalised():
    """one p+p collision at b = 0 (disk criterion: a certain hit): rho_binary, TA1 and TA2 each integrate to 1 - exp(-12.5),
    and the MC-Glauber rho to (1 - alpha) / 2 * 2 + alpha"""
    import supermc_b200 as smc
    alpha = 0.118
    ctx = smc.Context(smc.capi.default_params(which_mc_model=5, sub_model=1, aproj=1, atarg=1, ecm=2760.0, alpha=alpha, maxx=13.0, maxy=13.0,
                                              dx=0.1, dy=0.1, finalfactor=1.0, collision_criterion=1, cc_fluctuation_model=0, max_batch=4))
    w = ctx.k.width
    x0, y0 = 0.2345, -1.0123
    row = np.array([[x0, y0, 0.0, x0 - 4 * w, x0 + 4 * w, y0 - 4 * w, y0 + 4 * w, 1.0]])
    out = ctx.run_from_positions([dict(b=0.0, proj=row, targ=row, given_w=1)], smc.RUN_MOMENTS | smc.RUN_THICKNESS | smc.RUN_RHO_BINARY)
    assert (out[0]["ncoll"], out[0]["npart1"], out[0]["npart2"], out[0]["status"]) == (1, 1, 1, 0)
    cell = 0.1 * 0.1
    full = 1 - np.exp(-12.5)
    for which in (smc.GRID_RHO_BINARY, smc.GRID_TA1, smc.GRID_TA2):
        tot = ctx.grid(0, which).sum() * cell
        assert abs(tot - full) < 2e-6, (which, tot)
    # wounded nucleons are cut at the +-4 w box (quirk Q3): erf(4 / sqrt 2)^2 of the Gaussian; the collision term at 5 w
    from math import erf, sqrt
    expect = (1 - alpha) / 2 * 2 * erf(4 / sqrt(2)) ** 2 + alpha * full
    tot = ctx.grid(0, smc.GRID_RHO).sum() * cell
    assert abs(tot - expect) < 5e-5, (tot, expect)
    assert abs(out[0]["dsdy"] - tot) < 1e-12
    ctx.close()
