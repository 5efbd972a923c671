"""The reference's 3-D extension (scripts/generate_3d_profiles/profile_3d.cpp): participants of an event
as Gaussians in (eta_s, x, y).  Fixture tests/golden/profile3d.npz = the unmodified reference sources behind a harness that
also writes the rapidities / widths it drew (tests/golden/make_profile3d_golden.py).
CPU: oracle restatement == reference lattice, bit for bit.  GPU: smc_profile3d from the same rapidities / widths <= 1e-12
relative (exp rounding only: same summation order), its own draws follow the reference's distribution (KS)."""
import os
import numpy as np
import pytest

from helpers import GOLDEN, rel_err


def _fx():
    z = np.load(os.path.join(GOLDEN, "profile3d.npz"))
    nx, ny, neta, dx, dy, deta, ecm = z["grid"]
    return z, int(nx), int(ny), int(neta), dx, dy, deta, ecm


@pytest.mark.parametrize("flag", [0, 1, 3])
def test_oracle_profile3d_equals_reference(flag, oracle_lib):
    z, nx, ny, neta, dx, dy, deta, ecm = _fx()
    rho = oracle_lib.profile3d(nx, ny, neta, dx, dy, deta, z["src_%d" % flag])
    assert np.array_equal(rho, z["rho_%d" % flag]) and rho.max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("flag", [0, 1, 3])
def test_gpu_profile3d_matches_reference(flag):
    import supermc_b200 as smc
    z, nx, ny, neta, dx, dy, deta, ecm = _fx()
    src = z["src_%d" % flag]
    rho, eu, su = smc.capi.profile3d(src[:, 0], src[:, 1], src[:, 2].astype(int), nx, ny, neta, dx, dy, deta, ecm, random_flag=flag,
                                     eta=src[:, 3], sigma3=src[:, 4:7])
    ref = z["rho_%d" % flag]
    assert np.array_equal(rho == 0, ref == 0)
    assert rel_err(rho, ref).max() < 1e-12
    if flag == 0:      # nothing is drawn: the library's own set-up must give the reference's rapidities and widths
        rho0, eu, su = smc.capi.profile3d(src[:, 0], src[:, 1], src[:, 2].astype(int), nx, ny, neta, dx, dy, deta, ecm, random_flag=0)
        assert np.array_equal(eu, src[:, 3]) and np.abs(su - src[:, 4:7]).max() < 1e-15 and rel_err(rho0, ref).max() < 1e-12


@pytest.mark.gpu
def test_gpu_profile3d_draws_follow_the_reference_distribution():
    """the rapidity law of profile_3d::sample_eta_distribution_from_array: 20,000 draws per side against the tabulated density"""
    import supermc_b200 as smc
    from scipy import stats
    n = 20000
    ids = np.concatenate([np.ones(n, dtype=np.int32), 2 * np.ones(n, dtype=np.int32)])
    _, eta, sig = smc.capi.profile3d(np.zeros(2 * n), np.zeros(2 * n), ids, 5, 5, 5, 1.0, 1.0, 1.0, 19.6, random_flag=3, seed=7)
    yb = np.arctanh(np.sqrt(1. - 1. / (19.6 / 2.) ** 2))
    e = np.linspace(-yb, yb, 1000)
    f = np.where(np.abs(e) > 2.5, np.exp(-(np.abs(e) - 2.5) ** 2 / 0.5), 1.0)
    for side, dens in ((1, (1 - e / yb) * f), (2, (1 + e / yb) * f)):
        cdf = np.concatenate([[0], np.cumsum(0.5 * (dens[1:] + dens[:-1]))]); cdf /= cdf[-1]
        p = stats.kstest(eta[ids == side], lambda v: np.interp(v, e, cdf)).pvalue
        assert p > 0.01, (side, p)
    assert np.all(np.abs(sig[:, 2] - 0.5) <= 0.3 + 1e-12) and sig[:, 0].std() > 0.1 and (sig[:, 0] != sig[:, 1]).all()
