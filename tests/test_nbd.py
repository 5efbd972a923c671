"""NBD multiplicity fluctuations (cc_fluctuation_model 1, 2; MCnucl.cpp:868-905).
  * the oracle's literal restatement of NBD::rand reproduces the reference's own sampler sample for sample
    (tests/golden/nbd_ref.npz: 10 (p, r) pairs x 20000 consecutive draws of one drand48 stream);
  * the closed-form law of that sampler (what the CUDA path inverts per cell) fits the reference's samples;
  * -m gpu: the fluctuated lattice equals the oracle's cell by cell, given the same Philox uniforms."""
import os
import numpy as np
import pytest

from helpers import GOLDEN, Golden, event_in_from

Z = np.load(os.path.join(GOLDEN, "nbd_ref.npz"))


def test_oracle_sampler_equals_reference_sample_for_sample(oracle_lib):
    port = oracle_lib
    st = port.Stream48(seed=int(Z["seed"]))
    for (p, r), ref in zip(Z["pairs"], Z["samples"]):
        got = np.array([port.nbd_rand(float(p), float(r), st) for _ in range(len(ref))])
        assert np.array_equal(got, ref), (p, r)


def test_closed_form_law_fits_reference_samples(oracle_lib):
    from scipy import stats
    port = oracle_lib
    for (p, r), ref in zip(Z["pairs"], Z["samples"]):
        k0, pr = port.nbd_law(float(p), float(r))
        assert ref.min() >= k0 and ref.max() < k0 + len(pr)            # the truncated support
        hist = np.bincount(ref - k0, minlength=len(pr)).astype(float)
        exp = pr * len(ref)
        keep = exp > 5
        if keep.sum() < 2:
            assert hist[np.argmax(pr)] == len(ref)                       # 6 std < 1: always the same value (0)
            continue
        o = np.append(hist[keep], hist[~keep].sum()); e = np.append(exp[keep], exp[~keep].sum())
        if e[-1] == 0:
            o, e = o[:-1], e[:-1]
        chi2 = ((o - e) ** 2 / e).sum()
        assert stats.chi2.sf(chi2, len(e) - 1) > 1e-3, (p, r, chi2, len(e))
        # and it is not the plain NBD: the mean is pulled by the truncation
        assert abs((pr * (k0 + np.arange(len(pr)))).sum() - ref.mean()) < 5 * ref.std() / np.sqrt(len(ref)) + 1e-12


def test_quantile_inverts_the_law(oracle_lib):
    port = oracle_lib
    rng = np.random.default_rng(5)
    for p, r in [(0.5714285714285714, 0.75), (0.3, 2.5), (0.8, 5.0), (0.02, 0.75)]:
        k0, pr = port.nbd_law(p, r)
        cdf = np.cumsum(pr)
        for u in rng.random(200):
            k = port.nbd_quantile(p, r, float(u))
            i = k - k0
            if abs(u - (cdf[i - 1] if i > 0 else 0.0)) < 1e-12 or abs(u - cdf[i]) < 1e-12:
                continue
            assert (cdf[i - 1] if i > 0 else 0.0) <= u < cdf[i] + 1e-15, (p, r, u, k)


@pytest.mark.gpu
@pytest.mark.parametrize("model", [1, 2])
def test_gpu_fluctuated_density_equals_oracle(model, oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden("pbpb2760_glb")
    g.par["dx"] = g.par["dy"] = 0.4        # nb = rho dx dy of order 1: on the 0.1 fm lattice the truncated sampler returns 0 almost everywhere
    cfg = g.oracle_cfg(port)
    tries = [t for t in g.tries() if int(t["hdr"][4]) and int(t["hdr"][2]) + int(t["hdr"][3]) > 150][:2]      # mid-central and central
    assert len(tries) == 2
    evs = [event_in_from(t, port, cfg) for t in tries]
    smooth_ctx = smc.Context(g.smc_params(smc.capi, max_batch=8, cc_fluctuation_model=0))
    smooth_ctx.run_from_positions(evs, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO | smc.RUN_THICKNESS)
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=8, cc_fluctuation_model=model, cc_fluctuation_k=0.75))
    refs = []
    for e in range(len(evs)):
        rho = smooth_ctx.grid(e, smc.GRID_RHO)
        u = port.cell_uniforms(int(g.par["randomseed"]), e, 0, rho.size).reshape(rho.shape)
        refs.append(port.fluctuate_density(cfg, model, 0.75, rho, u, smooth_ctx.grid(e, smc.GRID_TA1), smooth_ctx.grid(e, smc.GRID_TA2)))
    for flags in (smc.RUN_MOMENTS | smc.RUN_KEEP_RHO | smc.RUN_THICKNESS, smc.RUN_MOMENTS):
        out = ctx.run_from_positions(evs, flags)
        for e in range(len(evs)):
            ref = refs[e]
            got = ctx.grid(e, smc.GRID_RHO)
            n_ref, n_got = np.rint(ref * cfg.dx * cfg.dy), np.rint(got * cfg.dx * cfg.dy)
            assert (n_ref > 0).sum() > 50 and n_ref.max() >= 2                                  # the event really fluctuates
            # integer output: equal cell by cell (the pmf recurrence of the kernel and the lgamma form of the oracle differ
            # by ~1e-14 relative, so only a uniform within that distance of a CDF step could differ; none does here)
            bad = np.argwhere(n_ref != n_got)
            assert len(bad) == 0, (model, e, [(tuple(q), n_ref[tuple(q)], n_got[tuple(q)], u[tuple(q)]) for q in bad[:4]])
            assert abs(out[e]["dsdy"] - got.sum() * cfg.dx * cfg.dy) <= 1e-9 * max(out[e]["dsdy"], 1.0)
            # moments of the fluctuated lattice, by the oracle
            boxes = np.concatenate([tries[e]["proj"][tries[e]["proj_part"].astype(int), 3:7], tries[e]["targ"][tries[e]["targ_part"].astype(int), 3:7], np.zeros((1, 4))])
            mom = port.eccentricities(cfg, got, boxes)
            assert np.abs(out[e]["mom"][:, :4] - mom["mom"][:, :4]).max() < 1e-9
    ctx.close(); smooth_ctx.close()
