"""-m gpu: edge cases and error behaviour of the C ABI, and size-independent properties at the BASELINE sizes
(Pb+Pb 2.76 TeV, 261 x 261): empty and ragged calls, capacity overflow, rejected options, call-order errors; linearity in
finalFactor, dS/dy == sum(rho) dx dy, bit-reproducibility, independence of the batch size at full size."""
import numpy as np
import pytest

from helpers import Golden

pytestmark = pytest.mark.gpu

PBPB = dict(which_mc_model=5, sub_model=1, aproj=208, atarg=208, ecm=2760.0, alpha=0.118, cc_fluctuation_model=6,
            cc_fluctuation_gamma_theta=0.75, maxx=13.0, maxy=13.0, dx=0.1, dy=0.1, finalfactor=1.0, randomseed=11)


def test_empty_and_ragged_calls():
    import supermc_b200 as smc
    ctx = smc.Context(smc.capi.default_params(max_batch=64, **PBPB))
    ev0 = ctx.run_events(0, 0)
    assert len(ev0) == 0
    a = ctx.run_events(0, 1)                       # a single event
    b = ctx.run_events(0, 64 * 3 + 5)              # three full batches and a ragged tail, through the slot pipeline
    assert a["status"][0] == 0 and (b["status"] == 0).all()
    assert a[0].tobytes() == b[0].tobytes()
    c = ctx.run_events(64 * 3, 5)                  # the tail alone: the same rows
    assert b[64 * 3:].tobytes() == c.tobytes()
    ctx.close()


def test_collision_list_overflow_is_reported_per_event():
    """an event whose Ncoll exceeds ncoll_cap gets status SMC_ERR_OVERFLOW (4) and the call still succeeds; the others are
    untouched (the reference has no such limit: its lists are std::vectors)"""
    import supermc_b200 as smc
    big = smc.Context(smc.capi.default_params(max_batch=128, **PBPB)).run_events(0, 256)
    cap = int(np.percentile(big["ncoll"], 60))
    ctx = smc.Context(smc.capi.default_params(max_batch=128, ncoll_cap=cap, **PBPB))
    ev = ctx.run_events(0, 256)
    over = big["ncoll"] > cap
    assert over.any() and (~over).any()
    assert (ev["status"][over] == 4).all() and (ev["status"][~over] == 0).all()
    assert ev[~over].tobytes() == big[~over].tobytes()
    assert np.array_equal(ev["ncoll"], big["ncoll"])          # the count itself is exact, only the list is capped
    ctx.close()


@pytest.mark.parametrize("bad", [dict(collision_criterion=4), dict(shape_of_entropy=7), dict(shape_of_nucleons=9), dict(dx=0.0),
                                 dict(which_mc_model=3), dict(aproj=0), dict(ny=0)])
def test_unsupported_options_are_rejected(bad):
    import supermc_b200 as smc
    with pytest.raises(smc.capi.SmcError):
        smc.Context(smc.capi.default_params(**dict(PBPB, **bad)))


def test_call_order_errors():
    import supermc_b200 as smc
    ctx = smc.Context(smc.capi.default_params(max_batch=32, **dict(PBPB, which_mc_model=1, sub_model=7)))
    with pytest.raises(smc.capi.SmcError):          # MC-KLN density without a table (MCnucl.cpp:636-640)
        ctx.run_events(0, 4)
    ctx.close()
    ctx = smc.Context(smc.capi.default_params(max_batch=32, **PBPB))
    with pytest.raises(smc.capi.SmcError):          # no batch yet
        ctx.participants(0)
    ctx.run_events(0, 8)
    with pytest.raises(smc.capi.SmcError):          # scan mode keeps no whole lattice
        ctx.grids(0, 8, smc.GRID_RHO)
    with pytest.raises(smc.capi.SmcError):          # outside the last batch
        ctx.collisions(8)
    with pytest.raises(smc.capi.SmcError):          # grids and lists are kept for one device batch
        ctx.run_events(0, 33, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO)
    with pytest.raises(smc.capi.SmcError):
        ctx.avg_run(0, 4)                           # smc_avg_begin first
    ctx.close()


def test_full_size_properties():
    """Pb+Pb 2.76 TeV at the BASELINE lattice, 2048 events: (1) bit-reproducible, (2) independent of the batch size,
    (3) every output density is linear in finalFactor while the eccentricities do not depend on it,
    (4) dS/dy = sum(rho) dx dy of the lattice the getters return, (5) |eps_n| <= 1 and <r^0> = 1."""
    import supermc_b200 as smc
    n = 2048
    a = smc.Context(smc.capi.default_params(max_batch=2048, **PBPB)).run_events(0, n)
    b = smc.Context(smc.capi.default_params(max_batch=2048, **PBPB)).run_events(0, n)
    assert a.tobytes() == b.tobytes()
    c = smc.Context(smc.capi.default_params(max_batch=192, **PBPB)).run_events(0, n)
    assert a.tobytes() == c.tobytes()
    f = smc.Context(smc.capi.default_params(max_batch=2048, **dict(PBPB, finalfactor=40.0))).run_events(0, n)
    assert np.allclose(f["total"], 40.0 * a["total"], rtol=1e-13) and np.allclose(f["dsdy"], a["dsdy"], rtol=1e-14)
    assert np.allclose(f["mom"], a["mom"], rtol=1e-9, atol=1e-12)
    ecc = np.hypot(a["mom"][:, :, 0], a["mom"][:, :, 1])
    assert (ecc <= 1.0 + 1e-12).all() and np.allclose(a["rn0"], 1.0)
    ctx = smc.Context(smc.capi.default_params(max_batch=64, **PBPB))
    ev = ctx.run_events(0, 48, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO)
    g = ctx.grids(0, 48, smc.GRID_RHO)
    assert np.allclose(g.reshape(48, -1).sum(axis=1) * 0.1 * 0.1, ev["dsdy"], rtol=1e-12)
    assert np.allclose(ev["mom"], a["mom"][:48], rtol=1e-11, atol=1e-13) and np.array_equal(ev["ncoll"], a["ncoll"][:48])   # profile mode == scan mode
    ctx.close()
