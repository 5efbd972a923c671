"""-m gpu: averaged profiles (operation 3).  The recenter / rotate / redeposit sequence of
generate_profile_average, orders 2 and 3, on a golden event of the unmodified reference
(tests/golden/pbpb2760_rotate.npz): rho after each step, and the final positions / AABBs including the
stale-base-box behaviour."""
import numpy as np
import pytest

from helpers import Golden, event_in_from, rel_err

pytestmark = pytest.mark.gpu
AVG_SD = 0


def test_rotation_sequence_matches_reference(oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden("pbpb2760_rotate"); cfg = g.oracle_cfg(port)
    ff = g.par["finalfactor"]
    for it in range(g.ntries):
        t = g.tr(it)
        if "rot3/rho" not in t:
            continue
        ctx = smc.Context(g.smc_params(smc.capi, max_batch=8))
        ctx.avg_begin(2, 3, with_rp=True, branches=1)
        ev = event_in_from(t, port, cfg)
        ev["proj_extra"] = t["proj_x"]; ev["targ_extra"] = t["targ_x"]
        out = ctx.avg_run_from_positions([ev])
        assert out[0]["ncoll"] == int(t["hdr"][1]) and ctx.avg_count() == 1
        for order in (2, 3):
            for variant, key in ((1, "rp%d/rho" % order), (0, "rot%d/rho" % order)):
                got = ctx.avg_get(order, variant, AVG_SD) / ff
                ref = t[key]
                err = rel_err(got, ref)
                nbad = int((err > 1e-8).sum())
                assert nbad == 0, (order, variant, nbad, err.max())
        # positions and AABBs after the last rotation
        for side, key in ((0, "rot3/proj"), (1, "rot3/targ")):
            got = ctx.nucleons(0, side); ref = t[key]
            part = ref[:, 7] > 0
            # the reference moves participants only (its spectators are separate copies, MCnucl.cpp:1125-1149)
            assert np.abs(got[part][:, 0:2] - ref[part][:, 0:2]).max() < 1e-11
            assert np.abs(got[part][:, 3:7] - ref[part][:, 3:7]).max() < 1e-11
        col = ctx.collisions(0)
        assert np.abs(col[:, 0:2] - t["rot3/coll_xy"]).max() < 1e-11
        ctx.close()


AVG_FILES = [  # (file stem of src/MakeDensity.cpp:755-1228, variant 0 rotated / 1 reaction plane, quantity of SMC_AVG_*)
    ("%sAvg_order_%d", 0, 0), ("%sAvg_RP_order_%d", 1, 0),
    ("TATB_from%s_order_%d", 0, 1), ("TATB_from%s_RP_order_%d", 1, 1),
    ("rho_binary_from%s_order_%d", 0, 2), ("rho_binary_from%s_RP_order_%d", 1, 2),
    ("nuclear_thickness_TA_from%s_order_%d", 0, 3), ("nuclear_thickness_TA_from%s_RP_order_%d", 1, 3),
    ("nuclear_thickness_TB_from%s_order_%d", 0, 4), ("nuclear_thickness_TB_from%s_RP_order_%d", 1, 4),
    ("spectator_density_A_from%s_order_%d", 0, 5), ("spectator_density_B_from%s_order_%d", 0, 6)]


def _avg3_events(g, port, cfg):
    evs = []
    for it in range(g.ntries):
        t = g.tr(it)
        ev = event_in_from(t, port, cfg)
        ev["proj_extra"] = t["proj_x"]; ev["targ_extra"] = t["targ_x"]
        evs.append(ev)
    return evs


def test_all_averaged_quantities_match_the_reference_files(oracle_lib):
    """operation 3 end to end against the 48 files the unmodified reference wrote for the same five events
    (tests/golden/make_avg_golden.py): entropy and energy branch, rotated and reaction-plane variants, all seven
    averaged quantities, orders 2 and 3 -- 1e-8 per cell (the files carry 12 digits)"""
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden("pbpb2760_avg3"); cfg = g.oracle_cfg(port)
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=8))
    ctx.avg_begin(2, 3, with_rp=True, branches=3)
    out = ctx.avg_run_from_positions(_avg3_events(g, port, cfg))
    assert ctx.avg_count() == g.ntries and (out["status"] == 0).all()
    seen = 0
    for order in (2, 3):
        for branch, tag in ((0, "sd"), (1, "ed")):
            for stem, variant, quantity in AVG_FILES:
                name = stem % (tag if quantity == 0 else tag.capitalize(), order)
                ref = g.z["avg/" + name]
                got = ctx.avg_get(order, variant, quantity, branch)
                assert np.array_equal(got == 0, ref == 0), (name, "zero pattern")
                err = rel_err(got, ref).max()
                assert err <= 1e-8, (name, err)
                seen += 1
    assert seen == 48 == len(g.z["files"])
    ctx.close()


def test_average_over_sampled_events_is_smooth():
    """statistical sanity of the sampled path: the order-2 rotated average is elongated along x or y
    consistently, normalised to <dS/dy>, and the reaction-plane average is left-right symmetric"""
    import supermc_b200 as smc
    g = Golden("pbpb2760_glb")
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=256, bmin=6.0, bmax=8.0, randomseed=77))
    ctx.avg_begin(2, 2, with_rp=True, branches=3)
    n = 512
    out = ctx.avg_run(0, n)
    assert ctx.avg_count() == n
    sd = ctx.avg_get(2, 0, AVG_SD, 0); ed = ctx.avg_get(2, 0, AVG_SD, 1); rp = ctx.avg_get(2, 1, AVG_SD, 0)
    assert abs(sd.sum() * 0.01 / out["total"].mean() - 1) < 0.02
    assert abs(ed.sum() / sd.sum() - 1) < 0.02
    x = np.linspace(-13, 13, 261)
    x2 = (sd.sum(1) * x * x).sum() / sd.sum(); y2 = (sd.sum(0) * x * x).sum() / sd.sum()
    assert (y2 - x2) / (y2 + x2) > 0.1                  # rotated to the participant plane: eps_2 > 0 survives the average
    assert abs((rp.sum(1) * x).sum() / rp.sum()) < 0.05  # re-centred
    ctx.close()
