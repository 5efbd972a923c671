"""-m gpu parity on lattices the golden fixtures do not cover, against the CPU oracle on the same inputs:

* a lattice much smaller than the nuclei (maxx = 5, maxy = 4): every deposit window is clipped by the lattice
  edges, the hot-spot region sticks out of the grid;
* anisotropic spacing dx != dy (the column-interval masks of the deposit kernel are built from 1/dy);
* a coarse lattice (dx = dy = 0.25: windows of ~20 cells) and a fine one (dx = dy = 0.05 on a small box:
  windows of ~100 cells, several 32-column stripes per source).

Also a full-size, size-independent property: events sampled on the device, read back through the getters and
fed to smc_run_from_positions must reproduce themselves (counts, collision lists, density grid, moments) bit
for bit -- the two entry points share the collision / deposit / moment kernels but not their input path.
"""
import numpy as np
import pytest

from helpers import Golden, event_in_from, src8_from, coll8_from, rel_err

pytestmark = pytest.mark.gpu

GRID_TOL = 1e-9
MOM_TOL = 1e-9

LATTICES = [dict(maxx=5.0, maxy=4.0, dx=0.1, dy=0.1),
            dict(maxx=13.0, maxy=13.0, dx=0.1, dy=0.13),
            dict(maxx=13.0, maxy=12.0, dx=0.17, dy=0.08),
            dict(maxx=13.0, maxy=13.0, dx=0.25, dy=0.25),
            dict(maxx=6.0, maxy=6.0, dx=0.05, dy=0.05)]


@pytest.mark.parametrize("lat", LATTICES, ids=lambda d: "maxx%g_maxy%g_dx%g_dy%g" % (d["maxx"], d["maxy"], d["dx"], d["dy"]))
@pytest.mark.parametrize("name", ["pbpb2760_glb", "auau200_disk_nucleons", "ppb5020_glb_quarks"])
def test_other_lattices_match_oracle(name, lat, oracle_lib):
    import supermc_b200 as smc
    port = oracle_lib
    g = Golden(name)
    for k, v in lat.items():
        g.par[k] = v
    cfg = g.oracle_cfg(port)
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=16))
    assert (ctx.k.maxx_cells, ctx.k.maxy_cells) == (cfg.Maxx, cfg.Maxy)
    tries = [t for t in g.tries() if int(t["hdr"][4])][:3]
    evs = [event_in_from(t, port, g.oracle_cfg(port)) for t in tries]
    flags = smc.RUN_MOMENTS | smc.RUN_THICKNESS | smc.RUN_RHO_BINARY | smc.RUN_SPECTATORS
    out = ctx.run_from_positions(evs, flags)
    ff = g.par["finalfactor"]
    for it, t in enumerate(tries):
        o = out[it]
        assert (o["ncoll"], o["npart1"], o["npart2"]) == (int(t["hdr"][1]), int(t["hdr"][2]), int(t["hdr"][3]))
        p8 = src8_from(t["proj"], t["proj_part"]); t8 = src8_from(t["targ"], t["targ_part"]); c8 = coll8_from(t["coll"])
        rho_ref, dndy = port.density(cfg, p8, t8, c8)
        sp = t["spectators"]; s8 = np.zeros((len(sp), 8)); s8[:, :2] = sp[:, :2]
        refs = ((smc.GRID_RHO, rho_ref), (smc.GRID_TA1, port.thickness(cfg, p8)), (smc.GRID_TA2, port.thickness(cfg, t8)),
                (smc.GRID_RHO_BINARY, port.unit_gauss(cfg, c8)), (smc.GRID_SPEC_A, port.unit_gauss(cfg, s8[sp[:, 2] > 0])),
                (smc.GRID_SPEC_B, port.unit_gauss(cfg, s8[sp[:, 2] <= 0])))
        for which, ref in refs:
            got = ctx.grid(it, which)
            assert got.shape == ref.shape
            assert np.array_equal(got == 0, ref == 0), (name, lat, it, which, "zero pattern")
            if ref.max() > 0:
                assert rel_err(got, ref).max() <= GRID_TOL, (name, lat, it, which, rel_err(got, ref).max())
        if rho_ref.sum() <= 0:
            continue
        boxes = np.concatenate([p8[:, 2:6], t8[:, 2:6], np.zeros((len(c8), 4))])     # getHotSpots order, quirk Q13
        e = port.eccentricities(cfg, rho_ref * ff, boxes)
        assert np.abs(o["mom"][:, :4] - e["mom"][:, :4]).max() <= MOM_TOL, (name, lat, it)
        assert (np.abs(o["mom"][:, 4] - e["mom"][:, 4]) / np.abs(e["mom"][:, 4])).max() <= MOM_TOL
        assert abs(o["total"] - e["total"] * cfg.dx * cfg.dy) <= 1e-11 * abs(o["total"])
        assert abs(o["dsdy"] - dndy * cfg.dx * cfg.dy) <= 1e-11 * o["dsdy"]
    ctx.close()


def test_sampled_events_reproduce_themselves_from_positions():
    """4096 sampled Pb+Pb events (two batches) -> nucleon rows, collision weights read back -> run_from_positions:
    identical Npart/Ncoll/collision list, density grid and moments."""
    import supermc_b200 as smc
    import bench
    n, batch = 4096, 2048
    par = dict(bench.WORKLOAD); par.update(collision_criterion=1)        # disk criterion: no per-pair uniforms to carry
    ctx = smc.Context(smc.capi.default_params(max_batch=batch, randomseed=77, **par))
    ev = ctx.run_events(0, n, smc.RUN_MOMENTS)
    assert (ev["status"] == 0).all()
    # the getters address the last batch: events [n - batch, n)
    pick = [0, 1, 7, 100, 1023, 2047]
    evs, keep = [], []
    for s in pick:
        a, b = ctx.nucleons(s, 0), ctx.nucleons(s, 1)          # rows x y ncoll xL xR yL yR weight
        def rows(x):
            r = x.copy(); r[:, 2] = 0.0; return r          # the z slot of the getter carries the collision count
        c = ctx.collisions(s)
        evs.append(dict(b=float(ev["b"][n - batch + s]), proj=rows(a), targ=rows(b), given_w=1, coll_weight=c[:, 2:4].copy()))
        keep.append((ctx.grid(s, smc.GRID_RHO).copy(), c.copy(), ev[n - batch + s].copy()))
    ctx.close()
    ctx2 = smc.Context(smc.capi.default_params(max_batch=16, randomseed=77, **par))
    out = ctx2.run_from_positions(evs, smc.RUN_MOMENTS)
    for k, (rho, coll, e0) in enumerate(keep):
        assert (out[k]["ncoll"], out[k]["npart1"], out[k]["npart2"]) == (e0["ncoll"], e0["npart1"], e0["npart2"])
        assert np.array_equal(ctx2.collisions(k), coll)
        assert np.array_equal(ctx2.grid(k, smc.GRID_RHO), rho)
        assert np.array_equal(out[k]["mom"], e0["mom"]) and out[k]["total"] == e0["total"]
    ctx2.close()
