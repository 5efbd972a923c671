"""Small runs of every device path for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python profiles/sanitizer_run.py glb|kln|nbd|sqrt
(glb also runs the profile kinds: thickness, rho_binary, spectators; nbd also runs the operation-3 sequence)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supermc_b200 as smc
import bench
which = sys.argv[1]
if which == "glb":
    ctx = smc.Context(smc.capi.default_params(max_batch=48, randomseed=3, **bench.WORKLOAD))
    ev = ctx.run_events(0, 150); assert (ev["status"] == 0).all(); print("glb", ev["npart1"][:5], np.isfinite(ev["mom"]).all())
    ev = ctx.run_events(0, 40, smc.RUN_MOMENTS | smc.RUN_KEEP_RHO | smc.RUN_THICKNESS | smc.RUN_RHO_BINARY | smc.RUN_SPECTATORS); print(ctx.grid(3, smc.GRID_SPEC_A).sum())
elif which == "kln":
    os.environ["SMC_KLN_QUAD"] = "40,20,8"
    ctx = smc.Context(smc.capi.default_params(max_batch=48, randomseed=3, **bench.WORKLOAD_KLN)); ctx.build_kln_table()
    ev = ctx.run_events(0, 150); print("kln", ev["status"][:8], np.isfinite(ev["mom"]).all())
elif which == "nbd":
    p = dict(bench.WORKLOAD, cc_fluctuation_model=2, dx=0.4, dy=0.4)
    ctx = smc.Context(smc.capi.default_params(max_batch=48, randomseed=3, **p))
    ev = ctx.run_events(0, 100); print("nbd", ev["dsdy"][:6])
    ctx.avg_begin(2, 2); ctx.avg_run(0, 40); print("avg count", ctx.avg_count())
elif which == "sqrt":
    p = dict(bench.WORKLOAD, which_mc_model=7)
    ctx = smc.Context(smc.capi.default_params(max_batch=48, randomseed=3, **p))
    ev = ctx.run_events(0, 150); print("sqrt", ev["dsdy"][:6], np.isfinite(ev["mom"]).all())
elif which == "pos":          # the from-positions entry (golden positions and weights), the getters, the device sort
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import Golden, event_in_from
    g = Golden("pbpb2760_glb")
    ctx = smc.Context(g.smc_params(smc.capi, max_batch=64))
    evs = [event_in_from(t) for t in g.tries()]          # golden positions and weights; the pair uniforms come from Philox
    out = ctx.run_from_positions(evs, smc.RUN_MOMENTS | smc.RUN_THICKNESS | smc.RUN_RHO_BINARY | smc.RUN_SPECTATORS)
    print("pos", out["ncoll"][:6], ctx.collisions(3).shape, ctx.participants(3).shape, ctx.spectators(3).shape)
    ev = ctx.run_events(0, 300); print(ctx.centrality_sort(ev["total"])[:5])
ctx.close()
