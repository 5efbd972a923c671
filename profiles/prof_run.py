"""Small driver used under ncu: `python profiles/prof_run.py [n_events] [batch]` runs the bench workload once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import supermc_b200 as smc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
kln = len(sys.argv) > 3 and sys.argv[3] == "kln"
ctx = smc.Context(smc.capi.default_params(max_batch=batch, randomseed=20261017, **(bench.WORKLOAD_KLN if kln else bench.WORKLOAD)))
if kln:
    os.environ.setdefault("SMC_KLN_QUAD", "100,50,16")      # the table's accuracy does not matter for a kernel profile
    ctx.build_kln_table()
ctx.run_events(0, n)
ctx.set_profiling(True)
ctx.run_events(n, n)
print(ctx.stage_ms(), ctx.last_run_ms)
