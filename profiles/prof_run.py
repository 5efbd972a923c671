"""Small driver used under ncu: `python profiles/prof_run.py [n_events] [batch]` runs the bench workload once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import supermc_b200 as smc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = smc.Context(smc.capi.default_params(max_batch=batch, randomseed=20261017, **bench.WORKLOAD))
ctx.run_events(0, n)
ctx.set_profiling(True)
ctx.run_events(n, n)
print(ctx.stage_ms(), ctx.last_run_ms)
