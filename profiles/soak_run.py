"""Soak run: N minimum-bias events of the bench workload (and of MC-KLN) through smc_run_events; every row must be finite, accepted
(status 0), with Npart >= 2 and eccentricities in [0, 1]; prints the tails of the distributions.  usage: soak_run.py [N] [glauber|kln]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import supermc_b200 as smc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
kln = len(sys.argv) > 2 and sys.argv[2] == "kln"
ctx = smc.Context(smc.capi.default_params(max_batch=2048, randomseed=777, **(bench.WORKLOAD_KLN if kln else bench.WORKLOAD)))
if kln:
    ctx.build_kln_table()
chunk = 1 << 19
bad = 0; nonfinite = 0; eccbad = 0; tries = 0; npmax = 0; ncmax = 0; t0 = time.time()
for first in range(0, n, chunk):
    m = min(chunk, n - first)
    ev = ctx.run_events(first, m)
    bad += int((ev["status"] != 0).sum())
    mom = ev["mom"]
    nonfinite += int((~np.isfinite(mom)).sum() + (~np.isfinite(ev["total"])).sum())
    ecc = np.hypot(mom[:, :, 0], mom[:, :, 1])
    eccbad += int(((ecc < 0) | (ecc > 1.0 + 1e-9)).sum())
    tries += int(ev["tries"].sum()); npmax = max(npmax, int((ev["npart1"] + ev["npart2"]).max())); ncmax = max(ncmax, int(ev["ncoll"].max()))
print("%s: %d events in %.1f s: status != 0: %d, non-finite values: %d, eccentricities outside [0,1]: %d, tries/event %.3f, max Npart %d, max Ncoll %d"
      % ("kln" if kln else "glauber", n, time.time() - t0, bad, nonfinite, eccbad, tries / n, npmax, ncmax))
ctx.close()
