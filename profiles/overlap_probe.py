"""Does running two event batches concurrently (two contexts = two streams, two host threads) raise
throughput?  Probe for the dual-stream pipeline."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import supermc_b200 as smc
N, B = 32768, 2048
def mk(seed): return smc.Context(smc.capi.default_params(max_batch=B, randomseed=seed, **bench.WORKLOAD))
a, b = mk(1), mk(2)
oa = np.zeros(N, dtype=smc.capi.EVENT_OUT_DTYPE); ob = np.zeros(N, dtype=smc.capi.EVENT_OUT_DTYPE)
a.run_events(0, N, out=oa); b.run_events(0, N, out=ob)
t0 = time.perf_counter(); a.run_events(N, N, out=oa); a.run_events(2 * N, N, out=oa); t1 = time.perf_counter()
print("one context : %.0f events/s" % (2 * N / (t1 - t0)))
def work(c, o):
    c.run_events(N, N, out=o); c.run_events(2 * N, N, out=o)
ta = threading.Thread(target=work, args=(a, oa)); tb = threading.Thread(target=work, args=(b, ob))
t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join(); t1 = time.perf_counter()
print("two contexts: %.0f events/s" % (4 * N / (t1 - t0)))
