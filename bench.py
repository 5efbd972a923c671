#!/usr/bin/env python3
"""bench.py -- accepted events/sec of the superMC hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md 8(d) input 2): MC-Glauber Pb+Pb 2.76 TeV minimum-bias
eccentricity scan, 261x261 grid (maxx=maxy=13, dx=dy=0.1), operation 9, orders 1..9, Gamma weights.
A "step" is one pass of the hot path (sample -> collide -> deposit -> moments) over one batch of
`--events-per-step` events per GPU; events are sharded across ranks by global event id (weak scaling,
no data-path collective: only the per-event rows leave the GPU).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path through the C ABI
    python bench.py --impl reference ...                     # the reference's own CPU code on host cores

value  : events / device time of the step (CUDA events on the library's stream; inputs = event ids only)
e2e    : same metric through the public C-ABI call smc_run_events with HOST buffers, wall clock, the
         host->device copy of the event ids and the device->host read of every result row included
roofline: the dominant kernel against the FP64 pipe (this path is bound by the FP64
         CUDA-core pipe, not by HBM or tensor cores), algorithmic FLOPs per event from 8(d).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(which_mc_model=5, sub_model=1, aproj=208, atarg=208, ecm=2760.0, alpha=0.118, cc_fluctuation_model=6,
                cc_fluctuation_gamma_theta=0.75, maxx=13.0, maxy=13.0, dx=0.1, dy=0.1, bmin=0.0, bmax=20.0, npmin=2, npmax=500,
                shape_of_nucleons=2, collision_criterion=2, shape_of_entropy=2, finalfactor=1.0, ecc_from_order=1, ecc_to_order=9)
REF_ARGS = ["which_mc_model=5", "sub_model=1", "Aproj=208", "Atarg=208", "ecm=2760", "alpha=0.118", "cc_fluctuation_model=6",
            "cc_fluctuation_Gamma_theta=0.75", "maxx=13", "maxy=13", "dx=0.1", "dy=0.1", "operation=9", "finalFactor=1",
            "bmin=0", "bmax=20", "Npmin=2", "Npmax=500", "shape_of_nucleons=2", "collision_criterion=2", "shape_of_entropy=2",
            "ecc_from_order=1", "ecc_to_order=9", "use_sd=1", "use_ed=1"]
WORKLOAD_NAME = "MC-Glauber Pb+Pb 2.76 TeV min-bias eccentricity scan (operation 9), 261x261 grid, orders 1-9"
# second half of the BASELINE.json metric ("MC-Glauber & MC-KLN Pb+Pb"): same scan with the kT-factorised MC-KLN density
# (SURVEY.md 8(d) input 4 without the rcBK tables: KLN uGD, lambda = 0.138 as scripts/generateAvgprofile.py:211-222 sets it
# for 2.76 TeV; no multiplicity fluctuations).  `--workload kln`; the default bench line stays MC-Glauber.
WORKLOAD_KLN = dict(WORKLOAD, which_mc_model=1, sub_model=7, cc_fluctuation_model=0, **{"lambda": 0.138})
WORKLOAD_KLN_NAME = "MC-KLN Pb+Pb 2.76 TeV min-bias eccentricity scan (operation 9), 261x261 grid, orders 1-9, 211x211 dN/dy table built on the device"


# the other scan configurations of BASELINE.json / SURVEY.md 8(d) input 5 (parity-test cases; `--workload <name>` prints a full
# line for each, reference binary beside it): overrides of WORKLOAD and of REF_ARGS
def _ref_args(**kv):
    out = [x for x in REF_ARGS if x.split("=")[0] not in kv]
    return out + ["%s=%s" % (k, v) for k, v in kv.items()]
OTHER_WORKLOADS = {
    "ppb": ("MC-Glauber p+Pb 5.02 TeV min-bias eccentricity scan (operation 9), 261x261 grid, orders 1-9",
            dict(aproj=1, atarg=208, ecm=5020.0), _ref_args(Aproj=1, Atarg=208, ecm=5020)),
    "auau": ("MC-Glauber Au+Au 200 GeV min-bias eccentricity scan (operation 9), 261x261 grid, orders 1-9",
             dict(aproj=197, atarg=197, ecm=200.0, alpha=0.14, cc_fluctuation_gamma_theta=0.61),
             _ref_args(Aproj=197, Atarg=197, ecm=200, alpha=0.14, cc_fluctuation_Gamma_theta=0.61)),
    "sqrt": ("Pb+Pb 2.76 TeV sqrt(TA TB) scaling (which_mc_model=7) eccentricity scan (operation 9), 261x261 grid, orders 1-9",
             dict(which_mc_model=7), _ref_args(which_mc_model=7)),
    "nbd": ("MC-Glauber Pb+Pb 2.76 TeV with NBD multiplicity fluctuations (cc_fluctuation_model=1, k=0.75), eccentricity scan (operation 9), 261x261 grid",
            dict(cc_fluctuation_model=1, cc_fluctuation_k=0.75), _ref_args(cc_fluctuation_model=1, cc_fluctuation_k=0.75)),
}


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("SMC_BENCH_NO_CLOCKS"):      # diagnostic: is nvidia-smi's start-up visible in the timed region?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        if not self.lines:      # a timed region shorter than nvidia-smi's first sample: one query right after it
            try:
                self.lines = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                            capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops(ev, width, dx, kln=False, A=208, B=208):
    """F_dep = 2[Np n4^2 (rho_WN) + Nc n5^2 (rho_BC)] (operation 9 with MC-Glauber needs no
    TA1/TA2, so that term is not claimed), F_mom = 200 per non-zero cell.  MC-KLN: F_dep = 2 Np n5^2 (TA1 + TA2,
    participants only -- quirk Q6); the 6-point table lookup (~30 FLOP per cell) is counted with the moments."""
    import numpy as np
    n5, n4 = 2 * 5 * width / dx, 8 * width / dx
    npart = (ev["npart1"] + ev["npart2"]).astype(np.float64); nc = ev["ncoll"].astype(np.float64)
    f_dep = 2.0 * (npart * n5 * n5) if kln is True else 2.0 * (npart * n4 * n4) if kln == "sqrt" else 2.0 * (npart * n4 * n4 + nc * n5 * n5)
    f_mom = 200.0 * ev["nonzero_cells"].astype(np.float64)
    # collisions 6AB FLOP per try, hard-core scan 3A^2 per nucleus per try
    f_smp = ev["tries"].astype(np.float64) * (6.0 * A * B + 3.0 * A * A + 3.0 * B * B)
    return float(f_dep.sum()), float(f_mom.sum()), float(f_smp.sum())


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written); fallback: the HBM figure B200_PROFILING.md states for this pool"""
    peaks = {"hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))); peaks["source"] = "MEASURED_PEAKS.json"
    except Exception:
        pass
    return peaks


def ref_paths():
    exe = os.path.join(ROOT, "oracle", "_ref", "superMC_ref.e")
    run = os.path.join(ROOT, "oracle", "_ref", "run_zero")
    return exe, run


def run_reference_once(nproc, nev_each, seed0, args=None):
    """`nproc` concurrent copies of the reference binary (its own 8-process mode,
    CollectDataAccordingToSettings.py:110-115), separate working directories; returns wall seconds."""
    exe, run = ref_paths()
    dirs = []
    for i in range(nproc):
        d = tempfile.mkdtemp(prefix="smcref_")
        os.makedirs(os.path.join(d, "data"))
        for f in ("parameters.dat", "EOS", "tables"):
            os.symlink(os.path.join(run, f), os.path.join(d, f))
        dirs.append(d)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([exe] + (args or REF_ARGS) + ["nev=%d" % nev_each, "randomSeed=%d" % (seed0 + i)], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i, d in enumerate(dirs)]
    for p in procs:
        p.wait()
    dt = time.perf_counter() - t0
    n_ok = 0
    for d in dirs:
        try:
            with open(os.path.join(d, "data", "sn_ecc_eccp_10.dat")) as f:
                n_ok += sum(1 for _ in f)
        except OSError:
            pass
        subprocess.call(["rm", "-rf", d])
    return dt, n_ok


def cpu_baseline(cores, nev_each, args=None):
    """bounded sample of the same workload on the host cores; start-up (EOS + QuarkPos load) subtracted
    with a 1-event run as BASELINE.md section 3 prescribes."""
    exe, _ = ref_paths()
    if not os.path.exists(exe):
        return None
    t1, _ = run_reference_once(cores, 1, 77, args)
    t, n = run_reference_once(cores, nev_each, 177, args)
    loop = max(t - t1, 1e-3) if n > cores else t
    return dict(value=(n - cores) / loop if n > cores else n / t, unit="events/s", cores=cores, kind="reference",
                sample="%d concurrent process(es) x %d accepted events of the bench workload (sd+ed as in BASELINE.md), "
                       "oracle/_ref/superMC_ref.e = unmodified reference sources, g++ -O3, GSL shim; 1-event start-up run subtracted" % (cores, nev_each),
                wall_s=t, startup_s=t1)


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events-per-step", type=int, default=32768, help="events per GPU per step")
    ap.add_argument("--batch", type=int, default=2048, help="events resident per launch wave")
    ap.add_argument("--cpu-sample-events", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="glauber", choices=["glauber", "kln", "ebe", "scan-exe", "avg"] + sorted(OTHER_WORKLOADS),
                    help="glauber = BASELINE.json configs[1] (the headline line); kln = the same scan with the MC-KLN density; "
                         "ebe = BASELINE.json configs[0] through the drop-in executable, text output included; "
                         "scan-exe = configs[1] through the drop-in executable (superMC_b200.e operation=9, tables written); "
                         "avg = configs[2]: MC-KLN Au+Au 200 GeV averaged profiles (operation 3), one centrality window, all-reduce in the timed region")
    ap.add_argument("--ref-events-per-process", type=int, default=300,
                    help="reference arm: accepted events per host process and step (BASELINE.md section 3 asks >= 2000 for quoted numbers: about 90 s per step)")
    ap.add_argument("--exe-events", type=int, default=1000000, help="scan-exe: nev of one executable run")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_line(a)))
        return 0

    if a.workload == "scan-exe":
        if rank == 0:
            print(json.dumps(scan_exe_line(a)))
        return 0
    if a.workload == "avg":
        return avg_main(a, rank, world, local)
    if a.workload == "ebe":
        if rank == 0:
            print(json.dumps(ebe_line(a)))
        return 0
    # rank 0 prints ONE JSON line on stdout: libraries that write there (NCCL prints its version banner on stdout when
    # NCCL_DEBUG is set) are sent to stderr for the duration of the run
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np
    import torch
    import supermc_b200 as smc
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kln = a.workload == "kln"
    wl, wl_name = (WORKLOAD_KLN, WORKLOAD_KLN_NAME) if kln else (WORKLOAD, WORKLOAD_NAME)
    ref_args = None
    if a.workload in OTHER_WORKLOADS:
        wl_name, over, ref_args = OTHER_WORKLOADS[a.workload]
        wl = dict(WORKLOAD, **over)
    ctx = smc.Context(smc.capi.default_params(max_batch=a.batch, randomseed=20261017, **wl), device=local)
    table_s = None
    if kln:      # one-off start-up (MCnucl::makeTable, 12.5 min on one reference core), outside the timed region
        t_tab = time.perf_counter(); kln_table = ctx.build_kln_table(); table_s = time.perf_counter() - t_tab
    n = a.events_per_step
    out = np.zeros(n, dtype=smc.capi.EVENT_OUT_DTYPE)

    def barrier():
        torch.cuda.synchronize(local)
        if dist is not None:
            dist.barrier()

    step_id = [0]

    def step():
        first = (step_id[0] * world + rank) * n          # disjoint global event ids per (step, rank)
        ctx.run_events(first, n, smc.RUN_MOMENTS, out=out)
        step_id[0] += 1
        return ctx.last_run_ms

    for _ in range(a.warmup):
        step()
    clocks = ClockSampler(local); clocks.start()
    barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    dev_ms = 0.0
    f_dep = f_mom = f_smp = 0.0
    kept = []
    for _ in range(a.steps):
        dev_ms += step()
        kept.append({k: out[k].copy() for k in ("npart1", "npart2", "ncoll", "nonzero_cells", "tries")})   # roofline bookkeeping, evaluated after the timed region
    barrier()
    wall = time.perf_counter() - t0
    for cols in kept:
        fd, fm, fs = algorithmic_flops(cols, ctx.k.width, WORKLOAD["dx"], True if kln else ("sqrt" if wl["which_mc_model"] == 7 else False), wl["aproj"], wl["atarg"]); f_dep += fd; f_mom += fm; f_smp += fs
    launches = ctx.launches - l0
    ck = clocks.stop()
    # per-kernel durations: the same steps once more with CUDA events around every launch; this pass runs the
    # batches serially on one stream (the timed region above overlaps consecutive batches on two streams, where
    # an event-bracketed kernel time would include the other stream's work)
    ctx.set_profiling(True)
    step_id[0] -= a.steps
    for _ in range(a.steps):
        step()
    stage = ctx.stage_ms()
    ctx.set_profiling(False)
    tm = torch.tensor([wall, dev_ms * 1e-3], dtype=torch.float64, device="cuda:%d" % local)
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(tm[0]), float(tm[1])
    total_events = n * a.steps * world
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp64_peak = ctx.fp64_peak_tflops()
    flops = {"deposit": f_dep, "moments": f_mom, "sample_collide": f_smp}
    dom = max(("deposit", "moments", "sample_collide"), key=lambda k: stage[k])
    dom_flops = flops[dom]
    traffic = None
    try:      # DRAM bytes per event of each kernel from the committed `ncu --set full` capture of this command
        if kln:
            raise KeyError("the committed capture is of the MC-Glauber workload")
        tp = [os.path.join(ROOT, "profiles", f) for f in ("r02_dram_traffic.json", "r01_dram_traffic.json")]
        tj = json.load(open([f for f in tp if os.path.exists(f)][0]))
        traffic = tj[dom + "_kernel"]["dram_bytes_per_event"] * a.batch      # per launch (one launch = one batch)
    except Exception:
        pass
    achieved = dom_flops / (stage[dom] * 1e-3) / 1e12 if stage[dom] > 0 else 0.0
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    grid_bytes = 8.0 * ctx.G * n * a.steps      # the rho scratch grid is written once and read back per event
    line = {
        "metric": "events/sec", "value": total_events / dev_max, "unit": "events/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dev_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "events_per_step_per_gpu": n, "batch": a.batch, "sharding": "event-id ranges per rank, no collective",
                   "l2": "per-step working set (%s scratch %d MB + event records) exceeds the 126 MB L2" % ("TA1/TA2/rho" if kln else "rho", int((3 if kln else 1) * 8 * ctx.G * a.batch / 1e6))},
        "e2e": {"value": total_events / wall_max, "unit": "events/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": int(out.itemsize) * n},
        "gpu_launches": int(launches),
        "clocks": ck,
        "roofline": {"bound": "fp64", "kernel": dom + "_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                     "peak_source": "measured live: smc_measure_fp64_peak (8 independent DFMA chains/thread, all SMs); nominal B200 FP64 ~37-40 TFLOP/s",
                     "algorithmic_flops_per_event": {"deposit": f_dep / (n * a.steps), "moments": f_mom / (n * a.steps), "sample_collide": f_smp / (n * a.steps)},
                     "all_kernels_tflops": {k: (flops[k] / (stage[k] * 1e-3) / 1e12 if stage[k] > 0 else None) for k in flops},
                     "stage_ms_per_step": {k: v / a.steps for k, v in stage.items()},
                     "stage_note": "kernel durations from a second, serial pass over the same steps (CUDA events around every launch); the timed region overlaps consecutive batches on two streams",
                     "hbm": {"achieved_gbs": 2 * grid_bytes / (dev_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "note": "rho scratch write+read; not the binding resource"}},
    }
    if kln:
        line["config"]["kln_table_build_s"] = table_s
        line["roofline"]["hbm"] = {"achieved_gbs": (3 + 2 + 2) * grid_bytes / (dev_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                   "note": "upper bound (whole lattices): TA1/TA2/rho written, TA1/TA2 read by the lookup, rho read twice by the moments; the kernels only touch each event's bounding rectangle"}
    if not a.no_cpu_baseline and world == 1:       # reported at N=1 only
        if kln:
            # the reference rebuilds its dN/dy table at every start (12.5 min, MCnucl.cpp:911-960), so its binary cannot
            # be sampled within the bench budget: the per-event loop is timed with the oracle port on the same table
            cb = port_baseline(min(a.cpu_sample_events, 60), kln_table=kln_table, kln_dt=ctx.k.kln_dt)
        else:
            cb = cpu_baseline(1, a.cpu_sample_events, ref_args)
            if cb is None and ref_args is None:
                # the unmodified reference binary is not on this box: time the oracle port instead
                cb = port_baseline(a.cpu_sample_events)
        line["cpu_baseline"] = cb
    json_out.write(json.dumps(line) + "\n"); json_out.flush()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def host_info():
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return {"cpu_model": model, "logical_cores": os.cpu_count(), "compiler": "g++ -std=gnu++11 -fpermissive -O3 (oracle/ref_build/Makefile), GSL replaced by the header shim"}


def reference_modes(nev_each, repeats, modes):
    """the unmodified reference as (a) one process, (b) its own 8-process mode
    (CollectDataAccordingToSettings.py:110-115), (c) one process per logical core; event-loop rate = accepted events /
    (wall clock - wall clock of a 1-event run of the same process count); min and median over `repeats` runs."""
    out = {}
    ncore = max(1, min(os.cpu_count() or 1, 64))
    for name in modes:
        nproc = {"1_process": 1, "8_process": 8, "all_cores": ncore}[name]
        t1, _ = run_reference_once(nproc, 1, 5)
        rates = []
        for r in range(repeats):
            t, n = run_reference_once(nproc, nev_each, 1000 + 100 * r)
            rates.append(max(n - nproc, 0) / max(t - t1, 1e-3))
        rates.sort()
        out[name] = {"processes": nproc, "events_per_process": nev_each, "runs": repeats, "min": rates[0], "median": rates[len(rates) // 2],
                     "startup_s": t1}
    return out


def reference_line(a):
    exe, _ = ref_paths()
    if not os.path.exists(exe):
        return {"impl": "reference", "unavailable": "oracle/_ref/superMC_ref.e is not built (needs /root/reference at build time)"}
    ncore = max(1, min(os.cpu_count() or 1, 64))
    nev_each = a.ref_events_per_process
    t1, _ = run_reference_once(ncore, 1, 5)
    for w in range(min(a.warmup, 1)):
        run_reference_once(ncore, 2, 50 + w)
    rates, tot_t = [], 0.0
    for s in range(a.steps):
        t, n = run_reference_once(ncore, nev_each, 1000 + 100 * s)
        tot_t += max(t - t1, 1e-3); rates.append(max(n - ncore, 0) / max(t - t1, 1e-3))
    rates.sort()
    v = rates[len(rates) // 2]
    # the other two host configurations of BASELINE.md section 3, once each per call (median of 3 with --steps >= 3)
    other = reference_modes(nev_each, min(3, max(1, a.steps)), ["1_process", "8_process"])
    other["all_cores"] = {"processes": ncore, "events_per_process": nev_each, "runs": a.steps, "min": rates[0], "median": v, "startup_s": t1}
    return {"metric": "events/sec", "value": v, "unit": "events/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tot_t / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD_NAME, "events_per_step": ncore * nev_each, "host_processes": ncore},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": ncore, "kind": "reference",
                             "sample": "median of %d steps x %d processes x %d events of the bench workload (sd+ed), start-up run subtracted; "
                                       "BASELINE.md section 3 host configurations in `modes`" % (a.steps, ncore, nev_each),
                             "modes": other, **host_info()},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def clocks_of(fn, index=0):
    ck = ClockSampler(index); ck.start()
    r = fn()
    return r, ck.stop()


def scan_exe_line(a):
    """BASELINE.json configs[1] through the product's front door: supermc_b200/superMC_b200.e operation=9 with its twenty
    tables written (use_sd=1 use_ed=1, 130 formatted numbers per event and table set), whole-process wall clock; the event
    loop the program reports (`Time elapsed`, src/main.cpp:65-68) next to it; the reference binary beside it."""
    exe = os.path.join(ROOT, "supermc_b200", "superMC_b200.e")
    nev = a.exe_events
    args = [x for x in REF_ARGS] + ["gpu_batch=2048"]

    def ours(n):
        d = tempfile.mkdtemp(prefix="smcscan_"); os.makedirs(os.path.join(d, "data"))
        subprocess.check_call(["cp", os.path.join(ROOT, "supermc_b200", "parameters.dat"), d])
        os.sync()          # the previous run's 4 GB of dirty pages would throttle this run's writes
        t0 = time.perf_counter()
        so = subprocess.run([exe] + args + ["nev=%d" % n, "randomSeed=20261017"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
        dt = time.perf_counter() - t0
        loop = [float(l.split(":")[1]) for l in so.splitlines() if l.startswith("Time elapsed (in seconds)")]
        rows = sum(1 for _ in open(os.path.join(d, "data", "sn_ecc_eccp_10.dat")))
        nbytes = sum(os.path.getsize(os.path.join(d, "data", f)) for f in os.listdir(os.path.join(d, "data")))
        subprocess.call(["rm", "-rf", d])
        return dt, rows, (loop[0] if loop else None), nbytes
    for _ in range(max(a.warmup, 1)):
        t_start, _, _, _ = ours(64)

    def timed():
        tot, rows, loops, nb = 0.0, 0, [], 0
        for _ in range(a.steps):
            dt, r, lp, b = ours(nev); tot += dt; rows += r; loops.append(lp); nb += b
        return tot, rows, loops, nb
    (tot, rows, loops, nbytes), ck = clocks_of(timed)
    v = rows / tot
    loop_rate = nev * len(loops) / sum(loops) if all(loops) else None
    line = {"metric": "events/sec", "value": v, "unit": "events/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME + ", through superMC_b200.e with the sn_/en_ecc_eccp_1..10 tables written; whole-process wall clock "
                                   "(a 64-event run of the same binary, i.e. start-up, takes %.2f s)" % t_start,
                       "events_per_step": nev, "rows_written": rows, "text_bytes_per_step": nbytes // max(a.steps, 1),
                       "event_loop_events_per_s": loop_rate, "event_loop_s_reported_by_the_program": loops},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 12 * nev, "d2h_bytes_per_step": 440 * nev},
            "gpu_launches": None, "clocks": ck, "roofline": None}
    if not a.no_cpu_baseline:
        cb = cpu_baseline(1, min(a.cpu_sample_events, 150))
        if cb:
            line["cpu_baseline"] = cb
    return line


AVG_WORKLOAD = dict(which_mc_model=1, sub_model=7, aproj=197, atarg=197, ecm=200.0, cc_fluctuation_model=0, maxx=13.0, maxy=13.0, dx=0.1, dy=0.1,
                    shape_of_nucleons=2, collision_criterion=2, shape_of_entropy=2, finalfactor=1.0, ecc_from_order=1, ecc_to_order=9,
                    # 20-30 % of scripts/centrality_cut_tables/iebe_centralityCut_total_entropy_MCKLNAuAu200_sigmaNN_gauss_d0.9_noMultFluct.dat
                    # translated as scripts/generateAvgprofile.py:92-191 does (Npart and b windows of the two bounding rows)
                    npmin=127, npmax=234, bmin=5.7851427, bmax=8.6297897, **{"lambda": 0.218})
AVG_NAME = ("MC-KLN Au+Au 200 GeV averaged smooth profiles (operation 3), 261x261 grid, orders 2-3, sd+ed branches, rotated + reaction-plane, "
            "all seven averaged quantities, one centrality window (20-30 %)")
AVG_REF_ARGS = ["which_mc_model=1", "sub_model=7", "lambda=0.218", "Aproj=197", "Atarg=197", "ecm=200", "cc_fluctuation_model=0", "maxx=13", "maxy=13",
                "dx=0.1", "dy=0.1", "finalFactor=1", "operation=3", "Npmin=127", "Npmax=234", "bmin=5.7851427", "bmax=8.6297897", "average_from_order=2",
                "average_to_order=3", "use_sd=1", "use_ed=1", "use_block=1", "use_4col=0"]


def avg_main(a, rank, world, local):
    """BASELINE.json configs[2].  One step = one averaged-profile run of `--events-per-step` accepted events per GPU:
    smc_avg_begin -> smc_avg_run (per event and order: recentre / rotate / redeposit / accumulate, up to five density
    evaluations per order and branch) -> smc_avg_allreduce over all ranks (N > 1) -> one averaged lattice read back."""
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np
    import supermc_b200 as smc
    n = a.events_per_step if a.events_per_step != 32768 else 2048
    ctx = smc.Context(smc.capi.default_params(max_batch=min(a.batch, 1024), randomseed=20261017, **AVG_WORKLOAD), device=local)
    backend = "single"
    if world > 1:
        backend = ctx.comm_init(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("SMC_COMM_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 1)))
    t_tab = time.perf_counter(); ctx.build_kln_table(); table_s = time.perf_counter() - t_tab
    G = ctx.G
    step_id = [0]
    ar_ms = []

    def step():
        first = (step_id[0] * world + rank) * n
        ctx.avg_begin(2, 3, with_rp=True, branches=3)
        out = ctx.avg_run(first, n)
        if world > 1:
            ar_ms.append(ctx.avg_allreduce())
        g = ctx.avg_get(2, 0, 0, 0)              # the step's result leaves the device
        step_id[0] += 1
        return out, g
    for _ in range(a.warmup):
        step()
    del ar_ms[:]
    clocks = ClockSampler(local); clocks.start()
    if world > 1:      # after the sampler's start-up: a rank that enters the timed region early would time the others' skew
        ctx.comm_barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(a.steps):
        ts = time.perf_counter()
        out, g = step()
        step_ms.append(round(1e3 * (time.perf_counter() - ts), 2))
    if world > 1:
        ctx.comm_barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launches - l0
    ck = clocks.stop()
    if world > 1:      # max over ranks of the timed region
        walls, _ = ctx.comm_gather(np.array([wall]), world)
        wall = float(walls.max()) if rank == 0 else wall
    if rank != 0:
        ctx.close()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    total = n * a.steps * world
    # algorithmic HBM bytes per accepted event (8 G per lattice written, 16 G per accumulator update):
    # per order and branch two density evaluations (reaction plane, rotated) of 6 lattices (TA1 TA2 rho rho_binary spec_A
    # spec_B) written and read once by the accumulation, + 5 / 7 accumulator lattices updated; + the first density
    norders, nbranch = 2, 2
    dens_evals = norders * (1 + nbranch * 2)
    bytes_per_event = dens_evals * 6 * 8 * G + norders * nbranch * (6 * 8 * G + 7 * 8 * G) + norders * nbranch * (5 + 7) * 16 * G / max(min(a.batch, 1024), 1)
    achieved = total * bytes_per_event / wall / 1e9
    line = {"metric": "events/sec", "value": total / wall, "unit": "events/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": AVG_NAME, "events_per_step_per_gpu": n, "batch": min(a.batch, 1024), "kln_table_build_s": table_s, "step_ms": step_ms,
                       "collective": "one all-reduce of %d doubles per step (%s)" % (2 * 4 * 7 * G, backend),
                       "allreduce_ms": (sum(ar_ms) / len(ar_ms)) if ar_ms else None,
                       "l2": "each density evaluation of a batch writes %d MB of lattices (> 126 MB L2)" % int(6 * 8 * G * min(a.batch, 1024) / 1e6)},
            "e2e": {"value": total / wall, "unit": "events/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": int(out.itemsize) * n + 8 * G},
            "gpu_launches": int(launches), "clocks": ck,
            "roofline": {"bound": "hbm", "kernel": "deposit_kernel + accumulate_kernel (lattice writes and reads of the re-deposits)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None,
                         "note": "algorithmic bytes = whole lattices (8 G per lattice written or read, 16 G per accumulator update); the kernels only touch each "
                                 "event's bounding rectangle (~1/3 of the lattice), so the real traffic is lower; timing is wall clock of the C-ABI calls"}}
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = avg_cpu_baseline()
    json_out.write(json.dumps(line) + "\n"); json_out.flush()
    ctx.close()
    return 0


def avg_cpu_baseline(nev=40):
    """the unmodified reference on the same window, one process.  Its start-up is the dN/dy table build (tmax=24: a 70x70 table,
    enough for this centrality window: ~35 s of BASES Monte Carlo whose duration varies by seconds from run to run), so the
    event loop is timed from the moment the program writes data/dNdyTable.dat (the last thing MCnucl::makeTable does,
    src/MCnucl.cpp:959) to its exit, and the fixed cost of writing the 48 averaged files at the end is removed by the
    difference of two runs (4 and 4 + nev events)."""
    exe, run = ref_paths()
    if not os.path.exists(exe):
        return {"value": None, "unit": "events/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref/superMC_ref.e not built on this box"}

    nfiles = [0]

    def once(n):
        d = tempfile.mkdtemp(prefix="smcavg_"); os.makedirs(os.path.join(d, "data"))
        for f in ("parameters.dat", "EOS", "tables"):
            os.symlink(os.path.join(run, f), os.path.join(d, f))
        t0 = time.time()
        # tmax=40 covers this window's thickness (the reference exits with "increase the dimension of dndyTable" below ~28);
        # one subdivision instead of three keeps the table build short and does not change the work per event
        rc = subprocess.call([exe] + AVG_REF_ARGS + ["tmax=40", "tmax_subdivision=1", "nev=%d" % n, "randomSeed=3"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        nfiles[0] = len([f for f in os.listdir(os.path.join(d, "data")) if "Avg" in f or "order" in f])
        t_end = time.time()
        try:
            t_tab = os.path.getmtime(os.path.join(d, "data", "dNdyTable.dat"))
        except OSError:
            t_tab = t0
        subprocess.call(["rm", "-rf", d])
        return t_end - t_tab, t_tab - t0
    (l1, tab1), (l2, tab2) = once(4), once(4 + nev)
    if nfiles[0] == 0 or l2 <= l1:      # the reference wrote no averaged profile: do not report a rate
        return {"value": None, "unit": "events/s", "cores": 1, "kind": "reference", "sample": "the reference run of this window failed (event loops %.2f / %.2f s, %d output files)" % (l1, l2, nfiles[0])}
    return {"value": nev / max(l2 - l1, 1e-3), "unit": "events/s", "cores": 1, "kind": "reference",
            "sample": "%d accepted events of the same window (operation 3, all outputs; tmax=40, one table subdivision), oracle/_ref/superMC_ref.e: event loops of a %d- and a 4-event run "
                      "(%.2f s and %.2f s, timed from the write of data/dNdyTable.dat to exit) subtracted; the dN/dy table builds took %.1f and %.1f s"
                      % (nev, 4 + nev, l2, l1, tab2, tab1),
            **host_info()}


EBE_ARGS = ["which_mc_model=5", "sub_model=1", "Aproj=197", "Atarg=197", "ecm=200", "alpha=0.14", "cc_fluctuation_model=6",
            "cc_fluctuation_Gamma_theta=0.61", "maxx=13", "maxy=13", "dx=0.1", "dy=0.1", "finalFactor=1", "operation=1", "use_sd=1",
            "use_ed=0", "use_block=1", "use_4col=0", "randomSeed=9"]


def ebe_line(a):
    """BASELINE.json configs[0]: MC-Glauber Au+Au 200 GeV, event-by-event entropy density (one 261^2 text block per event)
    + eccentricities, through supermc_b200/superMC_b200.e -- wall clock of the whole process, start-up, CUDA
    initialisation and text formatting included; the same run with output_binary=1 (raw float64 lattices instead of text)
    shows what the text path costs; next to it the unmodified reference binary, one process and one process per host core
    (start-up runs subtracted)."""
    exe = os.path.join(ROOT, "supermc_b200", "superMC_b200.e")
    nev = 1000

    def ours(n, extra=()):
        d = tempfile.mkdtemp(prefix="smcebe_"); os.makedirs(os.path.join(d, "data"))
        subprocess.check_call(["cp", os.path.join(ROOT, "supermc_b200", "parameters.dat"), d])
        os.sync()          # the previous run's GBs of dirty pages would throttle this run's writes
        t0 = time.perf_counter()
        so = subprocess.run([exe] + EBE_ARGS + list(extra) + ["nev=%d" % n], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
        dt = time.perf_counter() - t0
        loop = [float(l.split(":")[1]) for l in so.splitlines() if l.startswith("Time elapsed (in seconds)")]      # main.cpp:65-68
        names = os.listdir(os.path.join(d, "data"))
        nfiles = len([f for f in names if f.startswith("sd_event_")])
        nbytes = sum(os.path.getsize(os.path.join(d, "data", f)) for f in names)
        subprocess.call(["rm", "-rf", d])
        return dt, nfiles, (loop[0] if loop else None), nbytes
    for _ in range(max(a.warmup, 1)):
        t_start, _, _, _ = ours(8)
    tot, files, loops, nbytes = 0.0, 0, [], 0
    for _ in range(a.steps):
        dt, nf, lp, nb = ours(nev); tot += dt; files += nf; loops.append(lp); nbytes += nb
    _, _, loop_bin, bytes_bin = ours(nev, ["output_binary=1"])
    v = nev * a.steps / tot
    loop_rate = nev * len(loops) / sum(loops) if all(loops) else None
    G8 = 8 * 261 * 261
    peaks = measured_peaks()
    line = {"metric": "events/sec", "value": v, "unit": "events/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "MC-Glauber Au+Au 200 GeV event-by-event profiles (operation 1), 261x261 grid, one text block file per event, "
                                   "whole-process wall clock incl. start-up (an 8-event run of the same binary takes %.2f s)" % t_start,
                       "events_per_step": nev, "files_written": files, "event_loop_s_reported_by_the_program": loops,
                       "event_loop_events_per_s": loop_rate, "text_bytes_per_step": nbytes // max(a.steps, 1),
                       "host_text_gb_per_s_in_the_loop": (nbytes / max(a.steps, 1)) / (sum(loops) / len(loops)) / 1e9 if all(loops) else None,
                       "binary_output": {"event_loop_events_per_s": nev / loop_bin if loop_bin else None, "bytes_per_step": bytes_bin,
                                         "note": "same run with output_binary=1: raw float64 lattices instead of %22.12g text"}},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": G8 * nev},
            "roofline": {"bound": "hbm", "kernel": "deposit_kernel (lattice writes) + device->host copy of every lattice", "achieved": (loop_rate or v) * 2 * G8 / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (loop_rate or v) * 2 * G8 / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "note": "algorithmic bytes = one write + one read of the 8 G-byte lattice per event; this mode is bound by the host side "
                                 "(text formatting and the copy of ~3 MB of text per event into the page cache), not by the GPU"}}
    rexe, run = ref_paths()
    if not a.no_cpu_baseline and os.path.exists(rexe):
        def ref(n, procs):
            ds = []
            for _ in range(procs):
                d = tempfile.mkdtemp(prefix="smcref_"); os.makedirs(os.path.join(d, "data")); ds.append(d)
                for f in ("parameters.dat", "EOS", "tables"):
                    os.symlink(os.path.join(run, f), os.path.join(d, f))
            t0 = time.perf_counter()
            ps = [subprocess.Popen([rexe] + EBE_ARGS + ["nev=%d" % n, "randomSeed=%d" % (7 + i)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i, d in enumerate(ds)]
            for q in ps:
                q.wait()
            dt = time.perf_counter() - t0
            for d in ds:
                subprocess.call(["rm", "-rf", d])
            return dt
        t1 = ref(1, 1); t = ref(41, 1)
        line["cpu_baseline"] = dict(value=40 / max(t - t1, 1e-3), unit="events/s", cores=1, kind="reference",
                                    sample="41 events of the same workload, oracle/_ref/superMC_ref.e, 1-event start-up run subtracted")
        P = os.cpu_count() or 1
        tp1 = ref(1, P); tp = ref(41, P)
        line["cpu_baseline_all_cores"] = dict(value=P * 40 / max(tp - tp1, 1e-3), unit="events/s", cores=P, kind="reference",
                                              sample="%d concurrent processes x 41 events (its own multi-process mode), start-up runs subtracted" % P)
    return line


def port_baseline(nev, kln_table=None, kln_dt=None):
    """oracle port (scalar C restatement) timed on one core: sample+collide+deposit+moments per event."""
    import numpy as np
    from oracle import port
    cfg = port.make_cfg(ecm=2760.0, alpha=0.118) if kln_table is None else port.make_cfg(ecm=2760.0, alpha=0.118, which_mc_model=1, sub_model=7, cc_fluct_model=0)
    nA = port.nucleus(208, cfg.width); nB = port.nucleus(208, cfg.width)
    st = port.Stream48(seed=5)
    t0 = time.perf_counter(); done = 0
    while done < nev:
        b = np.sqrt(400.0 * st.next())
        p, _ = port.populate(nA, b / 2, 0.0, stream=st); t, _ = port.populate(nB, -b / 2, 0.0, stream=st)
        r = port.collide(cfg, p, t, stream=st)
        if r["ncoll"] == 0:
            continue
        ia = np.nonzero(r["ncollA"])[0]; ib = np.nonzero(r["ncollB"])[0]
        p8 = np.zeros((len(ia), 8)); p8[:, :2] = p[ia, :2]; p8[:, 2:6] = p[ia, 3:7]; p8[:, 6] = 1
        t8 = np.zeros((len(ib), 8)); t8[:, :2] = t[ib, :2]; t8[:, 2:6] = t[ib, 3:7]; t8[:, 6] = 1
        c8 = np.zeros((r["ncoll"], 8)); c8[:, 0] = (p[r["pairs"][:, 0], 0] + t[r["pairs"][:, 1], 0]) / 2; c8[:, 1] = (p[r["pairs"][:, 0], 1] + t[r["pairs"][:, 1], 1]) / 2; c8[:, 6] = 1
        if kln_table is None:
            rho, _ = port.density(cfg, p8, t8, c8)
        else:
            rho, _ = port.density_kln(cfg, port.thickness(cfg, p8), port.thickness(cfg, t8), kln_table, kln_dt)
        boxes = np.concatenate([p8[:, 2:6], t8[:, 2:6], np.zeros((1, 4))])
        port.eccentricities(cfg, rho, boxes); port.eccentricities(cfg, rho, boxes)     # sd + ed
        done += 1
    dt = time.perf_counter() - t0
    return dict(value=nev / dt, unit="events/s", cores=1, kind="port",
                sample="%d accepted events, oracle/smc_oracle.c, one core%s" % (nev, "" if kln_table is None else "; dN/dy table taken as given (the reference spends 12.5 min building it at every start)"))


if __name__ == "__main__":
    sys.exit(main())
