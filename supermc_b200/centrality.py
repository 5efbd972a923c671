"""Centrality tables and the per-centrality wrapper around the host driver.

Host-side mirror of three reference scripts (plain arrays instead of HDF5 -- h5py is not a dependency):

* ``collision_data``            scripts/collect_into_hdf5.py:23-38   (b, Npart, Ncoll, dS/dy, dE/dy from the 49-column rows)
* ``write_centrality_table``    scripts/centrality_cut_h5.py:36-110  (sort events by the cut quantity, one row per
                                centrality bound 0.1 ... 0.9, 1 ... 100 %; same text format as the shipped
                                scripts/centrality_cut_tables/*.dat).  The sort itself is the device radix sort behind
                                ``smc_centrality_sort``; this module only formats what the sorted order selects.
* ``translate_centrality_cut``  scripts/generateAvgprofile.py:92-175 (centrality window -> dS/dy or Npart window plus
                                the b / Npart ranges the rejection loop needs)
* ``python -m supermc_b200.centrality run ...``  scripts/generateAvgprofile.py:178-296 / generateEbeprofiles.py: one run of
                                the host driver (superMC_b200.e, or supermc_b200.launch under torchrun) per centrality bin.
"""
import argparse
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NUCLEUS_NAME = {208: "Pb", 197: "Au", 129: "Xe", 238: "U", 63: "Cu", 1: "p", 2: "d", 3: "He3"}     # generateAvgprofile.py:62-82
NUCLEUS_NUMBER = {v: k for k, v in NUCLEUS_NAME.items()}
CENTRALITY_BOUNDS = [x * 0.1 for x in range(0, 10)] + [float(y) for y in range(1, 101)]            # centrality_cut_h5.py:72-74


def collision_data(rows_sn, rows_en=None):
    """(nev, 49+) rows of sn_ecc_eccp_10.dat -> (nev, 5): b, Npart, Ncoll, dS/dy, dE/dy (collect_into_hdf5.py:27-33).
    The reference stores the array as float32; so does this (the tables inherit that rounding)."""
    rows_sn = np.atleast_2d(np.asarray(rows_sn, dtype=np.float64))
    rows_en = rows_sn if rows_en is None else np.atleast_2d(np.asarray(rows_en, dtype=np.float64))
    out = np.zeros((rows_sn.shape[0], 5))
    out[:, 0] = rows_sn[:, 48]; out[:, 1] = rows_sn[:, 45]; out[:, 2] = rows_sn[:, 46]; out[:, 3] = rows_sn[:, 47]
    out[:, 4] = rows_en[:, 47]
    return out.astype(np.float32)


def sort_key(coll, cut_type, alpha=0.118):
    """the quantity events are ranked by, largest first (b: smallest first) -- centrality_cut_h5.py:46-68"""
    coll = np.asarray(coll, dtype=np.float64)
    if cut_type == "b":
        return -coll[:, 0]
    if cut_type == "Npart":
        return coll[:, 1]
    if cut_type == "total_entropy":
        return coll[:, 3]
    if cut_type == "GlauberMixed":
        return (1. - alpha) / 2. * coll[:, 1] + alpha * coll[:, 2]
    raise ValueError("invalid cutType %r" % (cut_type,))


def device_order(ctx, coll, cut_type, alpha=0.118):
    """descending order of the cut quantity from the device sort (smc_centrality_sort)"""
    return ctx.centrality_sort(np.ascontiguousarray(sort_key(coll, cut_type, alpha), dtype=np.float64))


def centrality_table_text(coll, order, cut_type, alpha=0.118):
    """Text of iebe_centralityCut_<cut_type>_<name>.dat (centrality_cut_h5.py:44-107).  `order` is the sorted event
    order (device_order); rows: upper centrality bound, cut value, then the Npart / b ranges of the bin."""
    coll = np.asarray(coll, dtype=np.float64); order = np.asarray(order)
    nevent = coll.shape[0]
    s = []
    if cut_type == "b":
        s.append("#centrality b(fm)\n"); s.append("%6.4e %18.8e\n" % (0.0, 0.0))
    elif cut_type == "Npart":
        s.append("#centrality Npart b_min(fm) b_max(fm)\n"); s.append("%6.4e  %18.8e  %18.8e  %18.8e\n" % (0.0, 500, 0.0, 0.0))
    elif cut_type in ("total_entropy", "GlauberMixed"):
        s.append("#centrality dS/dy Npart_min Npart_max b_min(fm) b_max(fm)\n")
        s.append("%6.4e  %18.8e  %18.8e  %18.8e  %18.8e  %18.8e\n" % (0.0, 1000000, 500, 500, 0.0, 0.0))
    else:
        raise ValueError("invalid cutType %r" % (cut_type,))
    for icen in range(1, len(CENTRALITY_BOUNDS)):
        lower, upper = CENTRALITY_BOUNDS[icen - 1], CENTRALITY_BOUNDS[icen]
        nsample = int(nevent * (upper - lower) / 100) - 1
        noffset = int(nevent * lower / 100)
        sel = coll[order[noffset:noffset + nsample], :]
        if sel.shape[0] == 0:
            raise ValueError("too few events (%d) for a %g-%g %% bin" % (nevent, lower, upper))     # the reference's min() of an empty list
        npart_min, npart_max, b_min, b_max = sel[:, 1].min(), sel[:, 1].max(), sel[:, 0].min(), sel[:, 0].max()
        if cut_type == "total_entropy":
            s.append("%6.4e  %18.8e  %18.8e  %18.8e  %18.8e  %18.8e\n" % (upper, sel[:, 3].min(), npart_min, npart_max, b_min, b_max))
        elif cut_type == "GlauberMixed":
            s.append("%6.4e  %18.8e  %18.8e  %18.8e  %18.8e  %18.8e\n"
                     % (upper, ((1. - alpha) / 2. * sel[:, 1] + alpha * sel[:, 2]).min(), npart_min, npart_max, b_min, b_max))
        elif cut_type == "Npart":
            s.append("%6.4e  %18.8e  %18.8e  %18.8e\n" % (upper, npart_min, b_min, b_max))
        else:
            s.append("%6.4e  %18.8e\n" % (upper, b_max))
    return "".join(s)


def table_file_name(cut_type, which_mc_model, aproj, atarg, ecm, cc_fluctuation_model):
    """generateAvgprofile.py:101-129"""
    model = {5: "MCGlb", 1: "MCKLN", 7: "Trento"}[int(which_mc_model)]
    fluct = "withMultFluct" if cc_fluctuation_model != 0 else "noMultFluct"
    a, b = int(aproj), int(atarg)
    nuc = NUCLEUS_NAME[a] + NUCLEUS_NAME[b] if a == b else NUCLEUS_NAME[min(a, b)] + NUCLEUS_NAME[max(a, b)]
    return "iebe_centralityCut_%s_%s_sigmaNN_gauss_d0.9_%s.dat" % (cut_type, model + nuc + ("%g" % ecm), fluct)


def translate_centrality_cut(table, lower, upper, cut_type="total_entropy"):
    """Centrality window [lower, upper] % -> parameters of the run (generateAvgprofile.py:131-175): the cut value is
    interpolated linearly between the table rows around each bound; Npart and b ranges are the extremes of the rows
    the window touches.  `table` is the array of a centrality table (np.loadtxt of the file)."""
    t = np.atleast_2d(np.asarray(table, dtype=np.float64))
    lo_i = int(t[:, 0].searchsorted(lower + 1e-30)); up_i = int(t[:, 0].searchsorted(upper))

    def interp(i, x):
        return (t[i - 1, 1] - t[i, 1]) / (t[i - 1, 0] - t[i, 0]) * (x - t[i - 1, 0]) + t[i - 1, 1]
    cut_upper, cut_low = interp(lo_i, lower), interp(up_i, upper)
    rows = t[lo_i - 1:up_i + 1]
    p = {}
    if cut_type == "total_entropy":
        p["cutdSdy"] = 1
        p["Npmin"], p["Npmax"] = rows[:, 2].min(), rows[:, 3].max()
        p["bmin"], p["bmax"] = rows[:, 4].min(), rows[:, 5].max()
        p["cutdSdy_lowerBound"], p["cutdSdy_upperBound"] = cut_low, cut_upper
    elif cut_type == "Npart":
        p["cutdSdy"] = 0
        p["bmin"], p["bmax"] = rows[:, 2].min(), rows[:, 3].max()
        p["Npmin"], p["Npmax"] = cut_low, cut_upper
    else:
        raise ValueError("centrality cut type %r is not one the wrapper knows (total_entropy, Npart)" % (cut_type,))
    return p


def model_parameters(model, ecm, collsys):
    """generateAvgprofile.py:178-214"""
    p = {}
    if model == "MCGlb":
        p.update(which_mc_model=5, sub_model=1, cc_fluctuation_model=6)
    elif model == "MCKLN":
        p.update(which_mc_model=1, sub_model=7, cc_fluctuation_model=0)
    elif model == "Trento":
        p.update(which_mc_model=7, sub_model=1, cc_fluctuation_model=6)
    else:
        raise ValueError("invalid initial model type %r" % (model,))
    p["ecm"] = ecm
    if ecm == 2760:
        p.update({"alpha": 0.118} if model == "MCGlb" else {"lambda": 0.138} if model == "MCKLN" else {})
    if ecm <= 200:
        p.update({"alpha": 0.14} if model == "MCGlb" else {"lambda": 0.218} if model == "MCKLN" else {})
    p["Aproj"], p["Atarg"] = NUCLEUS_NUMBER[collsys[0]], NUCLEUS_NUMBER[collsys[1]]
    return p


def assignment_args(params):
    return ["%s=%s" % (k, ("%.17g" % v) if isinstance(v, float) else v) for k, v in params.items()]


def minbias_rows(ev):
    """smc_event_out rows -> the (nev, 5) array of collision_data: b, Npart, Ncoll, dS/dy, dE/dy (== dS/dy, quirk Q2)"""
    out = np.zeros((len(ev), 5))
    out[:, 0] = ev["b"]; out[:, 1] = ev["npart1"] + ev["npart2"]; out[:, 2] = ev["ncoll"]; out[:, 3] = ev["total"]; out[:, 4] = ev["total"]
    return out


def minbias_table(a):
    """One minimum-bias run of `nev` accepted events, sharded over the ranks of torchrun by global event id (so the event
    set does not depend on the GPU count), rows gathered on rank 0, sorted by the device radix sort, table written in the
    format of scripts/centrality_cut_tables/*.dat.  scripts/centrality_cut_h5.py + collect_into_hdf5.py without the files."""
    import supermc_b200 as smc
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    p = {k.lower(): v for k, v in model_parameters(a.model, a.ecm, a.collsys).items()}
    p.update(finalfactor=1.0, randomseed=a.seed, maxx=13.0, maxy=13.0)
    for kv in a.extra:
        k, v = kv.split("=", 1); k = k.lower()
        p[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
    if p.get("which_mc_model") == 5:
        p.setdefault("cc_fluctuation_gamma_theta", 0.75 if a.ecm > 1000 else 0.61)
    ctx = smc.Context(smc.capi.default_params(**p), device=local)
    if p["which_mc_model"] == 1:
        ctx.build_kln_table()
    if world > 1:
        ctx.comm_init(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("SMC_COMM_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 1)))
    lo, hi = a.nev * rank // world, a.nev * (rank + 1) // world
    ev = ctx.run_events(lo, hi - lo)
    ev = ev[ev["status"] == 0]
    rows = minbias_rows(ev)
    if world > 1:
        flat, _ = ctx.comm_gather(rows, world)
        rows = flat.reshape(-1, 5)
    rc = 0
    if rank == 0:
        coll = rows.astype(np.float32)           # collect_into_hdf5.py stores float32; the shipped tables inherit that rounding
        alpha = float(p.get("alpha", 0.118))
        text = centrality_table_text(coll, device_order(ctx, coll, a.cut, alpha), a.cut, alpha)
        name = table_file_name(a.cut, p["which_mc_model"], p["aproj"], p["atarg"], p["ecm"], p["cc_fluctuation_model"])
        os.makedirs(a.out, exist_ok=True)
        open(os.path.join(a.out, name), "w").write(text)
        print(os.path.join(a.out, name), "%d events" % len(coll))
    if world > 1:
        ctx.comm_barrier()
    ctx.close()
    return rc


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m supermc_b200.centrality")
    sub = ap.add_subparsers(dest="cmd", required=True)
    t = sub.add_parser("table", help="minimum-bias table (data/sn_ecc_eccp_10.dat) -> centrality cut table; sorts on the GPU")
    t.add_argument("data_dir"); t.add_argument("--cut", default="total_entropy"); t.add_argument("--alpha", type=float, default=0.118)
    t.add_argument("--name", default=None, help="table name suffix (default: the directory name)")
    r = sub.add_parser("run", help="averaged (operation 3) or event-by-event (operation 2) profiles for one centrality bin")
    r.add_argument("--model", default="MCGlb", choices=["MCGlb", "MCKLN", "Trento"]); r.add_argument("--ecm", type=float, default=2760)
    r.add_argument("--collsys", nargs=2, default=["Pb", "Pb"]); r.add_argument("--cen", default="0-5", help="lower-upper in %%")
    r.add_argument("--cut", default="total_entropy", choices=["total_entropy", "Npart"]); r.add_argument("--tables", required=True)
    r.add_argument("--operation", type=int, default=3); r.add_argument("--nev", type=int, default=1000)
    r.add_argument("--gpus", type=int, default=1); r.add_argument("--dry-run", action="store_true")
    r.add_argument("extra", nargs="*", help="further name=value parameters")
    m = sub.add_parser("minbias", help="minimum-bias scan in memory -> centrality cut table (no text tables in between).  Under torchrun "
                                       "every rank takes its range of global event ids, rank 0 gathers the per-event rows "
                                       "(smc_comm_gather_doubles) and sorts them on its GPU")
    m.add_argument("--model", default="MCGlb", choices=["MCGlb", "MCKLN", "Trento"]); m.add_argument("--ecm", type=float, default=2760)
    m.add_argument("--collsys", nargs=2, default=["Pb", "Pb"]); m.add_argument("--nev", type=int, default=100000)
    m.add_argument("--cut", default="total_entropy"); m.add_argument("--seed", type=int, default=1); m.add_argument("--out", default=".")
    m.add_argument("extra", nargs="*", help="further name=value parameters (names of smc_params, e.g. maxx=13)")
    a = ap.parse_args(argv)
    if a.cmd == "minbias":
        return minbias_table(a)
    if a.cmd == "table":
        import supermc_b200 as smc
        rows = np.loadtxt(os.path.join(a.data_dir, "sn_ecc_eccp_10.dat"))
        en = os.path.join(a.data_dir, "en_ecc_eccp_10.dat")
        coll = collision_data(rows, np.loadtxt(en) if os.path.exists(en) else None)
        ctx = smc.Context(smc.capi.default_params(max_batch=8))
        text = centrality_table_text(coll, device_order(ctx, coll, a.cut, a.alpha), a.cut, a.alpha)
        ctx.close()
        name = a.name or os.path.basename(os.path.abspath(a.data_dir))
        out = os.path.join(a.data_dir, "iebe_centralityCut_%s_%s.dat" % (a.cut, name))
        open(out, "w").write(text)
        print(out)
        return 0
    p = model_parameters(a.model, a.ecm, a.collsys)
    lower, upper = (float(x) for x in a.cen.split("-"))
    fn = os.path.join(a.tables, table_file_name(a.cut, p["which_mc_model"], p["Aproj"], p["Atarg"], p["ecm"], p["cc_fluctuation_model"]))
    if not os.path.exists(fn):
        print("Can not find the centrality cut table for the collision system\n" + fn, file=sys.stderr)
        return 1
    p.update(translate_centrality_cut(np.loadtxt(fn), lower, upper, a.cut))
    p.update(operation=a.operation, nev=a.nev, finalFactor=1.0, use_sd=1, use_ed=1)
    args = assignment_args(p) + list(a.extra)
    if a.gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
               "-m", "supermc_b200.launch", "parameters.dat"] + args
    else:
        cmd = [os.path.join(HERE, "superMC_b200.e")] + args
    print(" ".join(cmd))
    if a.dry_run:
        return 0
    os.makedirs("data", exist_ok=True)
    return subprocess.call(cmd, env=dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", "")))


if __name__ == "__main__":
    sys.exit(main())
