"""Shared test plumbing: golden fixtures -> oracle configs / C-ABI parameters / event inputs."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

KLN_SYSTEM = "auau200_kln"
SYSTEMS = ["pbpb2760_glb", "auau200_glb_quarks", "ppb5020_glb_quarks", "pbpb2760_sqrt_disk", "pbpb2760_uli",
           "auau200_disk_nucleons", "he3au200_glb", "cc200_glb", "uu193_deformed", "pbpb5020_lambda_width", "cuau200_glb", "oo200_glb", "auau200_nncorr", "pbpb2760_rotate"]


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.par = {str(k).lower(): float(v) for k, v in zip(self.z["params_keys"], self.z["params_vals"])}
        self.consts = self.z["consts"]
        self.ntries = int(self.z["ntries"])
        self.ecc_rows = self.z["ecc_rows"]
        self.quark_kind = str(self.z["quark_kind"])

    def tr(self, i):
        pre = "t%d/" % i
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def tries(self):
        return [self.tr(i) for i in range(self.ntries)]

    def oracle_cfg(self, port):
        p = self.par
        return port.make_cfg(maxx=p["maxx"], maxy=p["maxy"], dx=p["dx"], dy=p["dy"], ecm=p["ecm"], alpha=p.get("alpha", 0.118),
                             shape_of_nucleons=int(p["shape_of_nucleons"]), shape_of_entropy=int(p["shape_of_entropy"]),
                             collision_criterion=int(p["collision_criterion"]), which_mc_model=int(p["which_mc_model"]),
                             sub_model=int(p["sub_model"]), cc_fluct_model=int(p["cc_fluctuation_model"]),
                             gaussian_lambda=p.get("gaussian_lambda", 4.14))

    def smc_params(self, capi, **over):
        p = self.par
        kw = dict(which_mc_model=int(p["which_mc_model"]), sub_model=int(p["sub_model"]), alpha=p.get("alpha", 0.118),
                  aproj=int(p["aproj"]), atarg=int(p["atarg"]), proj_deformed=int(p.get("proj_deformed", 0)),
                  targ_deformed=int(p.get("targ_deformed", 0)), shape_of_nucleons=int(p["shape_of_nucleons"]),
                  collision_criterion=int(p["collision_criterion"]), shape_of_entropy=int(p["shape_of_entropy"]),
                  ecm=p["ecm"], bmin=p["bmin"], bmax=p["bmax"], npmin=int(p["npmin"]), npmax=int(p["npmax"]),
                  finalfactor=p["finalfactor"], maxx=p["maxx"], maxy=p["maxy"], dx=p["dx"], dy=p["dy"],
                  cc_fluctuation_model=int(p["cc_fluctuation_model"]),
                  cc_fluctuation_gamma_theta=p.get("cc_fluctuation_gamma_theta", 0.75), randomseed=int(p["randomseed"]),
                  include_nn_correlation=int(p.get("include_nn_correlation", 0)))
        if "gaussian_lambda" in p:
            kw["gaussian_lambda"] = p["gaussian_lambda"]
        if "lambda" in p:
            kw.update({"lambda": p["lambda"], "tmax": int(p["tmax"]), "tmax_subdivision": int(p["tmax_subdivision"])})
        if "ny" in p:
            kw.update(ny=int(p["ny"]), ymax=p["ymax"])
        kw.update(over)
        return capi.default_params(**kw)


def src8_from(nuc, idx):
    """participant rows for the oracle deposits: x y xL xR yL yR weight extra"""
    idx = np.asarray(idx, dtype=int)
    s = np.zeros((len(idx), 8))
    s[:, 0:2] = nuc[idx, 0:2]; s[:, 2:6] = nuc[idx, 3:7]; s[:, 6] = nuc[idx, 8]
    return s


def coll8_from(coll):
    c = np.zeros((len(coll), 8))
    if len(coll):
        c[:, 0:2] = coll[:, 0:2]; c[:, 6] = coll[:, 2]; c[:, 7] = coll[:, 3]
    return c


def event_in_from(t, port=None, cfg=None, with_uniforms=True):
    """golden try -> dict for Context.run_from_positions (nucleon rows x y z xL xR yL yR weight)"""
    proj, targ = t["proj"], t["targ"]
    p8 = np.concatenate([proj[:, :7], proj[:, 8:9]], axis=1)
    t8 = np.concatenate([targ[:, :7], targ[:, 8:9]], axis=1)
    ev = dict(b=float(t["hdr"][0]), proj=p8, targ=t8, given_w=1)
    coll = t["coll"]
    if len(coll):
        ev["coll_weight"] = coll[:, 2:4].copy()
    if with_uniforms and port is not None and int(cfg.collision_criterion) != 1:
        st = port.Stream48(state=t["hdr"][5:8])
        r = port.collide(cfg, proj[:, :7], targ[:, :7], stream=st, want_u=True)
        u = r["u"]
        ev["pair_uniform"] = np.where(u < 0, 2.0, u)     # never-tested pairs can never hit
    return ev


def rel_err(a, ref):
    """per-cell relative error against max(|ref|, 1e-12 max|ref|)  (level L2)"""
    scale = np.maximum(np.abs(ref), 1e-12 * np.abs(ref).max())
    return np.abs(a - ref) / scale
