#!/usr/bin/env python3
"""Golden vectors for supermc_b200/centrality.py, produced by the UNMODIFIED reference scripts (build container only):

* scripts/centrality_cut_h5.py is run as a program on the 10^5 reference events of ks_pbpb2760_ref.npz.  It needs h5py,
  which is not installed: a 10-line stand-in module (File(path).get("collision_data") -> array) is put on PYTHONPATH,
  the script itself is untouched.  Its two output tables are stored verbatim.
* scripts/generateAvgprofile.py is imported and translate_centrality_cut is called on shipped tables
  (scripts/centrality_cut_tables/) for a list of centrality windows; inputs (the table arrays) and outputs are stored.
"""
import importlib.util
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/scripts"
out = {}

# ---- centrality_cut_h5.py ----
z = np.load(os.path.join(ROOT, "tests", "golden", "ks_pbpb2760_ref.npz"))
coll = np.stack([z["b"], z["npart"], z["ncoll"], z["dsdy"], z["dsdy"]], axis=1).astype(np.float32)[:20000]
d = tempfile.mkdtemp()
open(os.path.join(d, "h5py.py"), "w").write(
    "import numpy as np\n"
    "class File:\n"
    "    def __init__(self, path, mode='r'): self.path = path\n"
    "    def get(self, name): return np.load(self.path + '.' + name + '.npy')\n"
    "    def close(self): pass\n")
np.save(os.path.join(d, "minbias.h5.collision_data.npy"), coll)
subprocess.check_call([sys.executable, os.path.join(REF, "centrality_cut_h5.py"), os.path.join(d, "minbias.h5")],
                      env=dict(os.environ, PYTHONPATH=d), stdout=subprocess.DEVNULL)
out["coll"] = coll
for cut in ("total_entropy", "Npart"):
    out["table_" + cut] = np.array(open(os.path.join(d, "iebe_centralityCut_%s_minbias.dat" % cut)).read())

# ---- generateAvgprofile.translate_centrality_cut ----
spec = importlib.util.spec_from_file_location("ref_avg", os.path.join(REF, "generateAvgprofile.py"))
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
cases = []
os.chdir(REF)
for model, ecm, sys_, cut, windows in (("MCGlb", 2760, ("Pb", "Pb"), "total_entropy", [(0, 5), (5, 10), (20, 30), (0.2, 0.7), (60, 80)]),
                                       ("MCGlb", 200, ("Au", "Au"), "Npart", [(0, 5), (10, 20), (40, 50)]),
                                       ("MCKLN", 200, ("Au", "Au"), "total_entropy", [(0, 10), (30, 40)])):
    ref.update_superMC_dict(model, ecm, sys_)
    for w in windows:
        sys.stdout = open(os.devnull, "w")
        try:
            ref.translate_centrality_cut(w, cut)
        finally:
            sys.stdout = sys.__stdout__
        p = ref.superMCParameters
        name = "iebe_centralityCut_%s_%s_sigmaNN_gauss_d0.9_%s.dat" % (
            cut, model + sys_[0] + sys_[1] + ("%g" % ecm), "withMultFluct" if p["cc_fluctuation_model"] != 0 else "noMultFluct")
        key = "tab%d" % len(cases)
        out[key] = np.loadtxt(os.path.join("centrality_cut_tables", name))
        cases.append([key, name, cut, w[0], w[1], p["which_mc_model"], p["Aproj"], p["Atarg"], ecm, p["cc_fluctuation_model"],
                      p["cutdSdy"], p.get("cutdSdy_lowerBound", 0.0), p.get("cutdSdy_upperBound", 0.0), p["Npmin"], p["Npmax"], p["bmin"], p["bmax"]])
out["cases"] = np.array(cases, dtype=object).astype(str)
path = os.path.join(ROOT, "tests", "golden", "centrality.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path) // 1024, "KB", len(cases), "cases")
