"""-m gpu: the drop-in executable superMC_b200.e (C++ host layer over the C ABI): operation 9 and 2
write the reference's data/ layouts, and the rows equal what the C ABI returns for the same event ids."""
import os
import shutil
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "supermc_b200", "superMC_b200.e")
ARGS = ["which_mc_model=5", "sub_model=1", "Aproj=208", "Atarg=208", "ecm=2760", "alpha=0.118", "maxx=13", "maxy=13",
        "finalFactor=1", "randomSeed=5", "cc_fluctuation_model=6"]


def _rundir(tmp_path):
    os.makedirs(tmp_path / "data")
    shutil.copy(os.path.join(ROOT, "supermc_b200", "parameters.dat"), tmp_path)
    return tmp_path


def test_operation9_table(tmp_path):
    import supermc_b200 as smc
    d = _rundir(tmp_path)
    subprocess.check_call([EXE] + ARGS + ["operation=9", "nev=300"], cwd=d, stdout=subprocess.DEVNULL)
    t10 = np.loadtxt(d / "data" / "sn_ecc_eccp_10.dat")
    assert t10.shape == (300, 49)
    assert (d / "data" / "en_ecc_eccp_10.dat").read_bytes() == (d / "data" / "sn_ecc_eccp_10.dat").read_bytes()   # quirk Q2
    t2 = np.loadtxt(d / "data" / "sn_ecc_eccp_2.dat")
    assert t2.shape == (300, 9) and np.array_equal(t2[:, 0:5], t10[:, 5:10])
    ctx = smc.Context(smc.capi.default_params(which_mc_model=5, sub_model=1, ecm=2760.0, alpha=0.118, maxx=13.0, maxy=13.0,
                                              finalfactor=1.0, randomseed=5, cc_fluctuation_model=6))
    ev = ctx.run_events(0, 300)
    assert np.array_equal(t10[:, 45], ev["npart1"] + ev["npart2"]) and np.array_equal(t10[:, 46], ev["ncoll"])
    assert np.allclose(t10[:, :45], ev["mom"].reshape(300, 45), rtol=2e-7, atol=1e-12)      # 8 printed digits
    assert np.allclose(t10[:, 48], ev["b"], rtol=2e-7)
    # the tables of the executable's pipeline (every number formatted once, rows assembled from cells) are, byte for byte,
    # the rows of the formatter that tests/test_host_layer.py pins to reference-written files
    import ctypes as C
    host = C.CDLL(os.path.join(ROOT, "supermc_b200", "libsupermc_host.so"))
    buf = C.create_string_buffer(4096)
    for order in (1, 2, 9, 10):
        want = b""
        for e in range(300):
            eo = smc.EventOut.from_buffer_copy(ev[e].tobytes())
            n = host.smc_host_format_ecc_row(C.byref(eo), order, 0, buf, 4096)
            assert n > 0
            want += buf.value
        assert (d / "data" / ("sn_ecc_eccp_%d.dat" % order)).read_bytes() == want, order


def test_operation2_files(tmp_path):
    d = _rundir(tmp_path)
    subprocess.check_call([EXE] + ARGS + ["operation=2", "nev=3", "use_4col=1"], cwd=d, stdout=subprocess.DEVNULL)
    for k in (1, 2, 3):
        for stem in ("sd_event_%d", "ed_event_%d", "rho_binary_event_%d", "nuclear_thickness_TA_event_%d", "nuclear_thickness_TB_event_%d",
                     "rhob_event_%d", "spectator_density_A_event_%d", "spectator_density_B_event_%d"):
            blk = np.loadtxt(d / "data" / ((stem % k) + "_block.dat"))
            assert blk.shape == (261, 261)
            col = np.loadtxt(d / "data" / ((stem % k) + "_4col.dat"))
            assert col.shape == (261 * 261, 4) and np.allclose(col[:, 3].reshape(261, 261), blk, rtol=1e-11)
        ta = np.loadtxt(d / "data" / ("nuclear_thickness_TA_event_%d_block.dat" % k)); tb = np.loadtxt(d / "data" / ("nuclear_thickness_TB_event_%d_block.dat" % k))
        assert np.allclose(np.loadtxt(d / "data" / ("rhob_event_%d_block.dat" % k)), ta + tb, rtol=1e-11)
        parts = np.loadtxt(d / "data" / ("ParticipantTable_event_%d.dat" % k)); row = np.loadtxt(d / "data" / ("sn_ecc_eccp_10_event_%d.dat" % k))
        assert len(parts) == int(row[45])
        assert abs(ta.sum() * 0.01 - (parts[:, 2] == 1).sum()) < 0.05 * (parts[:, 2] == 1).sum() + 1   # each participant deposits ~1
        assert len(np.loadtxt(d / "data" / ("BinaryCollisionTable_event_%d.dat" % k)).reshape(-1, 2)) == int(row[46])
        assert len(np.loadtxt(d / "data" / ("Spectators_event_%d.dat" % k))) == 416 - int(row[45])
        sd = np.loadtxt(d / "data" / ("sd_event_%d_block.dat" % k))
        assert abs(sd.sum() * 0.01 - row[47]) < 1e-6 * row[47]


def test_operation1_and_3_files(tmp_path):
    d = _rundir(tmp_path)
    subprocess.check_call([EXE] + ARGS + ["operation=1", "nev=2", "use_ed=0"], cwd=d, stdout=subprocess.DEVNULL)
    for k in (1, 2):
        assert np.loadtxt(d / "data" / ("sd_event_%d_block.dat" % k)).shape == (261, 261)
    assert len(np.loadtxt(d / "data" / "nucl1.data")) == 208 and (d / "data" / "binary.dat").stat().st_size > 0
    d3 = _rundir(tmp_path / "op3")
    subprocess.check_call([EXE] + ARGS + ["operation=3", "nev=40", "bmin=6", "bmax=8", "average_from_order=2", "average_to_order=3"], cwd=d3, stdout=subprocess.DEVNULL)
    names = sorted(os.listdir(d3 / "data"))
    assert len(names) == 48, names            # the reference writes 48 files for 2 orders with its default switches
    sd = np.loadtxt(d3 / "data" / "sdAvg_order_2_block.dat"); ed = np.loadtxt(d3 / "data" / "edAvg_order_2_block.dat")
    assert sd.shape == (261, 261) and sd.sum() > 0 and abs(ed.sum() / sd.sum() - 1) < 0.05


def test_light_ion_tables_through_the_executable(tmp_path):
    d = _rundir(tmp_path)
    os.makedirs(d / "tables")
    raw = np.loadtxt(os.path.join(ROOT, "tests", "golden", "carbon_configs_small.txt"))
    np.savetxt(d / "tables" / "carbon_plaintext.dat", raw, fmt="%.6f")
    rng = np.random.default_rng(3)
    np.savetxt(d / "tables" / "oxygen_plaintext.dat", rng.normal(0, 1.6, (50, 48)), fmt="%.6f")      # synthetic: the O table is a missing blob upstream
    subprocess.check_call([EXE, "which_mc_model=5", "sub_model=1", "Aproj=12", "Atarg=12", "ecm=200", "maxx=13", "maxy=13", "operation=9", "nev=100",
                           "randomSeed=2", "finalFactor=1", "bmax=8"], cwd=d, stdout=subprocess.DEVNULL)
    t = np.loadtxt(d / "data" / "sn_ecc_eccp_10.dat")
    assert t.shape == (100, 49) and t[:, 45].max() <= 24 and t[:, 45].min() >= 2
    for f in os.listdir(d / "data"):
        os.remove(d / "data" / f)
    subprocess.check_call([EXE, "which_mc_model=5", "sub_model=1", "Aproj=16", "Atarg=16", "ecm=200", "maxx=13", "maxy=13", "operation=9", "nev=50",
                           "randomSeed=2", "finalFactor=1", "bmax=8"], cwd=d, stdout=subprocess.DEVNULL)
    assert np.loadtxt(d / "data" / "sn_ecc_eccp_10.dat").shape == (50, 49)


def test_rcbk_tables_read_from_javier_directory(tmp_path):
    """sub_model=101: the executable reads javier/ft_rcbk_mv_qs02_0168_g1_119_*.dat (rcBKfunc.cpp:115-178) and appends
    the dN/dy table to data/dNdyTable.dat (MCnucl.cpp:1027-1048); a missing file is the reference's fatal error."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rcbk_synth
    import supermc_b200 as smc
    d = _rundir(tmp_path)
    args = [EXE, "which_mc_model=1", "sub_model=101", "Aproj=1", "Atarg=1", "ecm=2760", "maxx=6", "maxy=6", "tmax=8", "tmax_subdivision=3",
            "operation=9", "nev=20", "randomSeed=2", "finalFactor=1", "bmax=1", "cc_fluctuation_model=0"]
    r = subprocess.run(args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode != 0 and b"unable to open file javier/" in r.stdout
    kt, na = rcbk_synth.write_files(str(d / "javier"), 101)
    subprocess.check_call(args, cwd=d, stdout=subprocess.DEVNULL)
    tab = np.loadtxt(d / "data" / "dNdyTable.dat")
    ctx = smc.Context(smc.capi.default_params(which_mc_model=1, sub_model=101, aproj=1, atarg=1, ecm=2760.0, maxx=6.0, maxy=6.0, tmax=8,
                                              tmax_subdivision=3, cc_fluctuation_model=0, max_batch=8))
    ctx.load_rcbk_tables(kt, na)
    T = ctx.build_kln_table()
    n = T.shape[0]
    assert tab.shape == ((n - 1) * (n - 1), 4)
    assert np.allclose(tab[:, 3].reshape(n - 1, n - 1), T[1:, 1:], rtol=0, atol=1e-11)      # %22.12f
    assert np.loadtxt(d / "data" / "sn_ecc_eccp_10.dat").shape == (20, 49)


def test_centrality_tools_end_to_end(tmp_path):
    """minimum-bias run -> centrality table (sorted on the GPU) -> per-centrality averaged profiles through the wrapper
    (scripts/centrality_cut_h5.py + generateAvgprofile.py, natively: python -m supermc_b200.centrality)"""
    import sys
    from supermc_b200 import centrality as cen
    d = _rundir(tmp_path)
    subprocess.check_call([EXE] + ARGS + ["operation=9", "nev=4000"], cwd=d, stdout=subprocess.DEVNULL)     # >= 3000: the 0.1 % bins must not be empty
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, "-m", "supermc_b200.centrality", "table", "data"], cwd=d, env=env, stdout=subprocess.DEVNULL)
    tab = np.loadtxt(d / "data" / "iebe_centralityCut_total_entropy_data.dat")
    assert tab.shape == (110, 6) and np.all(np.diff(tab[1:, 1]) <= 0) and tab[-1, 0] == 100.0
    rows = np.loadtxt(d / "data" / "sn_ecc_eccp_10.dat")
    assert abs(tab[19, 1] - np.sort(rows[:, 47])[::-1][int(4000 * 0.10) - 2]) < 1e-3 * tab[19, 1]      # the 10 % row: smallest dS/dy of the 9-10 % bin
    os.makedirs(d / "tabs")
    shutil.copy(d / "data" / "iebe_centralityCut_total_entropy_data.dat", d / "tabs" / cen.table_file_name("total_entropy", 5, 208, 208, 2760.0, 6))
    for f in os.listdir(d / "data"):
        os.remove(d / "data" / f)
    subprocess.check_call([sys.executable, "-m", "supermc_b200.centrality", "run", "--model", "MCGlb", "--ecm", "2760", "--collsys", "Pb", "Pb",
                           "--cen", "0-10", "--tables", "tabs", "--nev", "12", "maxx=13", "maxy=13", "average_from_order=2", "average_to_order=2"],
                          cwd=d, env=env, stdout=subprocess.DEVNULL)
    g = np.loadtxt(d / "data" / "sdAvg_order_2_block.dat")
    assert g.shape == (261, 261) and g.sum() * 0.01 > tab[19, 1] * 0.9          # every averaged event passed the 0-10 % dS/dy cut


def test_mcnucl_adapter_drives_a_reference_style_loop(tmp_path):
    """host/MCnuclB200.h (the binding INTEGRATION.md describes) under a MakeDensity-style loop: the reference's own grid loops
    over getRho reproduce the columns the engine computed on the device, and the events are those of the C ABI"""
    import supermc_b200 as smc
    d = _rundir(tmp_path)
    out = subprocess.run([os.path.join(ROOT, "supermc_b200", "adapter_demo.e"), "100"] + ARGS, cwd=d, stdout=subprocess.PIPE, check=True, text=True).stdout
    t = np.array([[float(x) for x in ln.split()] for ln in out.splitlines() if ln and ln[0].isdigit()])
    assert t.shape == (100, 10)
    ctx = smc.Context(smc.capi.default_params(which_mc_model=5, sub_model=1, ecm=2760.0, alpha=0.118, maxx=13.0, maxy=13.0,
                                              finalfactor=1.0, randomseed=5, cc_fluctuation_model=6))
    ev = ctx.run_events(0, 100)
    assert np.array_equal(t[:, 1], ev["npart1"]) and np.array_equal(t[:, 2], ev["npart2"]) and np.array_equal(t[:, 3], ev["ncoll"])
    assert np.allclose(t[:, 9], ev["b"], rtol=1e-14)
    assert np.allclose(t[:, 4], t[:, 5], rtol=1e-12) and np.allclose(t[:, 5], ev["total"], rtol=1e-12)      # host loop over getRho == device sum
    assert np.allclose(t[:, 6], t[:, 7], rtol=1e-10)                                                         # <r^2> about the centre of mass
    assert (t[:, 8] > 0).all()                                                                               # TA1 lattice present
