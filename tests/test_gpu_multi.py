"""-m gpu: several ranks of ONE run.
* test_two_ranks_on_one_gpu: the drop-in executable superMC_b200.e as two processes (RANK 0/1, both on cuda:0 -- runs on
  a 1-GPU box): the merged operation-9 tables are byte-identical to the 1-rank run, per-event files of operation 1 carry
  global event ids, and the operation-3 average after smc_avg_allreduce (the peer-memory kernel over CUDA IPC, because
  NCCL refuses two ranks on one device) equals the 1-rank average.
* test_two_ranks_equal_one_rank: the same through torchrun on two GPUs (NCCL all-reduce); skipped on a 1-GPU box."""
import os
import shutil
import subprocess
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["which_mc_model=5", "sub_model=1", "Aproj=208", "Atarg=208", "ecm=2760", "alpha=0.118", "maxx=13", "maxy=13",
        "finalFactor=1", "randomSeed=5", "cc_fluctuation_model=6", "use_ed=0"]


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(tmp, world, extra, port):
    d = tmp / ("w%d" % world); os.makedirs(d / "data")
    shutil.copy(os.path.join(ROOT, "supermc_b200", "parameters.dat"), d)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "supermc_b200.launch", "parameters.dat"] + ARGS + extra
    subprocess.check_call(cmd, cwd=d, env=env, stdout=subprocess.DEVNULL)
    return d / "data"


def test_two_ranks_equal_one_rank(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    a = _run(tmp_path, 1, ["operation=9", "nev=501"], 29611); b = _run(tmp_path, 2, ["operation=9", "nev=501"], 29612)
    for f in ("sn_ecc_eccp_10.dat", "sn_ecc_eccp_2.dat"):
        assert (a / f).read_bytes() == (b / f).read_bytes()
    a = _run(tmp_path / "avg", 1, ["operation=3", "nev=64", "bmin=6", "bmax=8", "average_to_order=2"], 29613)
    b = _run(tmp_path / "avg", 2, ["operation=3", "nev=64", "bmin=6", "bmax=8", "average_to_order=2"], 29614)
    for f in ("sdAvg_order_2_block.dat", "sdAvg_RP_order_2_block.dat", "TATB_fromSd_order_2_block.dat", "spectator_density_A_fromSd_order_2_block.dat"):
        x, y = np.loadtxt(a / f), np.loadtxt(b / f)
        assert np.allclose(x, y, rtol=1e-10, atol=1e-14), f


def _exe_run(tmp, world, extra, port, ngpu=1):
    d = tmp / ("x%d" % world); os.makedirs(d / "data")
    shutil.copy(os.path.join(ROOT, "supermc_b200", "parameters.dat"), d)
    exe = os.path.join(ROOT, "supermc_b200", "superMC_b200.e")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r % ngpu), MASTER_ADDR="127.0.0.1", SMC_COMM_PORT=str(port))
        procs.append(subprocess.Popen([exe] + ARGS + extra, cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    return d / "data", outs


def test_two_ranks_on_one_gpu(tmp_path):
    a, _ = _exe_run(tmp_path / "t9", 1, ["operation=9", "nev=501"], 29711)
    b, _ = _exe_run(tmp_path / "t9", 2, ["operation=9", "nev=501"], 29712)
    assert sorted(os.listdir(a)) == sorted(os.listdir(b)) and not os.path.exists(str(b) + "_rank1")
    for f in sorted(os.listdir(a)):
        assert (a / f).read_bytes() == (b / f).read_bytes(), f
    assert len((a / "sn_ecc_eccp_10.dat").read_bytes().splitlines()) == 501
    # operation 1: per-event files with global ids, binary.dat appended in rank order
    e1 = ["operation=1", "nev=7", "use_4col=0", "use_block=1"]
    a, _ = _exe_run(tmp_path / "t1", 1, e1, 29713); b, _ = _exe_run(tmp_path / "t1", 2, e1, 29714)
    assert sorted(os.listdir(a)) == sorted(os.listdir(b))
    for f in sorted(os.listdir(a)):
        assert (a / f).read_bytes() == (b / f).read_bytes(), f
    # operation 3: accumulators summed across the two ranks
    e3 = ["operation=3", "nev=64", "bmin=6", "bmax=8", "average_to_order=2", "output_TATB=1", "output_spectator_density=1", "use_4col=0"]
    a, _ = _exe_run(tmp_path / "t3", 1, e3, 29715); b, outs = _exe_run(tmp_path / "t3", 2, e3, 29716)
    files = sorted(os.listdir(a))
    assert files == sorted(os.listdir(b)) and len(files) >= 6
    for f in files:
        x, y = np.loadtxt(a / f), np.loadtxt(b / f)
        assert np.allclose(x, y, rtol=1e-10, atol=1e-14), f


def _minbias(tmp, world, port, nev=3000):
    d = tmp / ("m%d" % world); os.makedirs(d)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", SMC_COMM_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-m", "supermc_b200.centrality", "minbias", "--nev", str(nev), "--seed", "5", "--out", str(d)],
                                      cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    f = [x for x in os.listdir(d) if x.startswith("iebe_centralityCut_")]
    assert len(f) == 1
    return d / f[0]


def test_minbias_centrality_table_in_memory(tmp_path):
    """python -m supermc_b200.centrality minbias: per-event rows gathered on rank 0 (smc_comm_gather_doubles) and sorted on
    its GPU.  Two ranks give the one-rank table byte for byte; the table agrees with the one built from the executable's
    text tables of the same run (those pass through 8 printed digits)."""
    one = _minbias(tmp_path, 1, 29731); two = _minbias(tmp_path, 2, 29732)
    assert one.read_bytes() == two.read_bytes()
    t = np.loadtxt(one)
    assert t.shape == (110, 6) and np.all(np.diff(t[1:, 1]) <= 0) and t[-1, 0] == 100.0
    d, _ = _exe_run(tmp_path / "txt", 1, ["operation=9", "nev=3000", "use_ed=1"], 29733)
    subprocess.check_call([sys.executable, "-m", "supermc_b200.centrality", "table", str(d)], env=dict(os.environ, PYTHONPATH=ROOT), stdout=subprocess.DEVNULL)
    ref = np.loadtxt(d / "iebe_centralityCut_total_entropy_data.dat")
    assert np.allclose(t, ref, rtol=2e-6, atol=1e-6)
