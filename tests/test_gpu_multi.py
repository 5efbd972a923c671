"""-m gpu, needs >= 2 GPUs (skipped otherwise): the torchrun launcher on 2 ranks reproduces the 1-GPU
operation-9 table byte for byte, and the all-reduced operation-3 average equals the 1-GPU average."""
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
