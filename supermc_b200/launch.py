"""Multi-GPU launcher: one process per GPU (torchrun), events sharded by global event id.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m supermc_b200.launch [parameters.dat] name=value ...

The reference's only parallel mode is "start 8 copies with different seeds"
(CollectDataAccordingToSettings.py:110-115); here every rank owns the contiguous event-id range
[rank*nev/world, (rank+1)*nev/world) of ONE Philox-keyed run, so the set of events does not depend on
the number of GPUs.  Data path: no collective for operations 1, 2, 9 (per-event rows and files are
independent; rank 0 concatenates the operation-9 tables in rank order, which reproduces the 1-GPU
file byte for byte).  Operation 3: the accumulator sums that live in each GPU's memory are combined
with one all-reduce behind the C ABI (smc_avg_allreduce: ncclAllReduce over NVLink), then rank 0 writes.
The helper functions below restate the host logic for the CPU tests (gloo, world size 2).
"""
import ctypes as C
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def shard_range(nev, rank, world):
    """global event ids [lo, hi) of `rank` -- must match MakeDensity::shard_range (host/MakeDensity.cpp)"""
    return nev * rank // world, nev * (rank + 1) // world


class _DevBuf:
    """zero-copy view of a device pointer for torch.as_tensor"""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def allreduce_sums(buf, count, dist):
    """sum the accumulator block and the accepted-event counter over all ranks (in place); returns the
    global count.  `buf` is a torch tensor (CUDA with NCCL; CPU with gloo in the tests)."""
    import torch
    cnt = torch.tensor([count], dtype=torch.int64, device=buf.device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return int(cnt.item())


def merge_rank_tables(data_dir, world, pattern="*_ecc_eccp_*.dat"):
    """append data_rank<r>/<table> (r = 1..world-1, rank order) to data/<table>, then drop the rank dirs"""
    for r in range(1, world):
        rd = "%s_rank%d" % (data_dir, r)
        for f in sorted(glob.glob(os.path.join(rd, pattern))):
            with open(os.path.join(data_dir, os.path.basename(f)), "ab") as dst, open(f, "rb") as src:
                dst.write(src.read())
            os.remove(f)
        try:
            os.rmdir(rd)
        except OSError:
            pass


def _host():
    L = C.CDLL(os.path.join(HERE, "libsupermc_host.so"))
    L.smc_host_create.restype = C.c_void_p
    L.smc_host_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.smc_host_error.restype = C.c_char_p
    L.smc_host_error.argtypes = [C.c_void_p]
    L.smc_host_context.restype = C.c_void_p
    L.smc_host_context.argtypes = [C.c_void_p]
    for f in ("smc_host_run", "smc_host_average_accumulate", "smc_host_average_write", "smc_host_destroy"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.smc_host_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
    return L


def main(argv=None):
    """One rank of a multi-GPU run (torchrun sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Everything, including the
    rendezvous, the single seed of a randomSeed < 0 run, the operation-3 all-reduce (smc_avg_allreduce: NCCL over NVLink)
    and the rank-ordered merge of the output files, happens in the C++ host layer (host/MakeDensity.cpp) behind the C ABI:
    this module only forwards argv.  `torchrun --no-python supermc_b200/superMC_b200.e ...` is the same thing."""
    argv = list(sys.argv[1:] if argv is None else argv)
    pfile = "parameters.dat"
    if argv and "=" not in argv[0]:
        pfile = argv.pop(0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    L = _host()
    args = (C.c_char_p * max(len(argv), 1))(*[a.encode() for a in argv])
    h = L.smc_host_create(pfile.encode(), len(argv), args, local, rank, world, b"data")
    try:
        if not L.smc_host_context(h):
            raise RuntimeError(L.smc_host_error(h).decode())
        rc = L.smc_host_run(h)
        if rc != 0:
            raise RuntimeError(L.smc_host_error(h).decode())
    finally:
        L.smc_host_destroy(h)
    return rc


if __name__ == "__main__":
    sys.exit(main())
