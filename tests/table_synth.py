"""Synthetic nucleon-configuration tables in the reference's file formats (TEST INFRASTRUCTURE).

The reference reads tables/oxygen_plaintext.dat (src/Nucleus.cpp:462-478: 48 coordinates per configuration),
tables/au197-sw-full_3Bchains-conf1820.dat (1820 configurations x 197 rows "x y z isospin dummy") and
tables/pb208-1.dat (10000 x 208 rows "x y z isospin"), src/Nucleus.cpp:481-522.  All three are missing blobs upstream
(.MISSING_LARGE_BLOBS), so parity on those code paths uses stand-ins generated here from a fixed legacy numpy seed
(RandomState streams are frozen by numpy's compatibility policy).  Coordinates are multiples of 1e-4 fm so that the
text files ("%.4f") and the in-memory tables hold the same doubles.
"""
import os
import numpy as np


def _quantise(a):
    return np.rint(a * 1e4) / 1e4


def _woods_saxon(rs, n, R, a):
    """n radii from r^2 / (1 + exp((r - R)/a)) on [0, R + 10 a] by rejection"""
    out = np.empty(0)
    rmax = R + 10 * a
    while len(out) < n:
        r = rmax * rs.random_sample(2 * n) ** (1.0 / 3.0)
        keep = rs.random_sample(2 * n) < 1.0 / (1.0 + np.exp((r - R) / a))
        out = np.concatenate([out, r[keep]])
    return out[:n]


def _cloud(rs, ncfg, A, R, a):
    r = _woods_saxon(rs, ncfg * A, R, a)
    ct = 1 - 2 * rs.random_sample(ncfg * A); ph = 2 * np.pi * rs.random_sample(ncfg * A)
    st = np.sqrt(1 - ct * ct)
    xyz = np.stack([r * st * np.cos(ph), r * st * np.sin(ph), r * ct], axis=1).reshape(ncfg, A, 3)
    xyz += rs.normal(0.0, 0.15, (ncfg, 1, 3))        # configurations are not centred: the NN-correlated sampler recentres
    return _quantise(xyz)


def oxygen(ncfg=400):
    return _cloud(np.random.RandomState(160016), ncfg, 16, 2.608, 0.513).reshape(ncfg, 48)


def au197(ncfg=1820):
    """the reference hard-codes n_configuration = 1820 (Nucleus.cpp:489)"""
    return _cloud(np.random.RandomState(197197), ncfg, 197, 6.42, 0.45).reshape(ncfg, 197 * 3)


def pb208(ncfg=10000):
    """the reference hard-codes n_configuration = 10000 (Nucleus.cpp:494)"""
    return _cloud(np.random.RandomState(208208), ncfg, 208, 6.67, 0.44).reshape(ncfg, 208 * 3)


def write_oxygen(path, cfg):
    with open(path, "w") as f:
        for row in cfg:
            f.write(" ".join("%.4f" % v for v in row) + "\n")


def write_nncorr(path, cfg, A, with_dummy):
    rs = np.random.RandomState(7)
    iso = rs.randint(0, 2, size=A)
    with open(path, "w") as f:
        for row in cfg:
            r = row.reshape(A, 3)
            if with_dummy:
                f.write("".join("%.4f %.4f %.4f %d %d\n" % (r[i, 0], r[i, 1], r[i, 2], iso[i], i) for i in range(A)))
            else:
                f.write("".join("%.4f %.4f %.4f %d\n" % (r[i, 0], r[i, 1], r[i, 2], iso[i]) for i in range(A)))


def install(tables_dir, params):
    """write the files a run with these parameters needs, if absent (params: reference names)"""
    A = {int(params.get("Aproj", 0)), int(params.get("Atarg", 0))}
    if 16 in A:
        p = os.path.join(tables_dir, "oxygen_plaintext.dat")
        if not os.path.exists(p):
            write_oxygen(p, oxygen())
    if int(params.get("include_NN_correlation", 0)) == 1:
        if 197 in A:
            p = os.path.join(tables_dir, "au197-sw-full_3Bchains-conf1820.dat")
            if not os.path.exists(p):
                write_nncorr(p, au197(), 197, True)
        if 208 in A:
            p = os.path.join(tables_dir, "pb208-1.dat")
            if not os.path.exists(p):
                write_nncorr(p, pb208(), 208, False)
