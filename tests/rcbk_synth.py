"""Synthetic rcBK uGD tables in the reference's file format (`javier/ft_rcbk_mv_qs02_*.dat`: 121 Y-bins x 101
rows `Y kT N_F N_A`, rcBKfunc.cpp:115-178).  The real tables are absent upstream, so
MC-KLN with sub_model 100/101 can only be exercised on these stand-ins."""
import os
import numpy as np

MAXY, MAXKT = 121, 101


def file_names(sub_model):
    if sub_model == 100:
        out = []
        for k in range(2, 61):
            lab = str(k // 10) if k % 10 == 0 else ("0%d" % k if k < 10 else str(k))
            out.append("ft_rcbk_mv_qs02_%s_ad.dat" % lab)
        return out, 0.1, 2
    return ["ft_rcbk_mv_qs02_0168_g1_119_%d.dat" % i for i in range(1, 31)], 0.168, 1


def make_tables(sub_model):
    """-> kt, N_F, N_A arrays [maxQ0][121][101] (smooth, positive, saturation-like)"""
    names, dq0, off = file_names(sub_model)
    nq = len(names)
    kt1 = 20.0 * (np.arange(MAXKT) / (MAXKT - 1.0)) ** 1.5 + 1e-3 * np.arange(MAXKT) / MAXKT
    Y = 0.1 * np.arange(MAXY)
    q02 = dq0 * (np.arange(nq) + off)
    kt = np.broadcast_to(kt1, (nq, MAXY, MAXKT)).copy()
    qs2 = q02[:, None, None] * np.exp(0.28 * Y)[None, :, None]
    na = (qs2 / (kt ** 2 + qs2)) ** 2 / (kt ** 2 + 0.05) * (1.0 + 0.1 * np.sin(3.0 * kt))
    nf = 0.5 * na
    return kt, nf, na


def write_files(directory, sub_model):
    names, _, _ = file_names(sub_model)
    kt, nf, na = make_tables(sub_model)
    os.makedirs(directory, exist_ok=True)
    Y = np.repeat(0.1 * np.arange(MAXY), MAXKT)
    for iq, name in enumerate(names):
        np.savetxt(os.path.join(directory, name), np.stack([Y, kt[iq].ravel(), nf[iq].ravel(), na[iq].ravel()], axis=1), fmt="%.10e")
    # what the reader gets back after the text round trip
    kt2 = np.array([np.loadtxt(os.path.join(directory, n))[:, 1] for n in names]).reshape(len(names), MAXY, MAXKT)
    na2 = np.array([np.loadtxt(os.path.join(directory, n))[:, 3] for n in names]).reshape(len(names), MAXY, MAXKT)
    return kt2, na2
