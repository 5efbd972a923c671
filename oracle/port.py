"""ctypes binding of oracle/libsmc_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class Cfg(C.Structure):
    _fields_ = [("Maxx", C.c_int), ("Maxy", C.c_int), ("Xmin", C.c_double), ("Ymin", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("width", C.c_double), ("dsq", C.c_double),
                ("siginNN", C.c_double), ("sigma_gg", C.c_double), ("alpha", C.c_double),
                ("shape_of_nucleons", C.c_int), ("shape_of_entropy", C.c_int), ("collision_criterion", C.c_int),
                ("which_mc_model", C.c_int), ("sub_model", C.c_int), ("cc_fluct_model", C.c_int)]


class Nucleus(C.Structure):
    _fields_ = [("A", C.c_int), ("rad", C.c_double), ("dr", C.c_double), ("rmaxCut", C.c_double),
                ("rwMax", C.c_double), ("beta2", C.c_double), ("beta4", C.c_double), ("deformed", C.c_int),
                ("width", C.c_double), ("quark_width", C.c_double), ("quark_R", C.c_double),
                ("quark_table", dp), ("quark_rows", C.c_int)]


class Rand48(C.Structure):
    _fields_ = [("x", C.c_uint64)]


class Kln(C.Structure):
    _fields_ = [("ecm", C.c_double), ("lambda_", C.c_double), ("siginNN200", C.c_double),
                ("model", C.c_int), ("pt_order", C.c_int)]


UFN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libsmc_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libsmc_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.smc_o_sigma_inel.restype = C.c_double; L.smc_o_sigma_inel.argtypes = [C.c_double]
        L.smc_o_drand48.restype = C.c_double
        L.smc_o_six_point.restype = C.c_double; L.smc_o_six_point.argtypes = [C.c_double] * 8
        L.smc_o_density.restype = C.c_double
        L.smc_o_density_quarks.restype = C.c_double
        L.smc_o_density_kln.restype = C.c_double
        L.smc_o_kln_integrand.restype = C.c_double
        L.smc_o_kln_dndy.restype = C.c_double
        L.smc_o_rcbk_func.restype = C.c_double
        L.smc_o_rcbk_dndy.restype = C.c_double
        L.smc_o_populate.restype = C.c_long
        L.smc_o_populate_table.restype = C.c_long
        L.smc_o_populate_deuteron.restype = C.c_long
        L.smc_o_hulthen_inv_cdf.restype = C.c_double
        L.smc_o_hulthen_inv_cdf.argtypes = [C.c_double]
        L.smc_o_uniform_rand48.restype = C.c_double
        L.smc_o_uniform_philox.restype = C.c_double
        L.smc_o_uniform_philox.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(dp)


def sigma_inel(ecm):
    return lib().smc_o_sigma_inel(float(ecm))


def gauss_params(shape, siginNN, lam=4.14, user_w=0.812):
    w, s = C.c_double(), C.c_double()
    lib().smc_o_gauss_params(int(shape), C.c_double(siginNN), C.c_double(lam), C.c_double(user_w), C.byref(w), C.byref(s))
    return w.value, s.value


def make_cfg(maxx=13.0, maxy=13.0, dx=0.1, dy=0.1, ecm=2760.0, alpha=0.118, shape_of_nucleons=2,
             shape_of_entropy=2, collision_criterion=2, which_mc_model=5, sub_model=1, cc_fluct_model=6,
             gauss_nucl_width=0.812, gaussian_lambda=4.14):
    c = Cfg()
    c.Xmin, c.Ymin, c.dx, c.dy = -maxx, -maxy, dx, dy
    c.Maxx = int((2 * maxx) / dx + 0.1) + 1
    c.Maxy = int((2 * maxy) / dy + 0.1) + 1
    c.siginNN = sigma_inel(ecm)
    c.width, c.sigma_gg = gauss_params(shape_of_nucleons, c.siginNN, lam=gaussian_lambda, user_w=gauss_nucl_width)
    c.dsq = 0.1 * c.siginNN / np.pi
    c.alpha = alpha
    c.shape_of_nucleons, c.shape_of_entropy, c.collision_criterion = shape_of_nucleons, shape_of_entropy, collision_criterion
    c.which_mc_model, c.sub_model, c.cc_fluct_model = which_mc_model, sub_model, cc_fluct_model
    return c


class Stream48:
    """drand48 clone usable as a sequential uniform source."""
    def __init__(self, seed=None, state=None):
        self.s = Rand48()
        if state is not None:
            lib().smc_o_seed48(C.byref(self.s), C.c_ushort(int(state[0])), C.c_ushort(int(state[1])), C.c_ushort(int(state[2])))
        else:
            lib().smc_o_srand48(C.byref(self.s), C.c_long(int(seed)))

    def next(self):
        return lib().smc_o_drand48(C.byref(self.s))

    def args(self):
        return C.cast(lib().smc_o_uniform_rand48, C.c_void_p), C.byref(self.s)


class PhiloxStream(C.Structure):
    _fields_ = [("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32), ("event", C.c_uint64), ("tr", C.c_uint32), ("nuc", C.c_int)]


class StreamPhilox:
    """the CUDA path's counter-based event stream, restated in the oracle (uniform source for the port)"""
    def __init__(self, seed, event, tr, nuc):
        self.s = PhiloxStream(seed & 0xffffffff, (seed >> 32) & 0xffffffff, event, tr, nuc)

    def u(self, kind, cand, slot):
        return lib().smc_o_uniform_philox(C.byref(self.s), kind, cand, slot)

    def args(self):
        return C.cast(lib().smc_o_uniform_philox, C.c_void_p), C.byref(self.s)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().smc_o_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def nucleus(A, width, deformed=0, quark_width=0.3, quark_table=None):
    n = Nucleus()
    qt = None if quark_table is None else np.ascontiguousarray(quark_table, dtype=np.float64)
    lib().smc_o_nucleus_init(C.byref(n), int(A), int(deformed), C.c_double(width), C.c_double(quark_width),
                             _d(qt) if qt is not None else None, 0 if qt is None else len(qt))
    n._keep = qt
    return n


def populate(n, xc, yc, stream=None, ufn=None):
    out = np.zeros((max(n.A, 1), 7))
    cxphi = np.zeros(2)
    if ufn is not None:
        cb = UFN(ufn)
        lib().smc_o_populate(C.byref(n), C.c_double(xc), C.c_double(yc), cb, None, _d(out), _d(cxphi))
    else:
        f, st = stream.args()
        lib().smc_o_populate(C.byref(n), C.c_double(xc), C.c_double(yc), f, st, _d(out), _d(cxphi))
    return out, cxphi


def populate_table(n, cfg3A, recentre, redraw, xc, yc, stream=None, ufn=None):
    out = np.zeros((n.A, 7))
    cfg3A = np.ascontiguousarray(cfg3A, dtype=np.float64)
    if ufn is not None:
        cb = UFN(ufn)
        lib().smc_o_populate_table(C.byref(n), _d(cfg3A), int(recentre), int(redraw), C.c_double(xc), C.c_double(yc), cb, None, _d(out))
    else:
        f, st = stream.args()
        lib().smc_o_populate_table(C.byref(n), _d(cfg3A), int(recentre), int(redraw), C.c_double(xc), C.c_double(yc), f, st, _d(out))
    return out


def populate_deuteron(n, xc, yc, stream):
    out = np.zeros((2, 7))
    f, st = stream.args()
    lib().smc_o_populate_deuteron(C.byref(n), C.c_double(xc), C.c_double(yc), f, st, _d(out))
    return out


def collide(cfg, proj7, targ7, stream=None, u_in=None, want_u=False):
    """-> dict(ncoll, ncollA, ncollB, firsthitB, pairs, u, tested)"""
    proj7 = np.ascontiguousarray(proj7, dtype=np.float64); targ7 = np.ascontiguousarray(targ7, dtype=np.float64)
    A, B = len(proj7), len(targ7)
    ncA = np.zeros(A, dtype=np.int32); ncB = np.zeros(B, dtype=np.int32); fh = np.zeros(B, dtype=np.int32)
    maxp = A * B
    pairs = np.zeros((maxp, 2), dtype=np.int32)
    u = np.zeros((A, B)) if want_u else None
    tested = C.c_long()
    if stream is not None:
        f, st = stream.args()
    else:
        f, st = None, None
    uin = None if u_in is None else np.ascontiguousarray(u_in, dtype=np.float64)
    n = lib().smc_o_collide(C.byref(cfg), A, _d(proj7), B, _d(targ7), f, st,
                            _d(uin) if uin is not None else None, _d(u) if u is not None else None,
                            ncA.ctypes.data_as(ip), ncB.ctypes.data_as(ip), fh.ctypes.data_as(ip),
                            pairs.ctypes.data_as(ip), maxp, C.byref(tested))
    return dict(ncoll=n, ncollA=ncA, ncollB=ncB, firsthitB=fh, pairs=pairs[:n].copy(), u=u, tested=tested.value)


def _src8(a):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 8)
    return a


def thickness(cfg, src8):
    s = _src8(src8); g = np.zeros((cfg.Maxx, cfg.Maxy))
    lib().smc_o_thickness(C.byref(cfg), len(s), _d(s), _d(g))
    return g


def unit_gauss(cfg, src8):
    s = _src8(src8); g = np.zeros((cfg.Maxx, cfg.Maxy))
    lib().smc_o_unit_gauss(C.byref(cfg), len(s), _d(s), _d(g))
    return g


def density(cfg, proj8, targ8, coll8):
    p, t, c = _src8(proj8), _src8(targ8), _src8(coll8)
    g = np.zeros((cfg.Maxx, cfg.Maxy))
    dndy = lib().smc_o_density(C.byref(cfg), len(p), _d(p), len(t), _d(t), len(c), _d(c), _d(g))
    return g, dndy


def populate_q(n, xc, yc, stream):
    """populate() + the valence-quark offsets (qx0 qy0 qx1 qy1 qx2 qy2 per nucleon, sorted order)"""
    q = np.zeros((max(n.A, 1), 6))
    lib().smc_o_quark_out(_d(q))
    try:
        rows, _ = populate(n, xc, yc, stream=stream)
    finally:
        lib().smc_o_quark_out(None)
    return rows, q


def collide_quarks(cfg, proj7, qA, targ7, qB, quark_width, stream=None, u_in=None, want_u=False):
    """collision_criterion 3: GaussianNucleonsCal::testFluctuatedCollision on the quark offsets"""
    qA = np.ascontiguousarray(qA, dtype=np.float64); qB = np.ascontiguousarray(qB, dtype=np.float64)
    lib().smc_o_quark_collide(_d(qA), _d(qB), C.c_double(quark_width))
    try:
        return collide(cfg, proj7, targ7, stream=stream, u_in=u_in, want_u=want_u)
    finally:
        lib().smc_o_quark_collide(None, None, C.c_double(0.0))


def density_quarks(cfg, proj8, qP, fP, targ8, qT, fT, coll8, quark_width):
    """shape_of_entropy 3: three quark Gaussians per wounded nucleon (offsets q*, weights f*) + the binary term"""
    p, t, c = _src8(proj8), _src8(targ8), _src8(coll8)
    qP = np.ascontiguousarray(qP, dtype=np.float64).reshape(-1, 6); qT = np.ascontiguousarray(qT, dtype=np.float64).reshape(-1, 6)
    fP = np.ascontiguousarray(fP, dtype=np.float64).reshape(-1, 3); fT = np.ascontiguousarray(fT, dtype=np.float64).reshape(-1, 3)
    g = np.zeros((cfg.Maxx, cfg.Maxy))
    dndy = lib().smc_o_density_quarks(C.byref(cfg), len(p), _d(p), _d(qP), _d(fP), len(t), _d(t), _d(qT), _d(fT), len(c), _d(c), C.c_double(quark_width), _d(g))
    return g, dndy


def profile3d(nx, ny, neta, dx, dy, deta, src7):
    """profile_3d::generate_3d_profile with the rapidities and widths given (rows x y id eta sx sy se)"""
    s = np.ascontiguousarray(src7, dtype=np.float64).reshape(-1, 7)
    rho = np.zeros((neta, nx, ny))
    lib().smc_o_profile3d(int(nx), int(ny), int(neta), C.c_double(dx), C.c_double(dy), C.c_double(deta), len(s), _d(s), _d(rho))
    return rho


def density_kln(cfg, TA1, TA2, table, dT):
    table = np.ascontiguousarray(table, dtype=np.float64)
    TA1 = np.ascontiguousarray(TA1); TA2 = np.ascontiguousarray(TA2)
    g = np.zeros((cfg.Maxx, cfg.Maxy))
    dndy = lib().smc_o_density_kln(C.byref(cfg), _d(TA1), _d(TA2), _d(table), table.shape[0], C.c_double(dT), _d(g))
    return g, dndy


def cm_angle(cfg, dens, n):
    dens = np.ascontiguousarray(dens, dtype=np.float64); o = np.zeros(4)
    lib().smc_o_cm_angle(C.byref(cfg), _d(dens), int(n), _d(o))
    return o


def eccentricities(cfg, dens, boxes4, from_order=1, to_order=9):
    """-> dict(mom (9,5) rows [Re e_n, Im e_n, Re e'_n, Im e'_n, <r^n>] for n=1..9, rn, total, xc, yc)"""
    dens = np.ascontiguousarray(dens, dtype=np.float64)
    b = np.ascontiguousarray(boxes4, dtype=np.float64).reshape(-1, 4)
    o = np.zeros(53)
    lib().smc_o_eccentricities(C.byref(cfg), _d(dens), len(b), _d(b), int(from_order), int(to_order), _d(o))
    mom = np.stack([o[0:10], o[10:20], o[20:30], o[30:40], o[40:50]], axis=1)
    return dict(mom=mom[1:10], rn=o[40:50].copy(), total=o[50], xc=o[51], yc=o[52])


def kln(ecm, lam, model=7, pt_order=1):
    k = Kln(); k.ecm = ecm; k.lambda_ = lam; k.siginNN200 = sigma_inel(200.0); k.model = model; k.pt_order = pt_order
    return k


class Rcbk(C.Structure):
    _fields_ = [("set", C.c_int), ("maxQ0", C.c_int), ("maxY", C.c_int), ("maxKt", C.c_int), ("dQ0", C.c_double),
                ("kt", dp), ("na", dp), ("y2", dp)]


def rcbk(sub_model, kt, na):
    """tabulated rcBK uGD: kt, na arrays [maxQ0][maxY][maxKt]; second derivatives of the natural spline per (iq, iy)"""
    kt = np.ascontiguousarray(kt, dtype=np.float64); na = np.ascontiguousarray(na, dtype=np.float64)
    y2 = np.zeros_like(na)
    nq, ny, nk = kt.shape
    L = lib()
    for iq in range(nq):
        for iy in range(ny):
            L.smc_o_spline_natural(_d(kt[iq, iy]), _d(na[iq, iy]), nk, _d(y2[iq, iy]))
    t = Rcbk(); t.set = int(sub_model); t.maxQ0, t.maxY, t.maxKt = nq, ny, nk
    t.dQ0 = 0.1 if sub_model == 100 else 0.168
    t.kt, t.na, t.y2 = _d(kt), _d(na), _d(y2)
    t._keep = (kt, na, y2)
    return t


def rcbk_func(t, qs2, x, kt2, alp):
    return lib().smc_o_rcbk_func(C.byref(t), C.c_double(qs2), C.c_double(x), C.c_double(kt2), C.c_double(alp))


def rcbk_dndy(k, t, y, ta, tb, npt=400, nkt=200, nphi=64):
    return lib().smc_o_rcbk_dndy(C.byref(k), C.byref(t), C.c_double(y), C.c_double(ta), C.c_double(tb), npt, nkt, nphi)


def kln_dndy(k, y, ta, tb, npt=400, nkt=200, nphi=64):
    return lib().smc_o_kln_dndy(C.byref(k), C.c_double(y), C.c_double(ta), C.c_double(tb), npt, nkt, nphi)


def kln_integrand(k, y, ta, tb, x3):
    x = (C.c_double * 3)(*x3)
    return lib().smc_o_kln_integrand(C.byref(k), C.c_double(y), C.c_double(ta), C.c_double(tb), x)


# ---- NBD multiplicity fluctuations (MCnucl.cpp:868-905, NBD.cpp, RandomVariable.cpp) ----
def nbd_rand(p, r, stream):
    """NBD::rand(p, r) restated literally, consuming `stream` (Stream48) like the reference consumes drand48"""
    L = lib(); L.smc_o_nbd_rand.restype = C.c_long; L.smc_o_nbd_rand.argtypes = [C.c_double, C.c_double, C.c_void_p]
    return L.smc_o_nbd_rand(p, r, C.byref(stream.s))


def nbd_law(p, r, cap=8192):
    """-> (k0, probabilities): the closed-form law of NBD::rand (a truncated NBD with fractional end cells)"""
    L = lib(); L.smc_o_nbd_law.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_long), C.POINTER(C.c_double), C.c_int]
    k0 = C.c_long(); w = (C.c_double * cap)()
    n = L.smc_o_nbd_law(p, r, C.byref(k0), w, cap)
    pr = np.array(w[:n]); return k0.value, pr / pr.sum()


def nbd_quantile(p, r, u):
    L = lib(); L.smc_o_nbd_quantile.restype = C.c_long; L.smc_o_nbd_quantile.argtypes = [C.c_double, C.c_double, C.c_double]
    return L.smc_o_nbd_quantile(p, r, u)


def fluctuate_density(cfg, model, cc_k, rho, u, TA1=None, TA2=None):
    """MCnucl::fluctuateCurrentDensity with the per-cell uniforms given (row-major like the grids)"""
    rho = np.array(rho, dtype=np.float64, order="C"); u = np.ascontiguousarray(u, dtype=np.float64)
    z = np.zeros_like(rho)
    a = np.ascontiguousarray(TA1 if TA1 is not None else z, dtype=np.float64); b = np.ascontiguousarray(TA2 if TA2 is not None else z, dtype=np.float64)
    L = lib(); L.smc_o_fluctuate_density.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.smc_o_fluctuate_density(C.byref(cfg), int(model), C.c_double(cc_k), a.ctypes.data, b.ctypes.data, u.ctypes.data, rho.ctypes.data)
    return rho


def cell_uniforms(seed, event, npass, ncell):
    """the CUDA path's per-cell NBD uniforms (smc_philox.h smc_uniform_cell): Philox(ctr = event lo/hi, pass<<8 | 10<<1, cell)"""
    out = np.empty(ncell)
    key = [seed & 0xffffffff, (seed >> 32) & 0xffffffff]
    c2 = ((npass << 8) | (10 << 1)) & 0xffffffff
    for q in range(ncell):
        o = philox([event & 0xffffffff, (event >> 32) & 0xffffffff, c2, q], key)
        out[q] = float((o[0] << 21) | (o[1] >> 11)) / 9007199254740992.0
    return out
