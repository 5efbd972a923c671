"""ctypes binding of include/supermc_b200.h (one-to-one; no logic lives here)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("SMC_LIB") or os.path.join(HERE, "libsupermc_b200.so")
dp = C.POINTER(C.c_double)

RUN_MOMENTS, RUN_KEEP_RHO, RUN_THICKNESS, RUN_RHO_BINARY, RUN_SPECTATORS, RUN_LISTS = 1, 2, 4, 8, 16, 32
GRID_RHO, GRID_TA1, GRID_TA2, GRID_RHO_BINARY, GRID_SPEC_A, GRID_SPEC_B = range(6)


class Params(C.Structure):
    _fields_ = [("which_mc_model", C.c_int), ("sub_model", C.c_int), ("lambda_", C.c_double),
                ("tmax", C.c_int), ("tmax_subdivision", C.c_int), ("alpha", C.c_double),
                ("aproj", C.c_int), ("atarg", C.c_int), ("proj_deformed", C.c_int), ("targ_deformed", C.c_int),
                ("include_nn_correlation", C.c_int), ("shape_of_nucleons", C.c_int),
                ("collision_criterion", C.c_int), ("shape_of_entropy", C.c_int), ("quark_width", C.c_double),
                ("gauss_nucl_width", C.c_double), ("ecm", C.c_double), ("bmin", C.c_double), ("bmax", C.c_double),
                ("npmin", C.c_int), ("npmax", C.c_int), ("cutdsdy", C.c_int),
                ("cutdsdy_lowerbound", C.c_double), ("cutdsdy_upperbound", C.c_double),
                ("randomseed", C.c_int64), ("finalfactor", C.c_double),
                ("ecc_from_order", C.c_int), ("ecc_to_order", C.c_int),
                ("maxx", C.c_double), ("maxy", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("cc_fluctuation_model", C.c_int), ("cc_fluctuation_gamma_theta", C.c_double),
                ("pt_order", C.c_int), ("gaussian_lambda", C.c_double), ("cc_fluctuation_k", C.c_double),
                ("ny", C.c_int), ("ymax", C.c_double), ("max_batch", C.c_int), ("ncoll_cap", C.c_int)]


class Constants(C.Structure):
    _fields_ = [("siginnn", C.c_double), ("siginnn200", C.c_double), ("width", C.c_double),
                ("sigma_gg", C.c_double), ("dsq", C.c_double), ("maxx_cells", C.c_int), ("maxy_cells", C.c_int),
                ("kln_dt", C.c_double), ("kln_tmax", C.c_int)]


class EventOut(C.Structure):
    _fields_ = [("b", C.c_double), ("npart1", C.c_int), ("npart2", C.c_int), ("ncoll", C.c_int),
                ("tries", C.c_int), ("nspec", C.c_int), ("status", C.c_int), ("dsdy", C.c_double),
                ("total", C.c_double), ("xc", C.c_double), ("yc", C.c_double), ("mom", (C.c_double * 5) * 9),
                ("rn0", C.c_double), ("nonzero_cells", C.c_int), ("reserved", C.c_int)]


class EventIn(C.Structure):
    _fields_ = [("b", C.c_double), ("na", C.c_int), ("nb", C.c_int), ("proj", dp), ("targ", dp),
                ("pair_uniform", dp), ("coll_weight", dp), ("n_coll_weight", C.c_int), ("use_given_weights", C.c_int),
                ("proj_extra", dp), ("targ_extra", dp)]


EVENT_OUT_DTYPE = np.dtype([("b", "f8"), ("npart1", "i4"), ("npart2", "i4"), ("ncoll", "i4"), ("tries", "i4"),
                            ("nspec", "i4"), ("status", "i4"), ("dsdy", "f8"), ("total", "f8"), ("xc", "f8"),
                            ("yc", "f8"), ("mom", "f8", (9, 5)), ("rn0", "f8"), ("nonzero_cells", "i4"), ("reserved", "i4")], align=True)


class Profile3dParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("neta", C.c_int), ("dx", C.c_double), ("dy", C.c_double), ("deta", C.c_double),
                ("ecm", C.c_double), ("random_flag", C.c_int), ("seed", C.c_int64)]


def profile3d(x, y, ids, nx=261, ny=261, neta=101, dx=0.1, dy=0.1, deta=0.1, ecm=19.6, random_flag=1, seed=1, eta=None, sigma3=None, device=0):
    """scripts/generate_3d_profiles: participants -> rho[neta][nx][ny]; returns (rho, eta_used, sigma3_used)"""
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64); ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = len(x)
    p = Profile3dParams(nx, ny, neta, dx, dy, deta, ecm, random_flag, seed)
    rho = np.zeros((neta, nx, ny)); eu = np.zeros(n); su = np.zeros((n, 3))
    e_in = None if eta is None else np.ascontiguousarray(eta, dtype=np.float64)
    s_in = None if sigma3 is None else np.ascontiguousarray(sigma3, dtype=np.float64)
    L = lib()
    L.smc_profile3d.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 8
    rc = L.smc_profile3d(int(device), C.byref(p), n, x.ctypes.data, y.ctypes.data, ids.ctypes.data, e_in.ctypes.data if e_in is not None else None,
                         s_in.ctypes.data if s_in is not None else None, rho.ctypes.data, eu.ctypes.data, su.ctypes.data)
    if rc != 0:
        raise SmcError("smc_profile3d: rc=%d" % rc)
    return rho, eu, su


class SmcError(RuntimeError):
    pass


_LIB = None


def build_library(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "csrc"), "-j8"],
                          stdout=None if verbose else subprocess.DEVNULL)
    return SO


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            raise SmcError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is the only compute path)" % SO)
        L = C.CDLL(SO)
        L.smc_last_error.restype = C.c_char_p
        L.smc_kernel_launches.restype = C.c_int64
        L.smc_last_run_ms.restype = C.c_double
        L.smc_measure_fp64_peak.restype = C.c_double
        L.smc_measure_hbm_write_peak.restype = C.c_double
        L.smc_run_events.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint, C.c_void_p]
        L.smc_run_from_positions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_void_p]
        L.smc_avg_run.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.smc_avg_run_from_positions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.smc_avg_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.smc_get_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.smc_centrality_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.smc_get_grids.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.smc_set_seed.argtypes = [C.c_void_p, C.c_int64]
        L.smc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.smc_comm_backend.restype = C.c_char_p
        L.smc_comm_backend.argtypes = [C.c_void_p]
        L.smc_comm_last_allreduce_ms.restype = C.c_double
        L.smc_comm_last_allreduce_ms.argtypes = [C.c_void_p]
        L.smc_comm_gather_doubles.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.smc_pinned_alloc.restype = C.c_void_p
        L.smc_pinned_alloc.argtypes = [C.c_size_t]
        L.smc_pinned_free.argtypes = [C.c_void_p]
        assert C.sizeof(EventOut) == EVENT_OUT_DTYPE.itemsize
        _LIB = L
    return _LIB


def default_params(**kw):
    p = Params()
    lib().smc_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Context:
    """smc_ctx wrapper: one per GPU."""

    def __init__(self, params=None, device=0, **kw):
        self.p = params if params is not None else default_params(**kw)
        self.h = C.c_void_p()
        rc = lib().smc_create(C.byref(self.p), int(device), C.byref(self.h))
        if rc != 0:
            msg = lib().smc_last_error(self.h).decode() if self.h else "smc_create failed"
            if self.h:
                lib().smc_destroy(self.h)
            self.h = None
            raise SmcError("smc_create: rc=%d: %s" % (rc, msg))
        self.k = Constants()
        self._ck(lib().smc_get_constants(self.h, C.byref(self.k)))
        self.G = self.k.maxx_cells * self.k.maxy_cells

    def _ck(self, rc):
        if rc != 0:
            raise SmcError("rc=%d: %s" % (rc, lib().smc_last_error(self.h).decode()))

    def close(self):
        if self.h:
            lib().smc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_quark_table(self, rows3):
        a = np.ascontiguousarray(rows3, dtype=np.float64)
        self._ck(lib().smc_load_quark_table(self.h, a.ctypes.data_as(dp), len(a)))

    def load_config_table(self, which, xyz):
        a = np.ascontiguousarray(xyz, dtype=np.float64)
        ncfg, A = a.shape[0], a.shape[1] // 3 if a.ndim == 2 else a.shape[1]
        self._ck(lib().smc_load_config_table(self.h, int(which), a.ctypes.data_as(dp), ncfg, A))

    def load_rcbk_tables(self, kt, na):
        kt = np.ascontiguousarray(kt, dtype=np.float64); na = np.ascontiguousarray(na, dtype=np.float64)
        q, y, k = kt.shape
        self._ck(lib().smc_load_rcbk_tables(self.h, kt.ctypes.data_as(dp), na.ctypes.data_as(dp), q, y, k))

    def build_kln_table(self):
        """-> (tmax, tmax), or (ny, tmax, tmax) with several rapidity slices"""
        ny = max(self.p.ny, 1)
        t = np.zeros((ny, self.k.kln_tmax, self.k.kln_tmax))
        self._ck(lib().smc_build_kln_table(self.h, t.ctypes.data_as(dp)))
        return t[0] if ny == 1 else t

    def set_kln_table(self, table, dt):
        t = np.ascontiguousarray(table, dtype=np.float64)      # (tmax, tmax) or (ny, tmax, tmax)
        assert t.ndim == 2 or t.shape[0] == max(self.p.ny, 1)
        self._ck(lib().smc_set_kln_table(self.h, t.ctypes.data_as(dp), t.shape[-1], C.c_double(dt)))

    def run_events(self, first_event_id, n, flags=RUN_MOMENTS, out=None):
        """-> structured array (EVENT_OUT_DTYPE) of n accepted events.  `out` may be a preallocated array."""
        if out is None:
            out = np.zeros(n * max(self.p.ny, 1), dtype=EVENT_OUT_DTYPE)      # ny > 1: row e*ny + iy
        self._ck(lib().smc_run_events(self.h, int(first_event_id), int(n), int(flags), out.ctypes.data))
        return out

    def _event_in_array(self, events):
        """events: list of dicts(b, proj (A,8), targ (B,8), pair_uniform (A,B)|None, coll_weight (n,2)|None, given_w)"""
        n = len(events)
        arr = (EventIn * n)()
        keep = []
        for i, ev in enumerate(events):
            pj = np.ascontiguousarray(ev["proj"], dtype=np.float64); tg = np.ascontiguousarray(ev["targ"], dtype=np.float64)
            keep += [pj, tg]
            arr[i].b = ev.get("b", 0.0); arr[i].na = len(pj); arr[i].nb = len(tg)
            arr[i].proj = pj.ctypes.data_as(dp); arr[i].targ = tg.ctypes.data_as(dp)
            pu = ev.get("pair_uniform")
            if pu is not None:
                pu = np.ascontiguousarray(pu, dtype=np.float64); keep.append(pu); arr[i].pair_uniform = pu.ctypes.data_as(dp)
            cw = ev.get("coll_weight")
            if cw is not None:
                cw = np.ascontiguousarray(cw, dtype=np.float64).reshape(-1, 2); keep.append(cw)
                arr[i].coll_weight = cw.ctypes.data_as(dp); arr[i].n_coll_weight = len(cw)
            arr[i].use_given_weights = int(ev.get("given_w", 1))
            for key in ("proj_extra", "targ_extra"):
                x = ev.get(key)
                if x is not None:
                    x = np.ascontiguousarray(x, dtype=np.float64)
                    if x.shape[1] < 20:      # rows of SMC_EXTRA_ROW: older fixtures carry 16 columns (no per-quark weights: 1/3 each)
                        x = np.concatenate([x, np.zeros((len(x), 20 - x.shape[1]))], axis=1); x[:, 15:18] = 1.0 / 3.0
                    x = np.ascontiguousarray(x); keep.append(x); setattr(arr[i], key, x.ctypes.data_as(dp))
        return arr, keep

    def run_from_positions(self, events, flags=RUN_MOMENTS):
        """events: list of dicts(b, proj (A,8), targ (B,8), pair_uniform (A,B)|None, coll_weight (n,2)|None, given_w)"""
        arr, keep = self._event_in_array(events)
        out = np.zeros(len(events) * max(self.p.ny, 1), dtype=EVENT_OUT_DTYPE)
        self._ck(lib().smc_run_from_positions(self.h, len(events), C.byref(arr), int(flags), out.ctypes.data))
        return out

    # ---- averaged profiles (operation 3) ----
    def avg_begin(self, from_order, to_order, with_rp=True, branches=3):
        self._ck(lib().smc_avg_begin(self.h, int(from_order), int(to_order), int(bool(with_rp)), int(branches)))

    def avg_run(self, first_event_id, n):
        out = np.zeros(n, dtype=EVENT_OUT_DTYPE)
        self._ck(lib().smc_avg_run(self.h, int(first_event_id), int(n), out.ctypes.data))
        return out

    def avg_run_from_positions(self, events):
        arr, keep = self._event_in_array(events)
        out = np.zeros(len(events), dtype=EVENT_OUT_DTYPE)
        self._ck(lib().smc_avg_run_from_positions(self.h, len(events), C.byref(arr), out.ctypes.data))
        return out

    def avg_get(self, order, variant, quantity, branch=0):
        g = np.zeros((self.k.maxx_cells, self.k.maxy_cells))
        self._ck(lib().smc_avg_get(self.h, int(order), int(variant), int(quantity), int(branch), g.ctypes.data))
        return g

    def avg_count(self):
        n = C.c_int64()
        self._ck(lib().smc_avg_count(self.h, C.byref(n)))
        return n.value

    def avg_set_count(self, n):
        self._ck(lib().smc_avg_set_count(self.h, C.c_int64(int(n))))

    def avg_device_buffer(self):
        """(device pointer, n_doubles) of the accumulator sums, for an all-reduce across GPUs"""
        p = C.c_void_p(); n = C.c_int64()
        self._ck(lib().smc_avg_device_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- several GPUs (one process each) ----
    def comm_init(self, rank, world, addr="127.0.0.1", port=29517):
        self._ck(lib().smc_comm_init(self.h, int(rank), int(world), addr.encode(), int(port)))
        return lib().smc_comm_backend(self.h).decode()

    def comm_barrier(self):
        self._ck(lib().smc_comm_barrier(self.h))

    def avg_allreduce(self):
        """-> device time of the all-reduce [ms]"""
        self._ck(lib().smc_avg_allreduce(self.h))
        return lib().smc_comm_last_allreduce_ms(self.h)

    def comm_gather(self, rows, world):
        """rank 0 receives every rank's rows (float64, any shape with a fixed trailing width) in rank order"""
        a = np.ascontiguousarray(rows, dtype=np.float64)
        cnt = np.zeros(world, dtype=np.int64)
        # two-phase: the sizes first, then the rows into a buffer of the right size on rank 0
        sizes = np.zeros(world)
        self._ck(lib().smc_comm_gather_doubles(self.h, np.array([float(a.size)]).ctypes.data, 1, sizes.ctypes.data, world, None))
        tot = int(sizes.sum()) if sizes.any() else a.size
        out = np.zeros(max(tot, 1))
        self._ck(lib().smc_comm_gather_doubles(self.h, a.ctypes.data, a.size, out.ctypes.data, out.size, cnt.ctypes.data))
        return out[:tot], cnt

    def set_seed(self, seed):
        self._ck(lib().smc_set_seed(self.h, int(seed)))

    @property
    def max_batch(self):
        return lib().smc_max_batch(self.h)

    def grids(self, first_slot, n, which):
        g = np.zeros((n, self.k.maxx_cells, self.k.maxy_cells))
        self._ck(lib().smc_get_grids(self.h, int(first_slot), int(n), int(which), g.ctypes.data))
        return g

    def grid(self, slot, which):
        g = np.zeros((self.k.maxx_cells, self.k.maxy_cells))
        self._ck(lib().smc_get_grid(self.h, int(slot), int(which), g.ctypes.data))
        return g

    def _rows(self, fn, slot, width, *extra):
        n = C.c_int()
        self._ck(fn(self.h, int(slot), *extra, None, C.byref(n)))
        a = np.zeros((max(n.value, 1), width))
        self._ck(fn(self.h, int(slot), *extra, a.ctypes.data_as(dp), C.byref(n)))
        return a[:n.value]

    def participants(self, slot):
        return self._rows(lib().smc_get_participants, slot, 8)

    def collisions(self, slot):
        return self._rows(lib().smc_get_collisions, slot, 6)

    def spectators(self, slot):
        return self._rows(lib().smc_get_spectators, slot, 3)

    def quarks(self, slot):
        return self._rows(lib().smc_get_quarks, slot, 6)

    def nucleons(self, slot, which):
        return self._rows(lib().smc_get_nucleons, slot, 8, int(which))

    def centrality_sort(self, key):
        k = np.ascontiguousarray(key, dtype=np.float64); perm = np.zeros(len(k), dtype=np.int64)
        self._ck(lib().smc_centrality_sort(self.h, k.ctypes.data, len(k), perm.ctypes.data))
        return perm

    def set_profiling(self, on=True):
        self._ck(lib().smc_set_profiling(self.h, int(bool(on))))

    def stage_ms(self):
        a = (C.c_double * 4)()
        self._ck(lib().smc_get_stage_ms(self.h, a))
        return dict(sample_collide=a[0], deposit=a[1], combine=a[2], moments=a[3])

    @property
    def launches(self):
        return lib().smc_kernel_launches(self.h)

    @property
    def last_run_ms(self):
        return lib().smc_last_run_ms(self.h)

    def fp64_peak_tflops(self):
        return lib().smc_measure_fp64_peak(self.h)

    def hbm_write_peak_gbs(self):
        return lib().smc_measure_hbm_write_peak(self.h)
