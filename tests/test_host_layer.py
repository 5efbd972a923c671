"""CPU: the C++ host layer (supermc_b200/host) -- the parameters.dat surface and the data/ text layouts --
against files the reference's own writers produced (tests/golden/text_formats.npz)."""
import ctypes as C
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    so = os.path.join(ROOT, "supermc_b200", "libsupermc_host.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(so)


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(ROOT, "tests", "golden", "text_formats.npz"))


def _text(fx, name):
    return fx["file/" + name].tobytes().decode()


def test_parameter_reader_semantics(host):
    txt = "# comment\nwhich_mc_model = 5        # (1) MC-CGC\n  Aproj=208\nEcm = 2760.   # GeV\n\nnev = 50000\n"
    v = C.c_double()
    get = lambda name, over="": (host.smc_host_param(txt.encode(), over.encode(), name.encode(), C.byref(v)), v.value)
    assert get("which_mc_model") == (0, 5.0)
    assert get("APROJ") == (0, 208.0)                  # names are case-insensitive (ParameterReader.cpp:61-69)
    assert get("ecm") == (0, 2760.0)
    assert get("nev", "nev=10 finalFactor=1") == (0, 10.0)      # argv overrides the file
    assert get("finalfactor", "nev=10 finalFactor=1") == (0, 1.0)
    assert get("not_there")[0] == 1                    # missing name is an error (reference: exit(-1))
    assert host.smc_host_param(b"bad line without equals\n", b"", b"x", C.byref(v)) == 1


def _event_out(row, smc):
    ev = smc.EventOut()
    for n in range(9):
        for k in range(5):
            ev.mom[n][k] = row[5 * n + k]
    npart, ncoll = int(row[45]), int(row[46])
    ev.npart1, ev.npart2, ev.ncoll, ev.total, ev.b = npart, 0, ncoll, row[47], row[48]
    return ev


def test_ecc_rows_byte_identical(host, fx):
    import supermc_b200 as smc
    buf = C.create_string_buffer(4096)
    for order in list(range(1, 10)) + [10]:
        lines = _text(fx, "h_ecc_%d.dat" % order).splitlines(keepends=True)
        assert len(lines) == len(fx["ecc_rows"]) == 2
        for i, row in enumerate(fx["ecc_rows"]):
            n = host.smc_host_format_ecc_row(C.byref(_event_out(row, smc)), order, 0, buf, 4096)
            assert n > 0 and buf.value.decode() == lines[i], (order, i)


def test_grid_writers_byte_identical(host, fx):
    last = max(int(k[1:k.index("/")]) for k in fx.files if k.endswith("/rho"))
    rho = np.ascontiguousarray(fx["t%d/rho" % last] * float(fx["consts"][6]))
    Maxx, Maxy = rho.shape
    buf = C.create_string_buffer(1 << 20)
    n = host.smc_host_format_block(rho.ctypes.data_as(C.POINTER(C.c_double)), Maxx, Maxy, buf, 1 << 20)
    assert n > 0 and buf.value.decode() == _text(fx, "ref_block.dat")
    hdr = fx["t%d/hdr" % last]
    host.smc_host_format_4col.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 6 + [C.c_char_p, C.c_int]
    n = host.smc_host_format_4col(rho.ctypes.data, Maxx, Maxy, -8.0, -8.0, 0.5, 0.5, 0.0, float(hdr[2] + hdr[3]), buf, 1 << 20)
    assert n > 0 and buf.value.decode() == _text(fx, "ref_4col.dat")


def test_list_writers_byte_identical(host, fx):
    buf = C.create_string_buffer(1 << 20)
    host.smc_host_format_list.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    tries = sorted({k[:k.index("/")] for k in fx.files if k.startswith("t") and "/" in k}, key=lambda s: int(s[1:]))
    acc = [t for t in tries if int(fx[t + "/hdr"][4])]
    for ie, t in enumerate(acc):
        proj, targ = fx[t + "/proj"], fx[t + "/targ"]
        pp, tp = fx[t + "/proj_part"].astype(int), fx[t + "/targ_part"].astype(int)
        rows = np.zeros((len(pp) + len(tp), 8))
        rows[:len(pp), 0:2] = proj[pp, 0:2]; rows[:len(pp), 2] = 1
        rows[len(pp):, 0:2] = targ[tp, 0:2]; rows[len(pp):, 2] = 2
        host.smc_host_format_list(1, rows.ctypes.data, len(rows), 8, buf, 1 << 20)
        assert buf.value.decode() == _text(fx, "ref_participants_%d.dat" % ie)
        coll = np.ascontiguousarray(fx[t + "/coll"])
        host.smc_host_format_list(0, coll.ctypes.data, len(coll), 6, buf, 1 << 20)
        assert buf.value.decode() == _text(fx, "ref_binary_%d.dat" % ie)
        sp = np.ascontiguousarray(fx[t + "/spectators"])
        host.smc_host_format_list(2, sp.ctypes.data, len(sp), 3, buf, 1 << 20)
        assert buf.value.decode() == _text(fx, "Spectators_event_%d.dat" % (1000 + ie))
    # quarks.data (appended by every dumpBinaryTable call, MCnucl.cpp:1198-1201): three valence quarks per wounded nucleon,
    # position at precision 3, Quark::getBoundingBox at the default precision 6
    want = ""
    for t in acc:
        rows = []
        for side, part in (("/proj", "/proj_part"), ("/targ", "/targ_part")):
            nuc, ex = fx[t + side], fx[t + side + "_x"]
            for i in fx[t + part].astype(int):
                for q in range(3):
                    qx, qy = ex[i, 4 + 3 * q], ex[i, 5 + 3 * q]
                    X, Y = qx + nuc[i, 0], qy + nuc[i, 1]
                    rows.append([X, Y, (qx - 1.2) + (X - qx), (qx + 1.2) + (X - qx), (qy - 1.2) + (Y - qy), (qy + 1.2) + (Y - qy)])
        rows = np.ascontiguousarray(rows)
        host.smc_host_format_list(3, rows.ctypes.data, len(rows), 6, buf, 1 << 20)
        want += buf.value.decode()
    assert want == _text(fx, "quarks.data")
    # nucl1.data / nucl2.data hold the last event (rewritten every event, MCnucl.cpp:1203-1209)
    t = acc[-1]
    for f, key in (("nucl1.data", "/proj"), ("nucl2.data", "/targ")):
        nu = np.ascontiguousarray(fx[t + key][:, :2])
        host.smc_host_format_list(0, nu.ctypes.data, len(nu), 2, buf, 1 << 20)
        assert buf.value.decode() == _text(fx, f)
