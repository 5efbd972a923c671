"""Reader for the record stream written by oracle/ref_build/ref_dump.cpp (test infrastructure)."""
import struct
import numpy as np


def read_records(path):
    """-> dict name -> ndarray (float64), insertion-ordered."""
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    p = 0
    while p < len(buf):
        (nl,) = struct.unpack_from("<q", buf, p); p += 8
        name = buf[p:p + nl].decode(); p += nl
        (nd,) = struct.unpack_from("<q", buf, p); p += 8
        dims = struct.unpack_from("<%dq" % nd, buf, p); p += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(buf, dtype="<f8", count=n, offset=p).reshape(dims).copy()
        p += 8 * n
    return out


def group_tries(rec):
    """split 't<k>/...' keys into a list of per-try dicts (ordered by try id); returns (globals, tries)."""
    glob, tries = {}, {}
    for k, v in rec.items():
        if k[0] == "t" and "/" in k and k[1:k.index("/")].isdigit():
            t = int(k[1:k.index("/")])
            tries.setdefault(t, {})[k[k.index("/") + 1:]] = v
        else:
            glob[k] = v
    return glob, [tries[t] for t in sorted(tries)]
