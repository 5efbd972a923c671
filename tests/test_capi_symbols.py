"""CPU: the C-ABI library loads, exports every entry point include/supermc_b200.h declares, and refuses
to compute without a GPU (no silent CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "supermc_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(smc_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    import supermc_b200 as smc
    L = smc.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert L.smc_abi_version() == 3


def test_struct_layouts_match_header():
    import supermc_b200 as smc
    assert C.sizeof(smc.EventOut) == smc.capi.EVENT_OUT_DTYPE.itemsize == 440
    p = smc.capi.default_params()
    assert (p.which_mc_model, p.aproj, p.ecm, p.maxx, p.finalfactor, p.cc_fluctuation_model) == (7, 208, 5020.0, 15.0, 40.0, 6)


def test_no_cpu_fallback():
    import torch
    import supermc_b200 as smc
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(smc.SmcError) as ei:
        smc.Context(smc.capi.default_params())
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    """the product path may never import, link or call anything under oracle/"""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "supermc_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"smc_oracle|from oracle|import oracle|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
