"""supermc_b200/centrality.py against the unmodified reference scripts (tests/golden/centrality.npz, made by
tests/golden/make_centrality_golden.py): the centrality-table writer byte for byte, the centrality-window translation
exactly.  The -m gpu test replaces numpy's argsort by the device sort."""
import os
import numpy as np
import pytest

from helpers import GOLDEN
from supermc_b200 import centrality as cen

Z = np.load(os.path.join(GOLDEN, "centrality.npz"))


@pytest.mark.parametrize("cut", ["total_entropy", "Npart"])
def test_table_writer_equals_reference_script(cut):
    coll = Z["coll"]
    order = np.argsort(-coll[:, {"total_entropy": 3, "Npart": 1}[cut]])       # the reference's own ranking (centrality_cut_h5.py:50-55)
    assert np.array_equal(np.argsort(-cen.sort_key(coll, cut)), order)
    assert cen.centrality_table_text(coll, order, cut) == str(Z["table_" + cut])


def test_collision_data_columns():
    rows = np.arange(2 * 49, dtype=np.float64).reshape(2, 49)
    c = cen.collision_data(rows)
    assert c.dtype == np.float32 and np.array_equal(c[0], [48, 45, 46, 47, 47]) and np.array_equal(c[1], [97, 94, 95, 96, 96])


def test_window_translation_equals_reference_script():
    for case in Z["cases"]:
        key, name, cut = case[0], case[1], case[2]
        lo, up = float(case[3]), float(case[4])
        model, a, b, ecm, fl = int(float(case[5])), int(float(case[6])), int(float(case[7])), float(case[8]), int(float(case[9]))
        assert cen.table_file_name(cut, model, a, b, ecm, fl) == name
        p = cen.translate_centrality_cut(Z[key], lo, up, cut)
        assert p["cutdSdy"] == int(float(case[10]))
        if cut == "total_entropy":
            assert p["cutdSdy_lowerBound"] == float(case[11]) and p["cutdSdy_upperBound"] == float(case[12])
        assert (p["Npmin"], p["Npmax"], p["bmin"], p["bmax"]) == tuple(float(x) for x in case[13:17]), case


def test_wrapper_builds_the_reference_command_line(tmp_path, capsys):
    name = str(Z["cases"][0][1])
    np.savetxt(tmp_path / name, Z[str(Z["cases"][0][0])])
    rc = cen.main(["run", "--model", "MCGlb", "--ecm", "2760", "--collsys", "Pb", "Pb", "--cen", "0-5", "--tables", str(tmp_path),
                   "--nev", "10", "--dry-run", "average_to_order=2"])
    line = capsys.readouterr().out.strip()
    assert rc == 0 and "which_mc_model=5" in line and "cutdSdy=1" in line and "operation=3" in line and line.endswith("average_to_order=2")
    lo = float([t for t in line.split() if t.startswith("cutdSdy_lowerBound=")][0].split("=")[1])
    assert lo == float(Z["cases"][0][11])
    assert cen.main(["run", "--collsys", "Au", "Au", "--ecm", "200", "--tables", str(tmp_path), "--dry-run"]) == 1      # no table: the reference exits too


@pytest.mark.gpu
def test_table_from_the_device_sort():
    import supermc_b200 as smc
    coll = Z["coll"]
    ctx = smc.Context(smc.capi.default_params(max_batch=8))
    order = cen.device_order(ctx, coll, "total_entropy")
    ctx.close()
    key = coll[:, 3].astype(np.float64)
    assert np.array_equal(np.sort(order), np.arange(len(key))) and np.all(np.diff(key[order]) <= 0)
    # the 20000 dS/dy keys of the fixture are distinct, so the order -- and with it every byte of the table -- is unique
    assert len(np.unique(key)) == len(key)
    assert cen.centrality_table_text(coll, order, "total_entropy") == str(Z["table_total_entropy"])
