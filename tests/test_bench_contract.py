"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the unmodified reference binary on
the host cores and prints ONE JSON line with the keys the driver reads; the algorithmic-work bookkeeping of the roofline
follows SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_prints_one_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "superMC_ref.e")):
        pytest.skip("oracle/_ref is not built here")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-events-per-process", "12"],
                         capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert out.returncode == 0 and len(lines) == 1, (out.stdout, out.stderr[-500:])
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "events/sec" and d["unit"] == "events/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "Pb+Pb" in d["config"]["workload"] and d["gpu_launches"] == 0
    # one process, the reference's own 8-process mode, one process per core
    modes = d["cpu_baseline"]["modes"]
    assert set(modes) == {"1_process", "8_process", "all_cores"} and modes["8_process"]["processes"] == 8
    assert all(m["median"] > 0 and m["min"] <= m["median"] for m in modes.values()) and d["cpu_baseline"]["cpu_model"] != ""


def test_algorithmic_work_follows_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    ev = np.zeros(1, dtype=[("npart1", "i4"), ("npart2", "i4"), ("ncoll", "i4"), ("nonzero_cells", "i4"), ("tries", "i4")])
    ev["npart1"], ev["npart2"], ev["ncoll"], ev["nonzero_cells"], ev["tries"] = 60, 57, 369, 10000, 2
    w, dx = 0.4941, 0.1                                   # Pb+Pb 2.76 TeV: n5 = 49.4, n4 = 39.5 cells
    f_dep, f_mom, f_smp = bench.algorithmic_flops(ev, w, dx)
    n5, n4 = 10 * w / dx, 8 * w / dx
    assert abs(f_dep - 2 * (117 * n4 * n4 + 369 * n5 * n5)) < 1e-6 and f_mom == 200.0 * 10000
    assert f_smp == 2 * (6.0 * 208 * 208 + 2 * 3.0 * 208 * 208)
    f_dep_kln, _, _ = bench.algorithmic_flops(ev, w, dx, kln=True)
    assert abs(f_dep_kln - 2 * 117 * n5 * n5) < 1e-6
