"""the reference's performance sweep (GJK::GPU::testing, examples/gpu/example.cu:258-383)
on this build, in small: same CSV header and row format as the reference writes (example.cu:281, 376), same column
layout as its published data file for the plotting script (data/data_32bit_4070, plotting/create_plots.py:21-30), every
run's distances equal to the single-thread CPU run's bit for bit.  The full sweep is scripts/sweep_csv.py
(profiles/r1c_data_32bit_b200.csv)."""
import importlib.util
import os
import re

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sweep_csv_formats_and_parity(tmp_path):
    spec = importlib.util.spec_from_file_location("sweep_csv", os.path.join(ROOT, "scripts", "sweep_csv.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = [(1000, 50), (1000, 200), (50, 500), (5000, 500)]
    out, plot = tmp_path / "sweep.csv", tmp_path / "sweep_plot.csv"
    rows, prow = mod.main(str(out), cases=cases, runs=2, plot_out=str(plot), check=True)
    lines = out.read_text().splitlines()
    assert lines[0] == "NumPolytopes,NumVertices,CPU_Time_ms,GPU_Time_ms"  # example.cu:281
    assert len(lines) == 1 + len(cases)
    for line, (n, nv) in zip(lines[1:], cases):
        m = re.fullmatch(r"(\d+),(\d+),(\d+\.\d{6}),(\d+\.\d{6})", line)  # "%d,%d,%.6f,%.6f" (example.cu:376)
        assert m and int(m.group(1)) == n and int(m.group(2)) == nv
        assert float(m.group(3)) > 0 and float(m.group(4)) > 0
    plines = plot.read_text().splitlines()
    assert plines[0] == "polytopes,Vertices,GPU_ms,CPU_ms"  # data/data_32bit_4070
    assert all(re.fullmatch(r"\d+,\d+,\d+\.\d{4},\d+\.\d{4}", ln) for ln in plines[1:])
    assert set(mod.PUBLISHED_4070) >= {(1000, 50), (50000, 500)}  # the published table the script prints beside its rows
