"""Reporting step: npbench_b200.report keeps npbench.db readable by the reference's
plot_results.py (results keeps the 13 columns of utilities.py:75-90) and adds the Gcell/s + roofline side table."""
import json
import os
import sqlite3

import pytest

from npbench_b200 import report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COLUMNS = ["id", "timestamp", "benchmark", "kind", "domain", "dwarf", "preset", "mode", "framework", "version",
               "details", "validated", "time"]


def _harness_row(conn, short, preset, framework, t, validated=1):
    # what Test.run writes (test.py:144-162)
    conn.execute("INSERT INTO results(timestamp, benchmark, kind, domain, dwarf, preset, mode, framework, version, details,"
                 " validated, time) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                 (1, short, "microbench", "Physics", "structured_grids", preset, "main", framework, "1.0", "default",
                  validated, t))


def test_roofline_side_table(tmp_path):
    db = str(tmp_path / "npbench.db")
    conn = sqlite3.connect(db)
    conn.execute(report.SQL_CREATE_RESULTS)
    for t in (0.5, 0.4, 0.6):
        _harness_row(conn, "heat3d", "L", "numpy", t)
    _harness_row(conn, "heat3d", "L", "b200", 0.3687e-3)
    _harness_row(conn, "gemm", "L", "numpy", 1.0)              # not ours: ignored
    conn.commit()
    assert report.refresh_roofline(conn, 6553.6) == 4
    cols = [r[1] for r in conn.execute("PRAGMA table_info(results)")]
    assert cols == REF_COLUMNS                                  # plot_results.py still sees the reference schema
    rows = {(b, p, f): (n, med, gc, frac) for b, p, f, n, med, gc, frac in report.summary(conn)}
    n, med, gc, frac = rows[("heat3d", "L", "b200")]
    units = 2 * 99 * 68 ** 3
    assert n == 1 and gc == pytest.approx(units / 0.3687e-3 / 1e9) and frac == pytest.approx(gc * 16 / 6553.6)
    assert rows[("heat3d", "L", "numpy")][0] == 3 and rows[("heat3d", "L", "numpy")][1] == 0.5
    conn.close()


def test_import_committed_bench_json(tmp_path):
    path = os.path.join(ROOT, "profiles", "r01_bench_n1.json")
    line = [ln for ln in open(path).read().splitlines() if ln.strip().startswith("{")][-1]
    bench = json.loads(line)
    conn = sqlite3.connect(str(tmp_path / "npbench.db"))
    conn.execute(report.SQL_CREATE_RESULTS)
    n = report.import_suite(conn, bench, timestamp=7)
    presets = [r for r in bench["suite"] if r["kernel"] in report.BENCH and r["preset"] in ("S", "M", "L", "paper")]
    assert n == len(presets) and n >= 20
    report.refresh_roofline(conn, bench["roofline"]["peak"])
    got = {(b, p): frac for b, p, f, _, _, _, frac in report.summary(conn) if f == "b200"}
    for r in presets:                                           # the side table reproduces bench.py's own fractions
        assert got[(report.BENCH[r["kernel"]]["short"], r["preset"])] == pytest.approx(r["frac_of_peak"], abs=2e-3)
    conn.close()


def test_cli(tmp_path, capsys):
    db = str(tmp_path / "x.db")
    report.main(["--db", db, "--bench-json", os.path.join(ROOT, "profiles", "r01_bench_n1.json")])
    out = capsys.readouterr().out
    assert "imported" in out and "heat3d" in out and "b200" in out


def test_channel_flow_step_counts_are_the_pinned_reference_values(pins_large):
    """report.py's unit count of channel_flow = steps until the reference's convergence loop stops
    (tests/golden/make_golden_large.py pins it from the unmodified reference at every preset)."""
    from npbench_b200 import report
    for preset, p in report.BENCH["channel_flow"]["presets"].items():
        assert p["steps"] == pins_large["channel_flow/" + preset]["stepcount"], preset
