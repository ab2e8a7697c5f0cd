"""Reporting step after the hot path.

The unmodified harness already writes one row per timed repetition into `npbench.db`
(table `results`, schema npbench/infrastructure/utilities.py:75-98, rows written by test.py:144-162;
`benchmark` is the bench_info short_name, `time` is seconds).  plot_results.py:90-168 reads that table
with `SELECT * FROM results`, so it must keep exactly its 13 columns.  This module therefore

  * leaves `results` untouched, except for optionally IMPORTING the device-timed suite of a bench.py JSON
    line as rows of framework "b200", details "cuda-events" (same columns, `time` in seconds), and
  * derives a side table `b200_roofline` with the metric of this repository for every row of the
    structured-grid kernels it knows (any framework, so the NumPy rows get their Gcell/s too):
    units of work, Gcell-updates/s, algorithmic bytes per unit, GB/s and the fraction of the HBM peak.

    python -m npbench_b200.report --db npbench.db [--bench-json profiles/r01_bench_n1.json] [--peak-gbs 6553.6]

Pure host-side Python (sqlite3 + json); no GPU and no oracle involved.
"""
import argparse
import json
import sqlite3
import time

# bench_info/<b>.json: short_name, kind, domain, dwarf, parameters (presets)
BENCH = {
    "jacobi_2d": dict(short="jacobi2d", kind="microbench", domain="Physics",
                      presets={"S": dict(TSTEPS=50, N=150), "M": dict(TSTEPS=80, N=350),
                               "L": dict(TSTEPS=200, N=700), "paper": dict(TSTEPS=1000, N=2800)}),
    "heat_3d": dict(short="heat3d", kind="microbench", domain="Physics",
                    presets={"S": dict(TSTEPS=25, N=25), "M": dict(TSTEPS=50, N=40),
                             "L": dict(TSTEPS=100, N=70), "paper": dict(TSTEPS=500, N=120)}),
    "fdtd_2d": dict(short="fdtd_2d", kind="microbench", domain="Physics",
                    presets={"S": dict(TMAX=20, NX=200, NY=220), "M": dict(TMAX=60, NX=400, NY=450),
                             "L": dict(TMAX=150, NX=800, NY=900), "paper": dict(TMAX=500, NX=1000, NY=1200)}),
    "hdiff": dict(short="hdiff", kind="microapp", domain="Weather",
                  presets={"S": dict(I=64, J=64, K=60), "M": dict(I=128, J=128, K=160),
                           "L": dict(I=384, J=384, K=160), "paper": dict(I=256, J=256, K=160)}),
    "vadv": dict(short="vadv", kind="microapp", domain="Weather",
                 presets={"S": dict(I=60, J=60, K=40), "M": dict(I=112, J=112, K=80),
                          "L": dict(I=180, J=180, K=160), "paper": dict(I=256, J=256, K=160)}),
    "jacobi_1d": dict(short="jacobi1d", kind="microbench", domain="Physics",
                      presets={"S": dict(TSTEPS=800, N=3200), "M": dict(TSTEPS=3000, N=12000),
                               "L": dict(TSTEPS=8500, N=34000), "paper": dict(TSTEPS=4000, N=32000)}),
    "seidel_2d": dict(short="seidel2d", kind="microbench", domain="Solver",
                      presets={"S": dict(TSTEPS=8, N=50), "M": dict(TSTEPS=15, N=100),
                               "L": dict(TSTEPS=40, N=200), "paper": dict(TSTEPS=100, N=400)}),
    "cavity_flow": dict(short="cavtflow", kind="microapp", domain="Physics",
                        presets={"S": dict(ny=61, nx=61, nt=25, nit=5), "M": dict(ny=121, nx=121, nt=50, nit=10),
                                 "L": dict(ny=201, nx=201, nt=100, nit=20), "paper": dict(ny=101, nx=101, nt=700, nit=50)}),
    # channel_flow's unit count depends on the data (steps until convergence); the reference values at the presets
    # (pinned from the unmodified reference in tests/golden/pins_large.json; tests/test_report.py checks them)
    "channel_flow": dict(short="chanflow", kind="microapp", domain="Physics",
                         presets={"S": dict(ny=61, nx=61, nit=5, steps=982), "M": dict(ny=121, nx=121, nit=10, steps=991),
                                  "L": dict(ny=201, nx=201, nit=20, steps=995), "paper": dict(ny=101, nx=101, nit=50, steps=989)}),
    "adi": dict(short="adi", kind="microbench", domain="Solver",
                presets={"S": dict(TSTEPS=5, N=100), "M": dict(TSTEPS=20, N=200),
                         "L": dict(TSTEPS=50, N=500), "paper": dict(TSTEPS=100, N=200)}),
}
SHORT2NAME = {v["short"]: k for k, v in BENCH.items()}

SQL_CREATE_RESULTS = """
CREATE TABLE IF NOT EXISTS results (
    id integer PRIMARY KEY,
    timestamp integer NOT NULL,
    benchmark text NOT NULL,
    kind text,
    domain text,
    dwarf text,
    preset text NOT NULL,
    mode text NOT NULL,
    framework text NOT NULL,
    version text NOT NULL,
    details text,
    validated integer,
    time real
);
"""   # identical to utilities.py:75-90

SQL_CREATE_ROOFLINE = """
CREATE TABLE IF NOT EXISTS b200_roofline (
    result_id integer PRIMARY KEY,
    benchmark text NOT NULL,
    preset text NOT NULL,
    framework text NOT NULL,
    details text,
    time real,
    units real,
    gcell_s real,
    bytes_per_unit real,
    gbps_algorithmic real,
    frac_of_hbm_peak real
);
"""


def units_and_bytes(bench: str, p: dict):
    """Unit of work and algorithmic bytes per unit."""
    if bench == "jacobi_2d":
        return 2.0 * (p["TSTEPS"] - 1) * (p["N"] - 2) ** 2, 16.0
    if bench == "heat_3d":
        return 2.0 * (p["TSTEPS"] - 1) * (p["N"] - 2) ** 3, 16.0
    if bench == "fdtd_2d":
        return float(p["TMAX"] * p["NX"] * p["NY"]), 48.0
    if bench == "hdiff":
        I, J, K = p["I"], p["J"], p["K"]
        return float(I * J * K), 8.0 * ((I + 4) * (J + 4) + 2 * I * J) / (I * J)
    if bench == "vadv":
        I, J, K = p["I"], p["J"], p["K"]
        return float(I * J * K), 8.0 * (6 * I + 1) / I
    if bench == "jacobi_1d":
        return 2.0 * (p["TSTEPS"] - 1) * (p["N"] - 2), 16.0
    if bench == "seidel_2d":
        return float((p["TSTEPS"] - 1) * (p["N"] - 2) ** 2), 16.0
    if bench == "adi":
        return 2.0 * p["TSTEPS"] * (p["N"] - 2) ** 2, 16.0
    if bench == "channel_flow":
        return float(p["steps"] * (p["nit"] + 2) * p["nx"] * (p["ny"] - 2)), 16.0
    if bench == "cavity_flow":
        return float(p["nt"] * (p["nit"] + 2) * (p["nx"] - 2) * (p["ny"] - 2)), 16.0
    raise KeyError(bench)


def import_suite(conn, bench_json: dict, version: str = "0.1.0", timestamp=None) -> int:
    """Insert the `suite` rows of a bench.py JSON line into `results` (framework b200, details cuda-events)."""
    ts = int(time.time()) if timestamp is None else int(timestamp)
    n = 0
    for row in bench_json.get("suite") or []:
        b = BENCH.get(row.get("kernel"))
        if b is None or "ms" not in row or row.get("preset") not in b["presets"]:
            continue                                         # scaled single-GPU grids are not NPBench presets
        conn.execute(
            "INSERT INTO results(timestamp, benchmark, kind, domain, dwarf, preset, mode, framework, version, details,"
            " validated, time) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
            (ts, b["short"], b["kind"], b["domain"], "structured_grids", row["preset"], "main", "b200", version,
             "cuda-events", 1, row["ms"] * 1e-3))
        n += 1
    conn.commit()
    return n


def refresh_roofline(conn, peak_gbs: float) -> int:
    """(Re)build b200_roofline from every `results` row of a kernel this backend covers."""
    conn.execute(SQL_CREATE_ROOFLINE)
    conn.execute("DELETE FROM b200_roofline")
    rows = conn.execute("SELECT id, benchmark, preset, framework, details, time FROM results").fetchall()
    n = 0
    for rid, short, preset, framework, details, t in rows:
        name = SHORT2NAME.get(short)
        if name is None or preset not in BENCH[name]["presets"] or not t or t <= 0:
            continue
        units, bpu = units_and_bytes(name, BENCH[name]["presets"][preset])
        gcell = units / t / 1e9
        conn.execute("INSERT INTO b200_roofline VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                     (rid, short, preset, framework, details, t, units, gcell, bpu, gcell * bpu,
                      gcell * bpu / peak_gbs))
        n += 1
    conn.commit()
    return n


def summary(conn):
    """Median per (benchmark, preset, framework): [(benchmark, preset, framework, n, median_s, gcell_s, frac)]."""
    out = []
    keys = conn.execute("SELECT DISTINCT benchmark, preset, framework FROM b200_roofline ORDER BY 1, 2, 3").fetchall()
    for b, p, f in keys:
        ts = sorted(r[0] for r in conn.execute(
            "SELECT time FROM b200_roofline WHERE benchmark=? AND preset=? AND framework=?", (b, p, f)))
        med = ts[len(ts) // 2] if len(ts) % 2 else 0.5 * (ts[len(ts) // 2 - 1] + ts[len(ts) // 2])
        units, bpu = units_and_bytes(SHORT2NAME[b], BENCH[SHORT2NAME[b]]["presets"][p])
        peak = conn.execute("SELECT gbps_algorithmic / frac_of_hbm_peak FROM b200_roofline WHERE benchmark=? LIMIT 1",
                            (b,)).fetchone()[0]
        out.append((b, p, f, len(ts), med, units / med / 1e9, units / med / 1e9 * bpu / peak))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--db", default="npbench.db")
    ap.add_argument("--bench-json", help="bench.py output (one JSON line) whose suite rows are imported as b200 results")
    ap.add_argument("--peak-gbs", type=float, default=6553.6, help="HBM peak for the roofline fraction (MEASURED_PEAKS.json)")
    args = ap.parse_args(argv)
    conn = sqlite3.connect(args.db)
    conn.execute(SQL_CREATE_RESULTS)
    if args.bench_json:
        with open(args.bench_json) as f:
            line = [ln for ln in f.read().splitlines() if ln.strip().startswith("{")][-1]
        print("imported %d suite rows" % import_suite(conn, json.loads(line)))
    print("b200_roofline rows: %d" % refresh_roofline(conn, args.peak_gbs))
    print("%-10s %-6s %-8s %3s %12s %10s %6s" % ("benchmark", "preset", "frmwrk", "n", "median s", "Gcell/s", "frac"))
    for b, p, f, n, med, gc, frac in summary(conn):
        print("%-10s %-6s %-8s %3d %12.6f %10.3f %6.3f" % (b, p, f, n, med, gc, frac))
    conn.close()


if __name__ == "__main__":
    main()
