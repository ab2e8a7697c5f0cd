"""Write profiles/rNN_summary.md from the committed bench lines, test logs and ncu summaries of a round.

    python tools/make_summary.py 02 > profiles/r02_summary.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(path):
    try:
        return json.loads(open(os.path.join(P, path)).read().strip().splitlines()[-1])
    except Exception:
        return None


def main():
    rr = sys.argv[1] if len(sys.argv) > 1 else "02"
    d = line("r%s_bench_n1.json" % rr)
    out = ["# Round %s — measured results and ncu evidence (B200, sm_100a)" % rr.lstrip("0"), "",
           "All kernel numbers: CUDA events on the launch stream, L2 flushed before every timed call (inputs of the headline are",
           "far larger than L2), SM clock %s MHz, throttle reasons %s (NVML sampled during the timed region).  Peak = %s GB/s, the" % (
               d["clocks"]["sm_mhz"], d["clocks"]["reasons"] or "none", d["roofline"]["peak"]),
           "measured HBM copy rate (`MEASURED_PEAKS.json`).  `frac` = algorithmic bytes / time / peak (BASELINE.md section 2 units).", "",
           "## bench.py, N = 1 (`r%s_bench_n1.json`)" % rr, "",
           "* headline %s: **%.1f Gcell/s**, %.2f ms per step, %d launches per step" % (
               d["config"]["workload"], d["value"], d["ms_per_step"], d["roofline"]["launches_per_step"]),
           "* roofline of the dominant kernel (%s): algorithmic frac %.3f, DRAM-counter frac %.3f (traffic %.2f GB per launch)" % (
               d["roofline"]["kernel"].split(":")[0], d["roofline"]["frac"], d["roofline"].get("frac_dram") or 0, (d["roofline"]["traffic"] or 0) / 1e9),
           "* e2e (pinned host arrays through `npb_jacobi2d_f64_host`): **%.1f Gcell/s**, %.0f ms per call, H2D %.2f GB + D2H %.2f GB; matches the device-array call bit for bit: %s" % (
               d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e9, d["e2e"]["d2h_bytes_per_step"] / 1e9, d["e2e"].get("matches_device_path")),
           "* CPU beside it: reference NumPy kernel %.3f Gcell/s on %d core (%s); C/OpenMP port %.2f Gcell/s on %d cores" % (
               d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"], d["cpu_port"]["value"], d["cpu_port"]["cores"]),
           "* parity field: %s" % json.dumps(d["parity"]["bit_exact"]), "",
           "### named records (BASELINE.json configs[0..3])", "",
           "| record | Gcell/s | ms | frac | kernel | DRAM bytes per launch (ncu) | e2e Gcell/s | harness -f b200 ms (validated) | harness -f numpy ms |", "|---|---|---|---|---|---|---|---|---|"]
    for k, v in d["records"].items():
        if not isinstance(v, dict) or "value" not in v:
            continue
        h = v.get("harness", {})
        hb, hn = h.get("b200", {}), h.get("numpy", {})
        out.append("| %s | %.1f | %.4f | %.3f | %s | %s | %s | %s (%s) | %s |" % (
            k, v["value"], v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["kernel"].split(" ")[0], v["roofline"]["traffic"],
            v.get("e2e", {}).get("value"), hb.get("wall_ms_median"), hb.get("validated"), hn.get("wall_ms_median")))
    out += ["", "### every kernel x preset (`suite`)", "", "| kernel | preset | ms | Gcell/s | frac of peak | launches/call |", "|---|---|---|---|---|---|"]
    for r in d["suite"]:
        out.append("| %s | %s | %.4f | %.2f | %.3f | %d |" % (r["kernel"], r["preset"], r["ms"], r["value"], r["frac_of_peak"], r["launches"]))
    out += ["", "## multi-GPU (`r%s_bench_n{2,4,8}.json`, `r%s_check_multigpu_n{2,4,8}.log`)" % (rr, rr), "",
            "| GPUs | jacobi_2d weak-scaled Gcell/s | one slab, same run | efficiency | e2e Gcell/s | parity | suite rows (Gcell/s, weak-scaling efficiency) |", "|---|---|---|---|---|---|---|"]
    for n in (2, 4, 8):
        m = line("r%s_bench_n%d.json" % (rr, n))
        if not m:
            continue
        one = m["single_gpu_same_workload"]["value"]
        rows = "; ".join("%s %.0f%s" % (r["kernel"], r["value"], (" (%.2f)" % r["weak_scaling_efficiency"]) if "weak_scaling_efficiency" in r else "")
                         for r in (m.get("suite") or []))
        out.append("| %d | %.0f | %.0f | %.3f | %s | %s | %s |" % (n, m["value"], one, m["value"] / (n * one), m["e2e"].get("value"), m["parity"]["bit_exact"], rows))
    for n in (2, 4, 8):         # re-runs late in the round (headline only), after the last kernel changes
        m = line("r%s_bench_n%d_late.json" % (rr, n))
        if m:
            one = m["single_gpu_same_workload"]["value"]
            out.append("| %d (late re-run: short chunks + two-state closing pass) | %.0f | %.0f | %.3f | %s | %s | — |" % (
                n, m["value"], one, m["value"] / (n * one), m["e2e"].get("value"), m["parity"]["bit_exact"]))
    for n in (2, 4, 8):
        p = os.path.join(P, "r%s_check_multigpu_n%d.log" % (rr, n))
        if os.path.exists(p):
            out.append("")
            out.append("`tools/check_multigpu.py`, %d ranks: `%s`" % (n, open(p).read().strip().splitlines()[-1]))
    p = os.path.join(P, "r%s_pytest_gpu.log" % rr)
    if os.path.exists(p):
        out += ["", "## tests", "", "`pytest -m gpu`: %s" % open(p).read().strip().splitlines()[-1]]
    p = os.path.join(P, "r%s_launches_bench.csv" % rr)
    if os.path.exists(p):
        import collections
        import csv
        agg = collections.defaultdict(lambda: [0, 0.0])
        scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
        for r in csv.reader(open(p)):
            if len(r) > 10 and r[0].isdigit():
                a = agg[r[4].split("(")[0].replace("<unnamed>::", "")]
                a[0] += 1
                a[1] += float(r[-1].replace(",", "")) * scale.get(r[-2], 1e-6)
        tot = sum(v[1] for v in agg.values()) or 1.0
        out += ["", "## ncu launch list of the bench command (`r%s_launches_bench.csv`: first 600 launches of `bench.py --steps 2 --warmup 1`, "
                "`gpu__time_duration.sum`, cold caches, serialised)" % rr, "",
                "The timed headline steps are the full-grid `jacobi2d_march_kernel<7, 1, 0>` / `<5, 1, 0>` / `<7, 1, 1>` (two-state closing pass) launches; "
                "the many short `<5, 1, 0>` / `<1, 1, 0>` launches are the row-chunk pipeline of the e2e host-buffer call.", "",
                "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
            out.append("| `%s` | %d | %.3f | %.1f %% |" % (k, v[0], v[1], 100.0 * v[1] / tot))
    out += ["", "## ncu summaries in this directory", ""]
    for f in sorted(os.listdir(P)):
        if f.startswith("r%s_ncu_" % rr) and f.endswith(".txt"):
            t = open(os.path.join(P, f)).read().splitlines()
            kv = {l[:75].strip(): l[75:].strip() for l in t[:20]}
            out.append("* `%s`: %s, %s, DRAM read %s + write %s, issue active %s, FP64 pipe %s" % (
                f, kv.get("Kernel Name", "")[:70], kv.get("gpu__time_duration.sum"), kv.get("dram__bytes_read.sum"), kv.get("dram__bytes_write.sum"),
                kv.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), kv.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")))
    print("\n".join(out))


if __name__ == "__main__":
    main()
