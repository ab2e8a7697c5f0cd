"""Pretty-print a bench.py JSON line (headline + suite table)."""
import json, sys
d = json.load(open(sys.argv[1]))
s = d.pop("suite", None) or []
for k in ("value", "ms_per_step", "gpu_launches"):
    print(k, d.get(k))
for k in ("roofline", "e2e", "cpu_baseline", "clocks"):
    v = d.get(k) or {}
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("note", "sample", "api", "peak_source")})
print("%-10s %-12s %10s %10s %8s %8s" % ("kernel", "preset", "ms", "Gcell/s", "frac", "launch"))
for r in s:
    if "error" in r:
        print(r)
    elif "workload" in r:
        print("%-10s %-70s %10.4f %10.2f %8.3f" % (r["kernel"], r["workload"][:70], r["ms"], r["value"], r["frac_of_peak_per_gpu"]))
    else:
        print("%-10s %-12s %10.4f %10.2f %8.3f %8d" % (r["kernel"], r["preset"], r["ms"], r["value"], r["frac_of_peak"], r["launches"]))
