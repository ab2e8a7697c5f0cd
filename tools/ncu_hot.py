"""Summarise an ncu source-page CSV: top SASS instructions by stall samples + stall-reason totals.
usage: ncu -i X.ncu-rep --page source --csv > x.csv ; python tools/ncu_hot.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = idx["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in reasons}
print("total samples", tot)
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print("top instructions:")
for n, r in sorted(enumerate(data), key=lambda t: -int(t[1][S] or 0))[:topn]:
    top = sorted(((int(r[idx[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
    print("%5d %5.1f%%  #%-4d %-70s %s" % (int(r[S]), 100.0 * int(r[S]) / max(tot, 1), n, r[idx["Source"]].strip()[:70],
                                           " ".join("%s:%d" % (b, a) for a, b in top if a)))
