"""SASS evidence for the hot kernels: per kernel of libnpb_b200.so the counts of the mnemonics that prove the
design (TMA tensor / bulk copies, tensor memory, mbarriers, shuffles, FP64 pipe), registers and spills.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt      (cuobjdump only; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "npbench_b200", "libnpb_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "STTM", "LDTM", "UTCBAR", "LDG", "STG", "LDS", "STS",
        "LDSM", "SHFL", "DFMA", "DADD", "DMUL", "MUFU", "BAR", "MEMBAR", "ATOM", "RED", "NANOSLEEP", "UCGABAR", "HMMA",
        "UTCHMMA", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.|\s|;)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            if op in KEYS:
                counts[cur][op] += 1
    names = demangle(list(counts))
    print("SASS summary of npbench_b200/libnpb_b200.so (sm_100a), produced by tools/sass_summary.py")
    print("mnemonics: UTMALDG/UTMASTG/UTMAPF = TMA tensor load/store/prefetch, UBLKCP = cp.async.bulk, SYNCS = mbarrier,")
    print("           STTM/LDTM = tcgen05.st/ld (tensor memory), UCGABAR = cluster barrier, LDL/STL = local-memory spills\n")
    for k, c in counts.items():
        nm = re.sub(r"\(anonymous namespace\)::", "", names.get(k, k))
        nm = re.sub(r"\(.*", "", nm)
        r = regs.get(k)
        tail = " regs=%d static_smem=%d local=%d" % r if r else ""
        print("%s\n    instructions=%d%s" % (nm, c["_total"], tail))
        print("    " + "  ".join("%s=%d" % (x, c[x]) for x in KEYS if c[x]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
