"""python -m npbench_b200.run [--reference DIR] [--overlay DIR] [--script run_benchmark.py] -- <NPBench CLI args>

Example (the BASELINE config #1 flow):
    python -m npbench_b200.run -- -b jacobi_2d -f numpy -p S
    python -m npbench_b200.run -- -b jacobi_2d -f b200 -p S
Runs the reference's own CLI, unmodified, against an overlay that contains the b200 plugin
(see npbench_b200/overlay.py).  npbench.db is written to the current directory.
"""
import argparse
import os
import sys
import tempfile
import zlib

from . import overlay as _ov


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    rest = []
    if "--" in argv:
        i = argv.index("--")
        argv, rest = argv[:i], argv[i + 1:]
    ap = argparse.ArgumentParser(prog="npbench_b200.run")
    ap.add_argument("--reference", default=None,
                    help="NPBench checkout (default: $NPBENCH_REF, /root/reference, baseline/_ref)")
    ap.add_argument("--overlay", default=None)
    ap.add_argument("--script", default="run_benchmark.py")
    a = ap.parse_args(argv)
    ref = _ov.find_reference(a.reference)
    # one overlay per checkout: symlinks into a different checkout must not be reused
    tag = "%08x" % zlib.crc32(ref.encode())
    ov = a.overlay or os.path.join(tempfile.gettempdir(), "npbench_b200_overlay_" + tag)
    _ov.run_cli(ref, ov, a.script, rest)


if __name__ == "__main__":
    main()
