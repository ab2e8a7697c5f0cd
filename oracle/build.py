"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE -- see stencil_oracle.c).

`python -m oracle.build` compiles oracle/stencil_oracle.c into
oracle/liboracle.so with contraction disabled (one IEEE rounding per operation,
as NumPy's ufunc passes do) and OpenMP enabled.  The reference is pure Python
(no C/C++ sources to compile), so there is no oracle/_ref build.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "stencil_oracle.c")
OUT = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
           "-ffp-contract=off", "-fno-fast-math", "-fno-builtin-pow", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
