"""Build libnpb_b200.so in-tree with nvcc for sm_100a (no GPU needed to build).

    python -m npbench_b200.build [--force] [--verbose]

-fmad=false: NumPy never contracts a*b+c into an FMA (one rounding per ufunc
pass); every kernel TU keeps NumPy's evaluation order, so disabling contraction
makes the results bit-identical to the reference.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnpb_b200.so")
SOURCES = ["runtime.cu", "host_api.cu", "init_fields.cu", "jacobi2d.cu", "heat3d.cu", "fdtd2d.cu",
           "hdiff.cu", "vadv.cu", "vadv_stream.cu", "jacobi1d.cu", "seidel2d.cu", "adi.cu", "cavity_flow.cu", "channel_flow.cu", "multi_device.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-O2",
              # host-side scalar arithmetic (adi / cavity / channel coefficients) follows the same
              # one-rounding-per-operation contract as the device code and the oracle build
              "-Xcompiler", "-ffp-contract=off"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "npb_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


LAST_BUILD = "not built in this process"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every TU and link; NPB_B200_FORCE_BUILD=1 (or force=True) ignores the mtime check.
    LAST_BUILD says whether this call compiled ("compiled N TUs in S s") or reused the library."""
    global LAST_BUILD
    force = force or os.environ.get("NPB_B200_FORCE_BUILD", "") not in ("", "0")
    if not force and not _stale():
        LAST_BUILD = "reused (up to date with csrc/ and include/)"
        return OUT
    import time
    t0 = time.time()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", OUT] + objs + ["-lcudart"], check=True)
    LAST_BUILD = "compiled %d TUs for sm_100a in %.0f s" % (len(SOURCES), time.time() - t0)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
