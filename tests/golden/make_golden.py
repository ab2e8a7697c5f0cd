"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

Imports the reference's own NumPy implementations and `initialize` functions

    npbench/benchmarks/polybench/{jacobi_2d,heat_3d,fdtd_2d}/<b>{,_numpy}.py
    npbench/benchmarks/weather_stencils/{hdiff,vadv}/<b>{,_numpy}.py

runs them, and writes
  * tests/golden/pins.json   -- sha256 + sum of every input and output of the NPBench
                                presets S and M (inputs come from the reference `initialize`);
  * tests/golden/cases.npz   -- full input/output arrays of small seeded cases (odd shapes,
                                random data, degenerate step counts) that the reference's own
                                inputs never exercise (heat_3d's NPBench input is a fixed point).
The GPU box has no /root/reference: tests only read these two files.
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("NPBENCH_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


def ref(relpath, module, func):
    return getattr(importlib.import_module("npbench.benchmarks.%s.%s" % (relpath.replace("/", "."), module)), func)


def digest(a):
    a = np.ascontiguousarray(a)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "sum": float(a.sum()), "shape": list(a.shape)}


PRESETS = {  # bench_info/*.json "parameters"
    "jacobi_2d": {"S": dict(TSTEPS=50, N=150), "M": dict(TSTEPS=80, N=350)},
    "heat_3d": {"S": dict(TSTEPS=25, N=25), "M": dict(TSTEPS=50, N=40)},
    "fdtd_2d": {"S": dict(TMAX=20, NX=200, NY=220), "M": dict(TMAX=60, NX=400, NY=450)},
    "hdiff": {"S": dict(I=64, J=64, K=60), "M": dict(I=128, J=128, K=160)},
    "vadv": {"S": dict(I=60, J=60, K=40), "M": dict(I=112, J=112, K=80)},
}


def main():
    j_init = ref("polybench/jacobi_2d", "jacobi_2d", "initialize")
    j_kern = ref("polybench/jacobi_2d", "jacobi_2d_numpy", "kernel")
    h_init = ref("polybench/heat_3d", "heat_3d", "initialize")
    h_kern = ref("polybench/heat_3d", "heat_3d_numpy", "kernel")
    f_init = ref("polybench/fdtd_2d", "fdtd_2d", "initialize")
    f_kern = ref("polybench/fdtd_2d", "fdtd_2d_numpy", "kernel")
    d_init = ref("weather_stencils/hdiff", "hdiff", "initialize")
    d_kern = ref("weather_stencils/hdiff", "hdiff_numpy", "hdiff")
    v_init = ref("weather_stencils/vadv", "vadv", "initialize")
    v_kern = ref("weather_stencils/vadv", "vadv_numpy", "vadv")

    pins = {"numpy": np.__version__}
    for preset, p in PRESETS["jacobi_2d"].items():
        A, B = j_init(p["N"])
        e = {"in": {"A": digest(A), "B": digest(B)}}
        j_kern(p["TSTEPS"], A, B)
        e["out"] = {"A": digest(A), "B": digest(B)}
        pins["jacobi_2d/" + preset] = e
    for preset, p in PRESETS["heat_3d"].items():
        A, B = h_init(p["N"])
        e = {"in": {"A": digest(A), "B": digest(B)}}
        h_kern(p["TSTEPS"], A, B)
        e["out"] = {"A": digest(A), "B": digest(B)}
        pins["heat_3d/" + preset] = e
    for preset, p in PRESETS["fdtd_2d"].items():
        ex, ey, hz, fict = f_init(p["TMAX"], p["NX"], p["NY"])
        e = {"in": {"ex": digest(ex), "ey": digest(ey), "hz": digest(hz), "_fict_": digest(fict)}}
        f_kern(p["TMAX"], ex, ey, hz, fict)
        e["out"] = {"ex": digest(ex), "ey": digest(ey), "hz": digest(hz)}
        pins["fdtd_2d/" + preset] = e
    for preset, p in PRESETS["hdiff"].items():
        inf, outf, coeff = d_init(p["I"], p["J"], p["K"])
        e = {"in": {"in_field": digest(inf), "out_field": digest(outf), "coeff": digest(coeff)}}
        d_kern(inf, outf, coeff)
        e["out"] = {"out_field": digest(outf)}
        pins["hdiff/" + preset] = e
    for preset, p in PRESETS["vadv"].items():
        dtr, us, u, w, up, ut = v_init(p["I"], p["J"], p["K"])
        e = {"in": {"utens_stage": digest(us), "u_stage": digest(u), "wcon": digest(w),
                    "u_pos": digest(up), "utens": digest(ut)}, "dtr_stage": dtr}
        v_kern(us, u, w, up, ut, dtr)
        e["out"] = {"utens_stage": digest(us)}
        pins["vadv/" + preset] = e
    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)

    # ---- small seeded cases with full arrays --------------------------------
    cases = {}
    rng = np.random.default_rng(20261017)

    def put(name, **arrs):
        for k, v in arrs.items():
            cases["%s.%s" % (name, k)] = np.asarray(v)

    for n, (ts, shape) in enumerate([(1, (7, 7)), (2, (7, 7)), (3, (67, 67)), (5, (64, 64)), (20, (131, 131)),
                                     (4, (37, 53)), (6, (3, 9)), (9, (130, 70))]):
        A = rng.random(shape); B = rng.random(shape)
        A0, B0 = A.copy(), B.copy()
        j_kern(ts, A, B)
        put("jacobi_2d.%d" % n, TSTEPS=ts, A_in=A0, B_in=B0, A_out=A, B_out=B)
    for n, (ts, shape) in enumerate([(1, (5, 5, 5)), (2, (5, 5, 5)), (3, (23, 23, 23)), (6, (21, 22, 23)),
                                     (4, (17, 17, 17)), (5, (9, 12, 15)), (3, (3, 3, 3)), (7, (12, 19, 40))]):
        A = rng.random(shape); B = rng.random(shape)
        A0, B0 = A.copy(), B.copy()
        h_kern(ts, A, B)
        put("heat_3d.%d" % n, TSTEPS=ts, A_in=A0, B_in=B0, A_out=A, B_out=B)
    for n, (tm, nx, ny) in enumerate([(0, 5, 6), (1, 5, 6), (3, 17, 33), (7, 64, 50), (12, 33, 129), (5, 2, 2),
                                      (4, 1, 9), (4, 9, 1)]):
        ex, ey, hz = rng.random((nx, ny)), rng.random((nx, ny)), rng.random((nx, ny))
        fict = rng.random((max(tm, 1),))
        i = dict(ex_in=ex.copy(), ey_in=ey.copy(), hz_in=hz.copy(), fict=fict.copy())
        f_kern(tm, ex, ey, hz, fict)
        put("fdtd_2d.%d" % n, TMAX=tm, ex_out=ex, ey_out=ey, hz_out=hz, **i)
    for n, (I, J, K) in enumerate([(1, 1, 1), (3, 5, 7), (17, 9, 33), (8, 8, 160), (5, 40, 2), (33, 3, 60)]):
        inf = rng.random((I + 4, J + 4, K)) - 0.5
        outf = rng.random((I, J, K)); coeff = rng.random((I, J, K))
        o = outf.copy()
        d_kern(inf, o, coeff)
        put("hdiff.%d" % n, in_field=inf, coeff=coeff, out_in=outf, out_field=o)
    for n, (I, J, K) in enumerate([(1, 1, 2), (2, 3, 3), (5, 7, 4), (4, 6, 17), (3, 3, 160), (9, 33, 40), (2, 40, 80)]):
        us, u, up, ut = (rng.random((I, J, K)) for _ in range(4))
        w = rng.random((I + 1, J, K))
        dtr = 3.0 / 20.0
        o = us.copy()
        v_kern(o, u, w, up, ut, dtr)
        put("vadv.%d" % n, utens_stage_in=us, u_stage=u, wcon=w, u_pos=up, utens=ut, dtr_stage=dtr,
            utens_stage_out=o)
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **cases)
    print("pins:", len(pins) - 1, "cases arrays:", len(cases),
          "npz bytes:", os.path.getsize(os.path.join(HERE, "cases.npz")))


if __name__ == "__main__":
    main()
