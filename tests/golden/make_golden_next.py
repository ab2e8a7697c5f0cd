"""Golden fixtures for the widening row: jacobi_1d, seidel_2d and (rank 2) adi.

    python tests/golden/make_golden_next.py            # needs /root/reference (read-only)

Imports the UNMODIFIED reference
    npbench/benchmarks/polybench/jacobi_1d/{jacobi_1d,jacobi_1d_numpy}.py
    npbench/benchmarks/polybench/seidel_2d/{seidel_2d,seidel_2d_numpy}.py
runs it and writes tests/golden/pins_next.json (sha256 + sum of inputs/outputs of the NPBench presets
S and M, seidel_2d also L) and tests/golden/cases_next.npz (full arrays of small seeded cases: random
data, odd sizes, degenerate step counts).  Kept apart from make_golden.py so that the fixtures of the
five hot-path kernels stay byte-identical.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import digest, ref  # noqa: E402  (also puts the reference on sys.path)

PRESETS = {  # bench_info/{jacobi_1d,seidel_2d}.json "parameters"
    "jacobi_1d": {"S": dict(TSTEPS=800, N=3200), "M": dict(TSTEPS=3000, N=12000)},
    "seidel_2d": {"S": dict(TSTEPS=8, N=50), "M": dict(TSTEPS=15, N=100), "L": dict(TSTEPS=40, N=200)},
    "adi": {"S": dict(TSTEPS=5, N=100), "M": dict(TSTEPS=20, N=200), "paper": dict(TSTEPS=100, N=200)},
    "cavity_flow": {"S": dict(ny=61, nx=61, nt=25, nit=5, rho=1.0, nu=0.1), "M": dict(ny=121, nx=121, nt=50, nit=10, rho=1.0, nu=0.1)},
    "channel_flow": {"S": dict(ny=61, nx=61, nit=5, rho=1.0, nu=0.1, F=1.0), "M": dict(ny=121, nx=121, nit=10, rho=1.0, nu=0.1, F=1.0),
                     "paper": dict(ny=101, nx=101, nit=50, rho=1.0, nu=0.1, F=1.0)},
}


def main():
    j_init = ref("polybench/jacobi_1d", "jacobi_1d", "initialize")
    j_kern = ref("polybench/jacobi_1d", "jacobi_1d_numpy", "kernel")
    s_init = ref("polybench/seidel_2d", "seidel_2d", "initialize")
    s_kern = ref("polybench/seidel_2d", "seidel_2d_numpy", "kernel")
    pins = {"numpy": np.__version__}
    for preset, p in PRESETS["jacobi_1d"].items():
        A, B = j_init(p["N"])
        e = {"in": {"A": digest(A), "B": digest(B)}}
        j_kern(p["TSTEPS"], A, B)
        e["out"] = {"A": digest(A), "B": digest(B)}
        pins["jacobi_1d/" + preset] = e
    for preset, p in PRESETS["seidel_2d"].items():
        A = s_init(p["N"])
        e = {"in": {"A": digest(A)}}
        s_kern(p["TSTEPS"], p["N"], A)
        e["out"] = {"A": digest(A)}
        pins["seidel_2d/" + preset] = e
    a_init = ref("polybench/adi", "adi", "initialize")
    a_kern = ref("polybench/adi", "adi_numpy", "kernel")
    for preset, p in PRESETS["adi"].items():
        u = a_init(p["N"])
        e = {"in": {"u": digest(u)}}
        r = a_kern(p["TSTEPS"], p["N"], u)
        assert r is u
        e["out"] = {"u": digest(u)}
        pins["adi/" + preset] = e
    c_init = ref("cavity_flow", "cavity_flow", "initialize")
    c_kern = ref("cavity_flow", "cavity_flow_numpy", "cavity_flow")
    for preset, p in PRESETS["cavity_flow"].items():
        u, v, pr, dx, dy, dt = c_init(p["ny"], p["nx"])
        e = {"in": {"u": digest(u), "v": digest(v), "p": digest(pr)}, "dx": dx, "dy": dy, "dt": dt}
        c_kern(p["nx"], p["ny"], p["nt"], p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"])
        e["out"] = {"u": digest(u), "v": digest(v), "p": digest(pr)}
        pins["cavity_flow/" + preset] = e
    h_init = ref("channel_flow", "channel_flow", "initialize")
    h_kern = ref("channel_flow", "channel_flow_numpy", "channel_flow")
    for preset, p in PRESETS["channel_flow"].items():
        u, v, pr, dx, dy, dt = h_init(p["ny"], p["nx"])
        e = {"in": {"u": digest(u), "v": digest(v), "p": digest(pr)}, "dx": dx, "dy": dy, "dt": dt}
        e["stepcount"] = int(h_kern(p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"], p["F"]))
        e["out"] = {"u": digest(u), "v": digest(v), "p": digest(pr)}
        pins["channel_flow/" + preset] = e
    with open(os.path.join(HERE, "pins_next.json"), "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)

    cases = {}
    rng = np.random.default_rng(20261018)

    def put(name, **arrs):
        for k, v in arrs.items():
            cases["%s.%s" % (name, k)] = np.asarray(v)

    for n, (ts, N) in enumerate([(1, 9), (2, 9), (2, 3), (3, 4), (5, 2), (40, 257), (7, 1000), (131, 700), (300, 5000),
                                 (65, 33)]):
        A = rng.random((N,)) - 0.5; B = rng.random((N,)) - 0.5
        A0, B0 = A.copy(), B.copy()
        j_kern(ts, A, B)
        put("jacobi_1d.%d" % n, TSTEPS=ts, A_in=A0, B_in=B0, A_out=A, B_out=B)
    for n, (ts, N) in enumerate([(1, 7), (2, 3), (2, 4), (2, 7), (3, 8), (5, 19), (9, 33), (4, 64), (12, 37), (3, 90)]):
        A = rng.random((N, N)) - 0.5
        A0 = A.copy()
        s_kern(ts, N, A)
        put("seidel_2d.%d" % n, TSTEPS=ts, N=N, A_in=A0, A_out=A)
    # adi cases draw AFTER the others so that the earlier fixtures keep their values
    for n, (ts, N) in enumerate([(1, 3), (1, 4), (2, 5), (3, 9), (4, 33), (7, 64), (2, 100), (5, 47)]):
        u = rng.random((N, N)) - 0.5
        u0 = u.copy()
        a_kern(ts, N, u)
        put("adi.%d" % n, TSTEPS=ts, N=N, u_in=u0, u_out=u)
    # cavity_flow cases draw after everything else; random (not physical) fields, rectangular grids, nit/nt edge values
    for n, (nx, ny, nt, nit) in enumerate([(3, 3, 1, 1), (5, 4, 2, 3), (9, 7, 3, 0), (17, 33, 4, 5), (40, 21, 0, 3),
                                           (64, 64, 3, 7), (31, 50, 5, 2)]):
        u, v, pr = rng.random((ny, nx)) - 0.5, rng.random((ny, nx)) - 0.5, rng.random((ny, nx)) - 0.5
        dx, dy = 2 / (nx - 1), 2 / (ny - 1)
        dt = .1 / ((nx - 1) * (ny - 1))
        i = dict(u_in=u.copy(), v_in=v.copy(), p_in=pr.copy())
        c_kern(nx, ny, nt, nit, u, v, dt, dx, dy, pr, 1.3, 0.07)
        put("cavity_flow.%d" % n, nx=nx, ny=ny, nt=nt, nit=nit, dt=dt, dx=dx, dy=dy, rho=1.3, nu=0.07, u_out=u, v_out=v, p_out=pr, **i)
    # channel_flow cases (drawn last): the reference's own initial fields on small / rectangular grids with other
    # nit, rho, nu, F, plus np.sum reference values for the pairwise-summation restatement
    for n, (nx, ny, nit, rho, nu, F) in enumerate([(3, 3, 1, 1.0, 0.1, 1.0), (9, 7, 3, 1.0, 0.1, 1.0), (17, 12, 0, 1.0, 0.2, 0.5),
                                                   (31, 21, 4, 1.3, 0.07, 1.0), (16, 33, 6, 0.9, 0.15, 2.0)]):
        u, v, pr, dx, dy, dt = h_init(ny, nx)
        sc = h_kern(nit, u, v, dt, dx, dy, pr, rho, nu, F)
        put("channel_flow.%d" % n, nx=nx, ny=ny, nit=nit, dt=dt, dx=dx, dy=dy, rho=rho, nu=nu, F=F, stepcount=sc,
            u_out=u, v_out=v, p_out=pr)
    for n, size in enumerate([1, 7, 8, 9, 100, 128, 129, 1000, 3721, 10201, 14641, 40401]):
        a = (rng.random(size) - 0.3) * 10.0 ** float(rng.integers(-3, 4))
        put("npsum.%d" % n, a=a, s=np.float64(np.sum(a)))
    np.savez_compressed(os.path.join(HERE, "cases_next.npz"), **cases)
    print("pins:", len(pins) - 1, "cases arrays:", len(cases),
          "npz bytes:", os.path.getsize(os.path.join(HERE, "cases_next.npz")))


if __name__ == "__main__":
    main()
