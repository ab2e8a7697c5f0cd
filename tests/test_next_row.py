"""Widening row: rank 1 jacobi_1d and seidel_2d, rank 2 adi.

CPU part: the oracle (oracle/stencil_oracle.c: npb_oracle_jacobi1d / npb_oracle_seidel2d) against outputs of the
unmodified reference (jacobi_1d_numpy.py:4-8, seidel_2d_numpy.py:4-13; fixtures from
tests/golden/make_golden_next.py).  GPU part (-m gpu): the CUDA kernels through the C ABI
(npb_jacobi1d_f64, npb_seidel2d_f64) against the oracle and the fixtures.  Bar: bit-exact.
"""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, assert_bit_equal, sha


@pytest.fixture(scope="module")
def pins_next():
    with open(os.path.join(GOLDEN, "pins_next.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def cases_next():
    z = np.load(os.path.join(GOLDEN, "cases_next.npz"))
    out = {}
    for key in z.files:
        bench, idx, field = key.split(".", 2)
        out.setdefault(bench, {}).setdefault(int(idx), {})[field] = z[key]
    return {b: [d[i] for i in sorted(d)] for b, d in out.items()}


# ------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("preset", ["S", "M"])
def test_oracle_jacobi_1d_presets(pins_next, preset):
    p = oracle.PRESETS["jacobi_1d"][preset]; pin = pins_next["jacobi_1d/" + preset]
    A, B = oracle.init_jacobi_1d(p["N"])
    assert sha(A) == pin["in"]["A"]["sha256"] and sha(B) == pin["in"]["B"]["sha256"]
    oracle.jacobi_1d(p["TSTEPS"], A, B)
    assert sha(A) == pin["out"]["A"]["sha256"] and sha(B) == pin["out"]["B"]["sha256"]


@pytest.mark.parametrize("preset", ["S", "M", "L"])
def test_oracle_seidel_2d_presets(pins_next, preset):
    p = oracle.PRESETS["seidel_2d"][preset]; pin = pins_next["seidel_2d/" + preset]
    A = oracle.init_seidel_2d(p["N"])
    assert sha(A) == pin["in"]["A"]["sha256"]
    oracle.seidel_2d(p["TSTEPS"], p["N"], A)
    assert sha(A) == pin["out"]["A"]["sha256"]


@pytest.mark.parametrize("preset", ["S", "M", "paper"])
def test_oracle_adi_presets(pins_next, preset):
    p = oracle.PRESETS["adi"][preset]; pin = pins_next["adi/" + preset]
    u = oracle.init_adi(p["N"])
    assert sha(u) == pin["in"]["u"]["sha256"]
    oracle.set_threads(4)
    try:
        oracle.adi(p["TSTEPS"], p["N"], u)
    finally:
        oracle.set_threads(1)
    assert sha(u) == pin["out"]["u"]["sha256"]


@pytest.mark.parametrize("preset", ["S", "M"])
def test_oracle_cavity_flow_presets(pins_next, preset):
    p = oracle.PRESETS["cavity_flow"][preset]; pin = pins_next["cavity_flow/" + preset]
    u, v, pr, dx, dy, dt = oracle.init_cavity_flow(p["ny"], p["nx"])
    assert (dx, dy, dt) == (pin["dx"], pin["dy"], pin["dt"])
    oracle.set_threads(4)
    try:
        oracle.cavity_flow(p["nx"], p["ny"], p["nt"], p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"])
    finally:
        oracle.set_threads(1)
    for name, a in (("u", u), ("v", v), ("p", pr)):
        assert sha(a) == pin["out"][name]["sha256"], name


@pytest.mark.parametrize("preset", ["S", "M", "paper"])
def test_oracle_channel_flow_presets(pins_next, preset):
    p = oracle.PRESETS["channel_flow"][preset]; pin = pins_next["channel_flow/" + preset]
    u, v, pr, dx, dy, dt = oracle.init_channel_flow(p["ny"], p["nx"])
    oracle.set_threads(8)
    try:
        sc = oracle.channel_flow(p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"], p["F"])
    finally:
        oracle.set_threads(1)
    assert sc == pin["stepcount"]
    for name, a in (("u", u), ("v", v), ("p", pr)):
        assert sha(a) == pin["out"][name]["sha256"], name


def test_oracle_np_sum(cases_next):
    for n, c in enumerate(cases_next["npsum"]):
        assert oracle.np_sum(c["a"]) == float(c["s"]), "np.sum case %d (n=%d)" % (n, c["a"].size)


def test_oracle_small_cases(cases_next):
    for n, c in enumerate(cases_next["channel_flow"]):
        u, v, pr, dx, dy, dt = oracle.init_channel_flow(int(c["ny"]), int(c["nx"]))
        sc = oracle.channel_flow(int(c["nit"]), u, v, dt, dx, dy, pr, float(c["rho"]), float(c["nu"]), float(c["F"]))
        assert sc == int(c["stepcount"]), "channel_flow.%d stepcount" % n
        for name, a in (("u", u), ("v", v), ("p", pr)):
            assert_bit_equal(a, c[name + "_out"], "channel_flow.%d %s" % (n, name))
    for n, c in enumerate(cases_next["cavity_flow"]):
        u, v, pr = c["u_in"].copy(), c["v_in"].copy(), c["p_in"].copy()
        oracle.cavity_flow(int(c["nx"]), int(c["ny"]), int(c["nt"]), int(c["nit"]), u, v, float(c["dt"]), float(c["dx"]),
                           float(c["dy"]), pr, float(c["rho"]), float(c["nu"]))
        for name, a in (("u", u), ("v", v), ("p", pr)):
            assert_bit_equal(a, c[name + "_out"], "cavity_flow.%d %s" % (n, name))
    for n, c in enumerate(cases_next["adi"]):
        u = c["u_in"].copy()
        oracle.adi(int(c["TSTEPS"]), int(c["N"]), u)
        assert_bit_equal(u, c["u_out"], "adi.%d" % n)
    for n, c in enumerate(cases_next["jacobi_1d"]):
        A, B = c["A_in"].copy(), c["B_in"].copy()
        oracle.jacobi_1d(int(c["TSTEPS"]), A, B)
        assert_bit_equal(A, c["A_out"], "jacobi_1d.%d A" % n); assert_bit_equal(B, c["B_out"], "jacobi_1d.%d B" % n)
    for n, c in enumerate(cases_next["seidel_2d"]):
        A = c["A_in"].copy()
        oracle.seidel_2d(int(c["TSTEPS"]), int(c["N"]), A)
        assert_bit_equal(A, c["A_out"], "seidel_2d.%d" % n)


# ------------------------------------------------------------------ CUDA kernels
@pytest.fixture(scope="module")
def nb():
    import npbench_b200 as nb_
    nb_.init(0)
    return nb_


def run_j1(nb, ts, A, B, host):
    if host:
        a, b = A.copy(), B.copy()
        nb.jacobi_1d(ts, a, b)
        return a, b
    dA, dB = nb.DeviceArray.from_host(A), nb.DeviceArray.from_host(B)
    nb.jacobi_1d(ts, dA, dB)
    return dA.to_host(), dB.to_host()


def run_s2(nb, ts, A, host):
    if host:
        a = A.copy()
        nb.seidel_2d(ts, A.shape[0], a)
        return a
    dA = nb.DeviceArray.from_host(A)
    nb.seidel_2d(ts, A.shape[0], dA)
    return dA.to_host()


def run_adi(nb, ts, u, host):
    if host:
        a = u.copy()
        r = nb.adi(ts, u.shape[0], a)
        assert r is a                       # the reference returns its (mutated) argument: adi_numpy.py:54
        return a
    d = nb.DeviceArray.from_host(u)
    assert nb.adi(ts, u.shape[0], d) is d
    return d.to_host()


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L", "paper"])
def test_gpu_adi_presets(nb, pins_next, preset):
    p = oracle.PRESETS["adi"][preset]
    u = oracle.init_adi(p["N"])
    g = run_adi(nb, p["TSTEPS"], u, host=False)
    if "adi/" + preset in pins_next:
        assert sha(g) == pins_next["adi/" + preset]["out"]["u"]["sha256"]
    oracle.set_threads(4)
    try:
        oracle.adi(p["TSTEPS"], p["N"], u)
    finally:
        oracle.set_threads(1)
    assert_bit_equal(g, u, "u")


@pytest.mark.gpu
@pytest.mark.parametrize("ts,n", [(1, 3), (1, 2), (2, 4), (3, 31), (2, 32), (4, 33), (6, 65), (3, 257), (9, 100)])
def test_gpu_adi_random(nb, ts, n):
    rng = np.random.default_rng(ts * 1000 + n)
    u = rng.random((n, n)) - 0.5
    g = run_adi(nb, ts, u, host=False)
    oracle.adi(ts, n, u)
    assert_bit_equal(g, u, "u")


def run_cavity(nb, c, host):
    args = (int(c["nx"]), int(c["ny"]), int(c["nt"]), int(c["nit"]))
    sc = (float(c["dt"]), float(c["dx"]), float(c["dy"]))
    if host:
        u, v, p = c["u_in"].copy(), c["v_in"].copy(), c["p_in"].copy()
        nb.cavity_flow(*args, u, v, *sc, p, float(c["rho"]), float(c["nu"]))
        return u, v, p
    u, v, p = (nb.DeviceArray.from_host(c[k]) for k in ("u_in", "v_in", "p_in"))
    nb.cavity_flow(*args, u, v, *sc, p, float(c["rho"]), float(c["nu"]))
    return u.to_host(), v.to_host(), p.to_host()


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True], ids=["device", "host"])
def test_gpu_cavity_flow_golden_cases(nb, cases_next, host):
    for n, c in enumerate(cases_next["cavity_flow"]):
        for name, g in zip(("u", "v", "p"), run_cavity(nb, c, host)):
            assert_bit_equal(g, c[name + "_out"], "cavity_flow.%d %s" % (n, name))


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L", "paper"])
def test_gpu_cavity_flow_presets(nb, pins_next, preset):
    p = oracle.PRESETS["cavity_flow"][preset]
    u, v, pr, dx, dy, dt = oracle.init_cavity_flow(p["ny"], p["nx"])
    c = dict(nx=p["nx"], ny=p["ny"], nt=p["nt"], nit=p["nit"], dt=dt, dx=dx, dy=dy, rho=p["rho"], nu=p["nu"],
             u_in=u, v_in=v, p_in=pr)
    for _ in range(2):                                       # second call replays the cached graph
        gu, gv, gp = run_cavity(nb, c, host=False)
    if "cavity_flow/" + preset in pins_next:
        pin = pins_next["cavity_flow/" + preset]["out"]
        assert (sha(gu), sha(gv), sha(gp)) == (pin["u"]["sha256"], pin["v"]["sha256"], pin["p"]["sha256"])
    oracle.set_threads(8)
    try:
        oracle.cavity_flow(p["nx"], p["ny"], p["nt"], p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"])
    finally:
        oracle.set_threads(1)
    assert_bit_equal(gu, u, "u"); assert_bit_equal(gv, v, "v"); assert_bit_equal(gp, pr, "p")


def run_channel(nb, ny, nx, nit, rho, nu, F, host):
    u, v, p, dx, dy, dt = oracle.init_channel_flow(ny, nx)
    if host:
        sc = nb.channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F)
        return sc, u, v, p
    d = [nb.DeviceArray.from_host(a) for a in (u, v, p)]
    sc = nb.channel_flow(nit, d[0], d[1], dt, dx, dy, d[2], rho, nu, F)
    return sc, d[0].to_host(), d[1].to_host(), d[2].to_host()


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True], ids=["device", "host"])
def test_gpu_channel_flow_golden_cases(nb, cases_next, host):
    for n, c in enumerate(cases_next["channel_flow"]):
        sc, u, v, p = run_channel(nb, int(c["ny"]), int(c["nx"]), int(c["nit"]), float(c["rho"]), float(c["nu"]),
                                  float(c["F"]), host)
        assert sc == int(c["stepcount"]), "channel_flow.%d stepcount" % n
        for name, g in (("u", u), ("v", v), ("p", p)):
            assert_bit_equal(g, c[name + "_out"], "channel_flow.%d %s" % (n, name))


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L", "paper"])
def test_gpu_channel_flow_presets(nb, pins_next, preset):
    p = oracle.PRESETS["channel_flow"][preset]
    sc, gu, gv, gp = run_channel(nb, p["ny"], p["nx"], p["nit"], p["rho"], p["nu"], p["F"], host=False)
    if "channel_flow/" + preset in pins_next:
        pin = pins_next["channel_flow/" + preset]
        assert sc == pin["stepcount"]
        assert (sha(gu), sha(gv), sha(gp)) == tuple(pin["out"][k]["sha256"] for k in ("u", "v", "p"))
    u, v, pr, dx, dy, dt = oracle.init_channel_flow(p["ny"], p["nx"])
    oracle.set_threads(8)
    try:
        sc_ref = oracle.channel_flow(p["nit"], u, v, dt, dx, dy, pr, p["rho"], p["nu"], p["F"])
    finally:
        oracle.set_threads(1)
    assert sc == sc_ref
    assert_bit_equal(gu, u, "u"); assert_bit_equal(gv, v, "v"); assert_bit_equal(gp, pr, "p")


@pytest.mark.gpu
def test_gpu_adi_zero_steps_raises_like_the_reference(nb):
    with pytest.raises(ZeroDivisionError):          # DT = 1.0 / TSTEPS, adi_numpy.py:14
        nb.adi(0, 8, np.zeros((8, 8)))


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True], ids=["device", "host"])
def test_gpu_golden_cases(nb, cases_next, host):
    for n, c in enumerate(cases_next["adi"]):
        g = run_adi(nb, int(c["TSTEPS"]), c["u_in"], host)
        assert_bit_equal(g, c["u_out"], "adi.%d" % n)
    for n, c in enumerate(cases_next["jacobi_1d"]):
        A, B = run_j1(nb, int(c["TSTEPS"]), c["A_in"], c["B_in"], host)
        assert_bit_equal(A, c["A_out"], "jacobi_1d.%d A" % n); assert_bit_equal(B, c["B_out"], "jacobi_1d.%d B" % n)
    for n, c in enumerate(cases_next["seidel_2d"]):
        A = run_s2(nb, int(c["TSTEPS"]), c["A_in"], host)
        assert_bit_equal(A, c["A_out"], "seidel_2d.%d" % n)


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L", "paper"])
def test_gpu_jacobi_1d_presets(nb, pins_next, preset):
    p = oracle.PRESETS["jacobi_1d"][preset]
    A, B = oracle.init_jacobi_1d(p["N"])
    gA, gB = run_j1(nb, p["TSTEPS"], A, B, host=False)
    if "jacobi_1d/" + preset in pins_next:
        pin = pins_next["jacobi_1d/" + preset]["out"]
        assert sha(gA) == pin["A"]["sha256"] and sha(gB) == pin["B"]["sha256"]
    oracle.set_threads(4)
    try:
        oracle.jacobi_1d(p["TSTEPS"], A, B)
    finally:
        oracle.set_threads(1)
    assert_bit_equal(gA, A, "A"); assert_bit_equal(gB, B, "B")


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L", "paper"])
def test_gpu_seidel_2d_presets(nb, pins_next, preset):
    p = oracle.PRESETS["seidel_2d"][preset]
    A = oracle.init_seidel_2d(p["N"])
    gA = run_s2(nb, p["TSTEPS"], A, host=False)
    if "seidel_2d/" + preset in pins_next:
        assert sha(gA) == pins_next["seidel_2d/" + preset]["out"]["A"]["sha256"]
    oracle.seidel_2d(p["TSTEPS"], p["N"], A)
    assert_bit_equal(gA, A, "A")


@pytest.mark.gpu
@pytest.mark.parametrize("ts,n", [(1, 10), (2, 3), (2, 2), (3, 1), (9, 31), (17, 32), (64, 33), (200, 777), (129, 4096),
                                  (1000, 100), (4, 100000)])
def test_gpu_jacobi_1d_random(nb, ts, n):
    rng = np.random.default_rng(ts + n)
    A, B = rng.random((n,)) - 0.5, rng.random((n,)) - 0.5
    gA, gB = run_j1(nb, ts, A, B, host=False)
    oracle.jacobi_1d(ts, A, B)
    assert_bit_equal(gA, A, "A"); assert_bit_equal(gB, B, "B")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["dispatch", "l2"])
@pytest.mark.parametrize("ts,n", [(1, 10), (2, 3), (2, 2), (2, 5), (3, 6), (7, 40), (30, 17), (5, 129), (3, 300), (12, 64),
                                  (9, 24), (6, 25), (4, 440), (3, 470), (20, 97), (3, 500)])
def test_gpu_seidel_2d_random(nb, ts, n, mode):
    """mode 0: distributed-shared-memory cluster kernel when 24 <= N and the grid fits (else L2 kernel); 1: L2 kernel."""
    rng = np.random.default_rng(ts * 100 + n)
    A = rng.random((n, n)) - 0.5
    nb.lib().seidel2d_set_mode(mode)
    try:
        gA = run_s2(nb, ts, A, host=False)
        path = nb.lib().seidel2d_last_path()
    finally:
        nb.lib().seidel2d_set_mode(0)
    if ts > 1 and n >= 3:
        fits = n >= 24 and ((n + 7) // 8) * n * 8 + 1024 <= 232448          # one 8-CTA cluster's shared memory
        assert path == (1 if (mode == 0 and fits) else 2)
    oracle.seidel_2d(ts, n, A)
    assert_bit_equal(gA, A, "A")
