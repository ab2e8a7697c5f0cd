"""The CPU oracle (oracle/stencil_oracle.c) against the reference's own outputs.

Golden data comes from tests/golden/make_golden.py, which imports the unmodified
reference NumPy kernels (jacobi_2d_numpy.py:4-10, heat_3d_numpy.py:4-20,
fdtd_2d_numpy.py:4-11, hdiff_numpy.py:5-29, vadv_numpy.py:9-78).  Bar: bit-exact.
"""
import numpy as np
import pytest

import oracle
from conftest import assert_bit_equal, sha


@pytest.fixture(scope="module", params=[1, 4], ids=["1thread", "4threads"])
def threads(request):
    oracle.set_threads(request.param)
    yield request.param
    oracle.set_threads(1)


# SURVEY.md section 4 pins (sha256[:16]) recorded independently during the survey
SURVEY_PINS = {
    ("jacobi_2d/S", "A"): "6fa8fb2fe9393cf5", ("jacobi_2d/S", "B"): "c99510e93631f61d",
    ("jacobi_2d/M", "A"): "a81d6fcf82a55e56", ("jacobi_2d/M", "B"): "7cf15549b1e2ead4",
    ("heat_3d/S", "A"): "0d79ad24bc16bcd9", ("heat_3d/M", "A"): "3dd9377c24ce238b",
    ("fdtd_2d/S", "ex"): "0abccf3ddd1e1033", ("fdtd_2d/S", "ey"): "436b6e6b347b2347",
    ("fdtd_2d/S", "hz"): "5f48442929d715e6", ("fdtd_2d/M", "ex"): "2430501880bb36d0",
    ("fdtd_2d/M", "ey"): "8814a051915f07c1", ("fdtd_2d/M", "hz"): "4ce37870161d30f9",
    ("hdiff/S", "out_field"): "5e120f438b2e1001", ("hdiff/M", "out_field"): "8499c41055706463",
    ("vadv/S", "utens_stage"): "d1ae5ea50a65b4fd", ("vadv/M", "utens_stage"): "e420e9b2387e2cb4",
}


def test_pins_file_matches_survey_pins(pins):
    for (key, field), h in SURVEY_PINS.items():
        assert pins[key]["out"][field]["sha256"].startswith(h), (key, field)


def _check(pin, name, arr, kind):
    assert sha(arr) == pin[kind][name]["sha256"], "%s %s" % (kind, name)


@pytest.mark.parametrize("preset", ["S", "M"])
def test_presets(pins, threads, preset):
    p = oracle.PRESETS["jacobi_2d"][preset]; pin = pins["jacobi_2d/" + preset]
    A, B = oracle.init_jacobi_2d(p["N"])
    _check(pin, "A", A, "in"); _check(pin, "B", B, "in")
    oracle.jacobi_2d(p["TSTEPS"], A, B)
    _check(pin, "A", A, "out"); _check(pin, "B", B, "out")

    p = oracle.PRESETS["heat_3d"][preset]; pin = pins["heat_3d/" + preset]
    A, B = oracle.init_heat_3d(p["N"])
    _check(pin, "A", A, "in"); _check(pin, "B", B, "in")
    oracle.heat_3d(p["TSTEPS"], A, B)
    _check(pin, "A", A, "out"); _check(pin, "B", B, "out")

    p = oracle.PRESETS["fdtd_2d"][preset]; pin = pins["fdtd_2d/" + preset]
    ex, ey, hz, fict = oracle.init_fdtd_2d(p["TMAX"], p["NX"], p["NY"])
    for n, a in (("ex", ex), ("ey", ey), ("hz", hz), ("_fict_", fict)):
        _check(pin, n, a, "in")
    oracle.fdtd_2d(p["TMAX"], ex, ey, hz, fict)
    for n, a in (("ex", ex), ("ey", ey), ("hz", hz)):
        _check(pin, n, a, "out")

    p = oracle.PRESETS["hdiff"][preset]; pin = pins["hdiff/" + preset]
    inf, outf, coeff = oracle.init_hdiff(p["I"], p["J"], p["K"])
    for n, a in (("in_field", inf), ("out_field", outf), ("coeff", coeff)):
        _check(pin, n, a, "in")
    oracle.hdiff(inf, outf, coeff)
    _check(pin, "out_field", outf, "out")

    p = oracle.PRESETS["vadv"][preset]; pin = pins["vadv/" + preset]
    dtr, us, u, w, up, ut = oracle.init_vadv(p["I"], p["J"], p["K"])
    assert dtr == pin["dtr_stage"]
    for n, a in (("utens_stage", us), ("u_stage", u), ("wcon", w), ("u_pos", up), ("utens", ut)):
        _check(pin, n, a, "in")
    oracle.vadv(us, u, w, up, ut, dtr)
    _check(pin, "utens_stage", us, "out")


def test_small_cases(cases, threads):
    for n, c in enumerate(cases["jacobi_2d"]):
        A, B = c["A_in"].copy(), c["B_in"].copy()
        oracle.jacobi_2d(int(c["TSTEPS"]), A, B)
        assert_bit_equal(A, c["A_out"], "jacobi_2d.%d A" % n); assert_bit_equal(B, c["B_out"], "jacobi_2d.%d B" % n)
    for n, c in enumerate(cases["heat_3d"]):
        A, B = c["A_in"].copy(), c["B_in"].copy()
        oracle.heat_3d(int(c["TSTEPS"]), A, B)
        assert_bit_equal(A, c["A_out"], "heat_3d.%d A" % n); assert_bit_equal(B, c["B_out"], "heat_3d.%d B" % n)
    for n, c in enumerate(cases["fdtd_2d"]):
        ex, ey, hz = c["ex_in"].copy(), c["ey_in"].copy(), c["hz_in"].copy()
        oracle.fdtd_2d(int(c["TMAX"]), ex, ey, hz, c["fict"].copy())
        for name, a in (("ex", ex), ("ey", ey), ("hz", hz)):
            assert_bit_equal(a, c[name + "_out"], "fdtd_2d.%d %s" % (n, name))
    for n, c in enumerate(cases["hdiff"]):
        out = c["out_in"].copy()
        oracle.hdiff(c["in_field"].copy(), out, c["coeff"].copy())
        assert_bit_equal(out, c["out_field"], "hdiff.%d" % n)
    for n, c in enumerate(cases["vadv"]):
        us = c["utens_stage_in"].copy()
        oracle.vadv(us, c["u_stage"].copy(), c["wcon"].copy(), c["u_pos"].copy(), c["utens"].copy(),
                    float(c["dtr_stage"]))
        assert_bit_equal(us, c["utens_stage_out"], "vadv.%d" % n)


def test_sweeps_helper_equals_kernel():
    rng = np.random.default_rng(3)
    A, B = rng.random((19, 23)), rng.random((19, 23))
    A2, B2 = A.copy(), B.copy()
    oracle.jacobi_2d(4, A, B)
    oracle.jacobi_2d_sweeps(6, A2, B2)
    assert_bit_equal(A2, A); assert_bit_equal(B2, B)
    A, B = rng.random((7, 8, 9)), rng.random((7, 8, 9))
    A2, B2 = A.copy(), B.copy()
    oracle.heat_3d(3, A, B)
    oracle.heat_3d_sweeps(4, A2, B2)
    assert_bit_equal(A2, A); assert_bit_equal(B2, B)


def test_slab_initialisers_match_full():
    A, B = oracle.init_jacobi_2d(40)
    a, b = oracle.init_jacobi_2d(40, row0=7, nrows=9)
    assert_bit_equal(a, A[7:16]); assert_bit_equal(b, B[7:16])
    A, B = oracle.init_heat_3d(12)
    a, b = oracle.init_heat_3d(12, row0=3, nrows=4)
    assert_bit_equal(a, A[3:7]); assert_bit_equal(b, B[3:7])
    ex, ey, hz, f = oracle.init_fdtd_2d(5, 20, 30)
    x, y, z, _ = oracle.init_fdtd_2d(5, 20, 30, row0=11, nrows=5)
    assert_bit_equal(x, ex[11:16]); assert_bit_equal(y, ey[11:16]); assert_bit_equal(z, hz[11:16])


def test_large_presets_against_reference_pins(pins_large):
    """Presets L and `paper` (tests/golden/make_golden_large.py: sha256 of the unmodified reference's outputs;
    jacobi_2d paper is a 160 s NumPy run there, 5 s here).  heat_3d also on seeded random inputs, because its
    NPBench input is a fixed point of the stencil."""
    oracle.set_threads(oracle.max_threads())
    try:
        for preset in ("L", "paper"):
            p = oracle.PRESETS["jacobi_2d"][preset]; pin = pins_large["jacobi_2d/" + preset]
            A, B = oracle.init_jacobi_2d(p["N"])
            oracle.jacobi_2d(p["TSTEPS"], A, B)
            _check(pin, "A", A, "out"); _check(pin, "B", B, "out")

            p = oracle.PRESETS["heat_3d"][preset]; pin = pins_large["heat_3d/" + preset]
            A, B = oracle.init_heat_3d(p["N"])
            oracle.heat_3d(p["TSTEPS"], A, B)
            _check(pin, "A", A, "out"); _check(pin, "B", B, "out")
            pin = pins_large["heat_3d_random/" + preset]
            rng = np.random.default_rng(pins_large["heat_3d_random_seed"])
            A = rng.random((p["N"],) * 3); B = rng.random((p["N"],) * 3)
            _check(pin, "A", A, "in"); _check(pin, "B", B, "in")
            oracle.heat_3d(p["TSTEPS"], A, B)
            _check(pin, "A", A, "out"); _check(pin, "B", B, "out")

            p = oracle.PRESETS["fdtd_2d"][preset]; pin = pins_large["fdtd_2d/" + preset]
            ex, ey, hz, fict = oracle.init_fdtd_2d(p["TMAX"], p["NX"], p["NY"])
            oracle.fdtd_2d(p["TMAX"], ex, ey, hz, fict)
            for n, a in (("ex", ex), ("ey", ey), ("hz", hz)):
                _check(pin, n, a, "out")

            p = oracle.PRESETS["hdiff"][preset]; pin = pins_large["hdiff/" + preset]
            inf, outf, coeff = oracle.init_hdiff(p["I"], p["J"], p["K"])
            _check(pin, "in_field", inf, "in"); _check(pin, "coeff", coeff, "in")
            oracle.hdiff(inf, outf, coeff)
            _check(pin, "out_field", outf, "out")

            p = oracle.PRESETS["vadv"][preset]; pin = pins_large["vadv/" + preset]
            dtr, us, u, w, up, ut = oracle.init_vadv(p["I"], p["J"], p["K"])
            assert dtr == pin["dtr_stage"]
            _check(pin, "utens_stage", us, "in"); _check(pin, "wcon", w, "in")
            oracle.vadv(us, u, w, up, ut, dtr)
            _check(pin, "utens_stage", us, "out")
    finally:
        oracle.set_threads(1)
