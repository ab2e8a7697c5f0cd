"""The NPBench plugin surface (framework_info/b200.json, B200Framework, <bench>_b200.py).

CPU part: against the real, unmodified reference harness when /root/reference is mounted
(build container): the overlay is built, `generate_framework("b200")` resolves our class,
the NumPy framework still runs through the same overlay, and `-f b200` fails loudly without
a GPU.  GPU part (no NPBench checkout on the box): the same plugin files driven by the
stand-in harness of tests/harness_standin.py, outputs compared bit-for-bit with the oracle.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT, assert_bit_equal

from npbench_b200 import overlay as _overlay

try:
    REF = _overlay.find_reference()          # $NPBENCH_REF, /root/reference, or the staged baseline/_ref
except FileNotFoundError:
    REF = ""
needs_ref = pytest.mark.skipif(not REF, reason="no NPBench checkout")


def test_descriptor_has_the_keys_the_harness_reads():
    info = json.load(open(os.path.join(ROOT, "npbench_b200/plugin/framework_info/b200.json")))["framework"]
    assert set(info) >= {"simple_name", "full_name", "prefix", "postfix", "class", "arch"}
    assert info["class"] == "B200Framework" and info["postfix"] == "b200" and info["arch"] == "gpu"


def test_plugin_modules_have_the_numpy_signatures():
    import inspect
    import importlib.util
    want = {"polybench/jacobi_2d/jacobi_2d_b200.py": ("kernel", ["TSTEPS", "A", "B"]),
            "polybench/heat_3d/heat_3d_b200.py": ("kernel", ["TSTEPS", "A", "B"]),
            "polybench/fdtd_2d/fdtd_2d_b200.py": ("kernel", ["TMAX", "ex", "ey", "hz", "_fict_"]),
            "polybench/jacobi_1d/jacobi_1d_b200.py": ("kernel", ["TSTEPS", "A", "B"]),
            "polybench/seidel_2d/seidel_2d_b200.py": ("kernel", ["TSTEPS", "N", "A"]),
            "polybench/adi/adi_b200.py": ("kernel", ["TSTEPS", "N", "u"]),
            "channel_flow/channel_flow_b200.py": ("channel_flow", ["nit", "u", "v", "dt", "dx", "dy", "p", "rho", "nu", "F"]),
            "cavity_flow/cavity_flow_b200.py": ("cavity_flow", ["nx", "ny", "nt", "nit", "u", "v", "dt", "dx", "dy", "p",
                                                                "rho", "nu"]),
            "weather_stencils/hdiff/hdiff_b200.py": ("hdiff", ["in_field", "out_field", "coeff"]),
            "weather_stencils/vadv/vadv_b200.py": ("vadv", ["utens_stage", "u_stage", "wcon", "u_pos", "utens",
                                                            "dtr_stage"])}
    for rel, (fn, args) in want.items():
        path = os.path.join(ROOT, "npbench_b200/plugin/npbench/benchmarks", rel)
        spec = importlib.util.spec_from_file_location("m", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        assert list(inspect.signature(getattr(m, fn)).parameters) == args


@needs_ref
def test_overlay_runs_unmodified_harness(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    ov = str(tmp_path / "ov")
    r = subprocess.run([sys.executable, "-m", "npbench_b200.run", "--reference", REF, "--overlay", ov, "--",
                        "-b", "jacobi_2d", "-f", "numpy", "-p", "S", "-r", "1"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert "NumPy - default - median" in r.stdout, r.stdout + r.stderr
    # the overlay did not modify the reference and registered our class with one line
    init = open(os.path.join(ov, "npbench/infrastructure/__init__.py")).read()
    ref_init = open(os.path.join(REF, "npbench/infrastructure/__init__.py")).read()
    assert init.startswith(ref_init) and init[len(ref_init):].strip() == "from .b200_framework import *"
    assert os.path.islink(os.path.join(ov, "npbench/infrastructure/framework.py"))
    assert os.path.exists(os.path.join(ov, "framework_info/b200.json"))


@needs_ref
def test_b200_through_real_harness_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "npbench_b200.run", "--reference", REF, "--overlay",
                        str(tmp_path / "ov"), "--", "-b", "hdiff", "-f", "b200", "-p", "S", "-r", "1"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device visible" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("bench", ["jacobi_2d", "heat_3d", "fdtd_2d", "hdiff", "vadv", "jacobi_1d", "seidel_2d", "adi", "cavity_flow", "channel_flow"])
def test_plugin_end_to_end_with_standin_harness(bench):
    import harness_standin as hs
    infra = hs.install()
    frm = infra.B200Framework("b200")
    b = infra.Benchmark(bench)
    assert isinstance(frm.version(), str) and frm.version()
    p = oracle.PRESETS[bench]["S"]
    if bench == "jacobi_2d":
        A, B = oracle.init_jacobi_2d(p["N"]); bdata = dict(TSTEPS=p["TSTEPS"], A=A, B=B)
        ref = lambda d: oracle.jacobi_2d(d["TSTEPS"], d["A"], d["B"])
    elif bench == "heat_3d":
        rng = np.random.default_rng(0)
        bdata = dict(TSTEPS=p["TSTEPS"], A=rng.random((p["N"],) * 3), B=rng.random((p["N"],) * 3))
        ref = lambda d: oracle.heat_3d(d["TSTEPS"], d["A"], d["B"])
    elif bench == "fdtd_2d":
        ex, ey, hz, f = oracle.init_fdtd_2d(p["TMAX"], p["NX"], p["NY"])
        bdata = dict(TMAX=p["TMAX"], ex=ex, ey=ey, hz=hz, _fict_=f)
        ref = lambda d: oracle.fdtd_2d(d["TMAX"], d["ex"], d["ey"], d["hz"], d["_fict_"])
    elif bench == "jacobi_1d":
        A, B = oracle.init_jacobi_1d(p["N"]); bdata = dict(TSTEPS=p["TSTEPS"], A=A, B=B)
        ref = lambda d: oracle.jacobi_1d(d["TSTEPS"], d["A"], d["B"])
    elif bench == "seidel_2d":
        bdata = dict(TSTEPS=p["TSTEPS"], N=p["N"], A=oracle.init_seidel_2d(p["N"]))
        ref = lambda d: oracle.seidel_2d(d["TSTEPS"], d["N"], d["A"])
    elif bench == "cavity_flow":
        u, v, pr, dx, dy, dt = oracle.init_cavity_flow(p["ny"], p["nx"])
        bdata = dict(nx=p["nx"], ny=p["ny"], nt=p["nt"], nit=p["nit"], u=u, v=v, dt=dt, dx=dx, dy=dy, p=pr, rho=p["rho"],
                     nu=p["nu"])
        ref = lambda d: oracle.cavity_flow(d["nx"], d["ny"], d["nt"], d["nit"], d["u"], d["v"], d["dt"], d["dx"], d["dy"],
                                           d["p"], d["rho"], d["nu"])
    elif bench == "channel_flow":
        u, v, pr, dx, dy, dt = oracle.init_channel_flow(p["ny"], p["nx"])
        bdata = dict(nit=p["nit"], u=u, v=v, dt=dt, dx=dx, dy=dy, p=pr, rho=p["rho"], nu=p["nu"], F=p["F"])
        ref = lambda d: oracle.channel_flow(d["nit"], d["u"], d["v"], d["dt"], d["dx"], d["dy"], d["p"], d["rho"], d["nu"],
                                            d["F"])
    elif bench == "adi":
        bdata = dict(TSTEPS=p["TSTEPS"], N=p["N"], u=oracle.init_adi(p["N"]))
        ref = lambda d: oracle.adi(d["TSTEPS"], d["N"], d["u"])
    elif bench == "hdiff":
        i, o, c = oracle.init_hdiff(p["I"], p["J"], p["K"]); bdata = dict(in_field=i, out_field=o, coeff=c)
        ref = lambda d: oracle.hdiff(d["in_field"], d["out_field"], d["coeff"])
    else:
        dtr, us, u, w, up, ut = oracle.init_vadv(p["I"], p["J"], p["K"])
        bdata = dict(dtr_stage=dtr, utens_stage=us, u_stage=u, wcon=w, u_pos=up, utens=ut)
        ref = lambda d: oracle.vadv(d["utens_stage"], d["u_stage"], d["wcon"], d["u_pos"], d["utens"], d["dtr_stage"])
    impl, _ = frm.implementations(b)[0]
    assert "__npb_b200_sync()" in frm.exec_str(b, impl) and "__npb_b200_sync()" in frm.setup_str(b, impl)
    out, times = hs.execute(frm, b, impl, bdata, repeat=3)
    nout = len(b.info["output_args"])
    # adi returns its argument like the reference does (adi_numpy.py:54): returned values come first (test.py:39-50)
    # ... and channel_flow its step count (channel_flow_numpy.py:170)
    assert len(times) == 3 and len(out) == nout + (1 if bench in ("adi", "channel_flow") else 0)
    if bench == "channel_flow":
        assert out[0] == 982                                   # reference step count at preset S (pins_next.json)
    got = [frm.copy_back_func()(a) for a in out[-nout:]]
    want = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in bdata.items()}
    ref(want)
    for name, g in zip(b.info["output_args"], got):
        assert_bit_equal(g, want[name], "%s %s" % (bench, name))
        # the harness's own criterion (utilities.py:154-180) trivially holds too
        assert np.allclose(want[name], g, rtol=1e-5, atol=1e-8)
    # inputs on the host were not touched (copy_func copies, like np.copy)
    for name in b.info["output_args"]:
        assert not np.shares_memory(bdata[name], g)
