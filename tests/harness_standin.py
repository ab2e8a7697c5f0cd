"""Stand-in for the part of the NPBench harness that drives a framework plugin.

TEST INFRASTRUCTURE.  The GPU box has no NPBench checkout, so the plugin files under
npbench_b200/plugin/ are exercised against this minimal re-statement of the interface
they plug into.  It reproduces, in our own words, only the behaviour the plugin relies on:
  * `Framework` base: reads framework_info/<name>.json next to the plugin tree and builds
    the argument names / setup / exec strings (framework.py:15-31, 89-162);
  * `Benchmark`: the bench_info facts of the five stencils (input/array/output args);
  * `execute`: what Test._execute + utilities.benchmark do with those strings
    (test.py:16-51, utilities.py:135-151): timeit with setup untimed, then one extra run
    whose namespace yields the outputs.
`install()` registers it as `npbench.infrastructure` so that the plugin's own
`from npbench.infrastructure import Benchmark, Framework` resolves.
"""
import importlib
import json
import os
import sys
import timeit
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "npbench_b200", "plugin")

BENCH_INFO = {   # bench_info/<name>.json: relative_path, module_name, func_name, args
    "jacobi_2d": dict(short_name="jacobi2d", relative_path="polybench/jacobi_2d", module_name="jacobi_2d",
                      func_name="kernel", input_args=["TSTEPS", "A", "B"], array_args=["A", "B"],
                      output_args=["A", "B"]),
    "heat_3d": dict(short_name="heat3d", relative_path="polybench/heat_3d", module_name="heat_3d",
                    func_name="kernel", input_args=["TSTEPS", "A", "B"], array_args=["A", "B"],
                    output_args=["A", "B"]),
    "fdtd_2d": dict(short_name="fdtd_2d", relative_path="polybench/fdtd_2d", module_name="fdtd_2d",
                    func_name="kernel", input_args=["TMAX", "ex", "ey", "hz", "_fict_"],
                    array_args=["ex", "ey", "hz", "_fict_"], output_args=["ex", "ey", "hz"]),
    "hdiff": dict(short_name="hdiff", relative_path="weather_stencils/hdiff", module_name="hdiff",
                  func_name="hdiff", input_args=["in_field", "out_field", "coeff"],
                  array_args=["in_field", "out_field", "coeff"], output_args=["out_field"]),
    # widening row: bench_info/{jacobi_1d,seidel_2d}.json
    "jacobi_1d": dict(short_name="jacobi1d", relative_path="polybench/jacobi_1d", module_name="jacobi_1d",
                      func_name="kernel", input_args=["TSTEPS", "A", "B"], array_args=["A", "B"],
                      output_args=["A", "B"]),
    "seidel_2d": dict(short_name="seidel2d", relative_path="polybench/seidel_2d", module_name="seidel_2d",
                      func_name="kernel", input_args=["TSTEPS", "N", "A"], array_args=["A"], output_args=["A"]),
    "adi": dict(short_name="adi", relative_path="polybench/adi", module_name="adi", func_name="kernel",
                input_args=["TSTEPS", "N", "u"], array_args=["u"], output_args=["u"]),
    "cavity_flow": dict(short_name="cavtflow", relative_path="cavity_flow", module_name="cavity_flow",
                        func_name="cavity_flow",
                        input_args=["nx", "ny", "nt", "nit", "u", "v", "dt", "dx", "dy", "p", "rho", "nu"],
                        array_args=["u", "v", "p"], output_args=["u", "v", "p"]),
    "channel_flow": dict(short_name="chanflow", relative_path="channel_flow", module_name="channel_flow",
                         func_name="channel_flow",
                         input_args=["nit", "u", "v", "dt", "dx", "dy", "p", "rho", "nu", "F"],
                         array_args=["u", "v", "p"], output_args=["u", "v", "p"]),
    "vadv": dict(short_name="vadv", relative_path="weather_stencils/vadv", module_name="vadv", func_name="vadv",
                 input_args=["utens_stage", "u_stage", "wcon", "u_pos", "utens", "dtr_stage"],
                 array_args=["utens_stage", "u_stage", "wcon", "u_pos", "utens"], output_args=["utens_stage"]),
}


class Benchmark:
    def __init__(self, bname):
        self.bname = bname
        self.info = BENCH_INFO[bname]


class Framework:
    def __init__(self, fname):
        self.fname = fname
        with open(os.path.join(PLUGIN, "framework_info", fname + ".json")) as f:
            self.info = json.load(f)["framework"]

    def version(self):
        raise RuntimeError("no distribution named %s" % self.fname)   # what pkg_resources would do

    def imports(self):
        return {}

    def copy_func(self):
        import numpy
        return numpy.copy

    def copy_back_func(self):
        return lambda x: x

    def _dev(self, a):
        return "__npb_%s_%s" % (self.info["prefix"], a)

    def implementations(self, bench):
        mod = importlib.import_module("npbench.benchmarks.%s.%s_%s" % (
            bench.info["relative_path"].replace("/", "."), bench.info["module_name"], self.info["postfix"]))
        return [(getattr(mod, bench.info["func_name"]), "default")]

    def args(self, bench, impl=None):
        return [self._dev(a) if a in bench.info["array_args"] else a for a in bench.info["input_args"]]

    def inout_args(self, bench, impl=None):
        return [self._dev(a) for a in bench.info["output_args"]]

    def setup_str(self, bench, impl=None):
        arrs = bench.info["array_args"]
        if not arrs:
            return "pass"
        return ", ".join(self._dev(a) for a in arrs) + " = " + ", ".join("__npb_copy(%s)" % a for a in arrs)

    def exec_str(self, bench, impl=None):
        return "__npb_result = __npb_impl(%s)" % ", ".join(self.args(bench, impl))


def install():
    """Expose the stand-in as `npbench.infrastructure` and the plugin tree as `npbench.benchmarks`."""
    if "npbench.infrastructure" in sys.modules and not getattr(sys.modules["npbench.infrastructure"], "_standin", False):
        raise RuntimeError("a real npbench is already imported")
    pkg = types.ModuleType("npbench")
    pkg.__path__ = [os.path.join(PLUGIN, "npbench")]
    infra = types.ModuleType("npbench.infrastructure")
    infra.__path__ = [os.path.join(PLUGIN, "npbench", "infrastructure")]
    infra.Benchmark, infra.Framework, infra._standin = Benchmark, Framework, True
    pkg.infrastructure = infra
    sys.modules["npbench"] = pkg
    sys.modules["npbench.infrastructure"] = infra
    mod = importlib.import_module("npbench.infrastructure.b200_framework")
    infra.B200Framework = mod.B200Framework
    return infra


def execute(frmwrk, bench, impl, bdata, repeat=1):
    """Test._execute + utilities.benchmark: returns (outputs, times)."""
    ctx = {"__npb_impl": impl, "__npb_copy": frmwrk.copy_func(), **bdata, **frmwrk.imports()}
    setup, stmt = frmwrk.setup_str(bench, impl), frmwrk.exec_str(bench, impl)
    times = timeit.repeat(stmt, setup=setup, repeat=repeat, number=1, globals={**ctx})
    exec(setup, ctx)
    exec(stmt, ctx)
    res = ctx["__npb_result"]
    out = [] if res is None else (list(res) if isinstance(res, (tuple, list)) else [res])
    out += [ctx[a] for a in frmwrk.inout_args(bench)]
    return out, times
