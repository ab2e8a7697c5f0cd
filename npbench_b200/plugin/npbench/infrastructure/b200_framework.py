# NPBench framework plugin for the B200-native stencil backend (libnpb_b200.so).
#
# Drop-in per frameworks.md: this file goes to npbench/infrastructure/, b200.json to
# framework_info/, the <bench>_b200.py modules next to their NumPy siblings, and one
# line `from .b200_framework import *` is appended to npbench/infrastructure/__init__.py
# (frameworks.md:39-42).  Nothing else in the harness changes.
#
# Pattern follows the in-tree GPU plugins (cupy_framework.py:22-58): device copies in
# the untimed setup string, a device synchronisation appended to the timed statement.
from typing import Any, Callable, Dict

from npbench.infrastructure import Benchmark, Framework

import npbench_b200
from npbench_b200 import DeviceArray, ShardedArray
from npbench_b200 import sharded as _sharded

__all__ = ["B200Framework"]

_SYNC = "__npb_b200_sync()"


def _to_device(a):
    """copy_func: called as __npb_copy(<array>) once per array_arg in setup_str."""
    return DeviceArray.from_host(a)


def _to_host(a):
    """copy_back_func: outputs back to NumPy for util.validate (test.py:107-110); gathers column shards."""
    return a.to_host() if isinstance(a, (DeviceArray, ShardedArray)) else a


# benchmarks whose array arguments are scattered over NPB_B200_GPUS devices (no run-time exchange needed)
_SHARDED_BENCHES = {"hdiff", "vadv"}


class B200Framework(Framework):
    """Framework subclass for `-f b200` (framework.py:12-162 is the interface)."""

    def __init__(self, fname: str):
        super().__init__(fname)
        import os
        # fails loudly when no B200 is visible (no CPU fallback).  NPB_B200_GPUS = N > 1: hdiff and vadv are split
        # along I over N device slots of this one process (npbench_b200/sharded.py); everything else runs on device 0
        npbench_b200.init(int(os.environ.get("NPB_B200_DEVICE", "0")))
        self.devices = _sharded.configure(_sharded.devices_from_env())

    def version(self) -> str:
        # the base class asks pkg_resources for a distribution called "b200" (framework.py:33-35)
        return "%s+lib%s" % (npbench_b200.__version__, npbench_b200.lib().version().decode())

    def imports(self) -> Dict[str, Any]:
        # merged into the exec namespace (test.py:90)
        return {"__npb_b200_sync": npbench_b200.sync, "__npb_b200_scatter": _sharded.scatter}

    def copy_func(self) -> Callable:
        return _to_device

    def copy_back_func(self) -> Callable:
        return _to_host

    def setup_str(self, bench: Benchmark, impl: Callable = None) -> str:
        # H2D copies finish before the timer starts (cupy_framework.py:32-45)
        name = bench.info["module_name"]
        if len(self.devices) > 1 and name in _SHARDED_BENCHES and len(bench.info["array_args"]):
            # copy_func only sees the array (framework.py:149), but in_field / wcon need their row overlap: emit one
            # scatter call per array argument that names the benchmark and the argument (framework.py:139-150)
            arg_str = self.out_arg_str(bench, impl)
            calls = ", ".join("__npb_b200_scatter({a}, '{b}', '{a}')".format(a=a, b=name) for a in bench.info["array_args"])
            return arg_str + " = " + calls + "; " + _SYNC
        base = super().setup_str(bench, impl)
        return _SYNC if base == "pass" else base + "; " + _SYNC

    def exec_str(self, bench: Benchmark, impl: Callable = None) -> str:
        # kernels are enqueued asynchronously; the timed statement ends with a device sync
        # (cupy_framework.py:47-58)
        return super().exec_str(bench, impl) + "; " + _SYNC
