"""npbench_b200 -- B200 (sm_100a) backend for NPBench's structured-grid stencil family.

    from npbench_b200 import jacobi_2d, heat_3d, fdtd_2d, hdiff, vadv, DeviceArray

The compute lives in libnpb_b200.so (hand-written CUDA, C ABI: include/npb_b200.h);
this package is the thin ctypes host side plus the NPBench plugin files
(npbench_b200/plugin/).  There is no CPU fallback.
"""
from ._lib import B200Error, init, lib  # noqa: F401
from .device_array import DeviceArray  # noqa: F401
from .sharded import ShardedArray  # noqa: F401
from .kernels import adi, cavity_flow, channel_flow, fdtd_2d, hdiff, heat_3d, jacobi_1d, jacobi_2d, seidel_2d, sync, vadv  # noqa: F401

__version__ = "0.1.0"
