"""DeviceArray: the object B200Framework.copy_func returns to the NPBench harness.

A C-contiguous float64 array living in B200 HBM, owned by the library's
caching allocator (npb_malloc / npb_free).  It carries `shape`/`dtype` because
the NPBench kernels derive their extents from the arrays
(hdiff_numpy.py:6, vadv_numpy.py:10).
"""
import ctypes

import numpy as np

from . import _lib


class DeviceArray:
    __slots__ = ("ptr", "shape", "dtype", "nbytes", "_owner")

    def __init__(self, shape, dtype=np.float64, _ptr=None, _owner=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        if self.dtype != np.float64:
            raise TypeError("the B200 stencil backend computes in float64 only")
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        if _ptr is None:
            p = ctypes.c_void_p()
            _lib.lib().malloc(max(self.nbytes, 1), ctypes.byref(p))
            self.ptr = p.value
            self._owner = None
        else:                       # a view into memory owned by someone else
            self.ptr = int(_ptr)
            self._owner = _owner if _owner is not None else True

    # -- construction / extraction -------------------------------------------
    @staticmethod
    def _host_f64(a) -> np.ndarray:
        """C-contiguous float64 view/copy of `a`.  Other dtypes are REJECTED, not upcast: the harness
        validates against NumPy run on the same arrays, so a silent cast would compare different
        precisions (kernels._kind rejects the same input on the host-buffer path)."""
        a = np.asarray(a)
        if a.dtype != np.float64:
            raise TypeError("the B200 stencil backend computes in float64 only (got %s); NPBench's default "
                            "datatype for these kernels is float64" % a.dtype)
        return np.ascontiguousarray(a)

    @classmethod
    def from_host(cls, a) -> "DeviceArray":
        """Framework.copy_func: np.ndarray -> device (async on the library stream)."""
        a = cls._host_f64(a)
        d = cls(a.shape)
        if d.nbytes:
            _lib.lib().h2d(d.ptr, a.ctypes.data, d.nbytes)
            _lib.lib().sync()       # `a` may be a temporary: finish before it can be freed
        return d

    def to_host(self) -> np.ndarray:
        """Framework.copy_back_func: device -> np.ndarray."""
        out = np.empty(self.shape, dtype=np.float64)
        if self.nbytes:
            _lib.lib().d2h(out.ctypes.data, self.ptr, self.nbytes)
        _lib.lib().sync()
        return out

    def copy_from_host(self, a) -> None:
        a = self._host_f64(a)
        assert a.shape == self.shape
        if self.nbytes:
            _lib.lib().h2d(self.ptr, a.ctypes.data, self.nbytes)
            _lib.lib().sync()

    @property
    def size(self):
        return self.nbytes // 8

    @property
    def ndim(self):
        return len(self.shape)

    def __array__(self, dtype=None, copy=None):
        a = self.to_host()
        return a if dtype is None else a.astype(dtype)

    def __del__(self):
        try:
            if getattr(self, "_owner", True) is None and self.ptr:
                _lib.lib()._raw_npb_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass
