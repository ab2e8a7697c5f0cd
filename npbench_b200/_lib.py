"""ctypes binding of libnpb_b200.so (include/npb_b200.h).

There is NO fallback: if the shared library is missing or no B200 is visible,
every call raises.  Nothing here imports the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPB_B200_LIB", os.path.join(_HERE, "libnpb_b200.so"))

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p
_dbl = ctypes.c_double
_int = ctypes.c_int
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/npb_b200.h declares
PROTOTYPES = {
    "npb_init": (_int, [_int]),
    "npb_shutdown": (_int, []),
    "npb_version": (ctypes.c_char_p, []),
    "npb_last_error": (ctypes.c_char_p, []),
    "npb_device_info": (_int, [ctypes.POINTER(_int)] * 3 + [ctypes.POINTER(_sz)] * 3),
    "npb_set_stream": (_int, [_vp]),
    "npb_get_stream": (_vp, []),
    "npb_sync": (_int, []),
    "npb_mg_init": (_int, [_int, ctypes.POINTER(_int)]),
    "npb_mg_count": (_int, []),
    "npb_mg_select": (_int, [_int]),
    "npb_mg_current": (_int, []),
    "npb_shard_bounds": (_int, [_i64, _int, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "npb_hdiff_f64_mg": (_int, [_int, ctypes.POINTER(_int), _i64, _i64, _i64, ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "npb_vadv_f64_mg": (_int, [_int, ctypes.POINTER(_int), _i64, _i64, _i64] + [ctypes.POINTER(_vp)] * 5 +
                        [_dbl, ctypes.POINTER(_i64)]),
    "npb_hdiff_f64_mg_host": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp]),
    "npb_vadv_f64_mg_host": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _dbl]),
    "npb_malloc": (_int, [_sz, ctypes.POINTER(_vp)]),
    "npb_free": (_int, [_vp]),
    "npb_pool_trim": (_int, []),
    "npb_host_alloc": (_int, [_sz, ctypes.POINTER(_vp)]),
    "npb_host_free": (_int, [_vp]),
    "npb_h2d": (_int, [_vp, _vp, _sz]),
    "npb_d2h": (_int, [_vp, _vp, _sz]),
    "npb_d2d": (_int, [_vp, _vp, _sz]),
    "npb_memset": (_int, [_vp, _int, _sz]),
    "npb_timer_start": (_int, []),
    "npb_timer_stop": (_int, [ctypes.POINTER(ctypes.c_float)]),
    "npb_launch_count": (ctypes.c_uint64, []),
    "npb_l2_flush": (_int, []),
    "npb_jacobi2d_f64": (_int, [_i64, _i64, _i64, _vp, _vp]),
    "npb_jacobi2d_block_f64": (_int, [_int, _i64, _i64, _vp, _vp, _i64, _i64]),
    "npb_jacobi2d_block2_f64": (_int, [_int, _i64, _i64, _vp, _vp, _vp, _i64, _i64]),
    "npb_jacobi2d_block_marches": (_int, [_i64, _i64]),
    "npb_jacobi2d_set_mode": (_int, [_int]),
    "npb_jacobi2d_last_path": (_int, []),
    "npb_jacobi2d_last_passes": (_int, []),
    "npb_jacobi2d_pass_plan": (_int, [_i64, _int, _vp, _int]),
    "npb_jacobi2d_march_rows_per_chunk": (_i64, [_int, _i64, _i64, _int]),
    "npb_jacobi2d_regtile_config": (_int, [_vp]),
    "npb_jacobi2d_regtile_plan": (_int, [_i64, _i64, _i64, _int, _vp]),
    "npb_jacobi2d_tile_rows": (_int, []),
    "npb_heat3d_f64": (_int, [_i64, _i64, _i64, _i64, _vp, _vp]),
    "npb_fdtd2d_set_mode": (_int, [_int]),
    "npb_fdtd2d_last_path": (_int, []),
    "npb_fdtd2d_regtile_config": (_int, [_vp]),
    "npb_fdtd2d_regtile_plan": (_int, [_i64, _i64, _i64, _int, _vp]),
    "npb_fdtd2d_pass_plan": (_int, [_i64, _int, _vp, _int]),
    "npb_heat3d_set_mode": (_int, [_int]),
    "npb_heat3d_last_path": (_int, []),
    "npb_heat3d_set_trace": (_int, [_vp]),
    "npb_heat3d_sweep_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _i64, _i64]),
    "npb_heat3d_march_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _i64, _i64]),
    "npb_fdtd2d_march_f64": (_int, [_int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64]),
    "npb_fdtd2d_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "npb_fdtd2d_step_f64": (_int, [_i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64]),
    "npb_hdiff_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _vp]),
    "npb_hdiff_set_mode": (_int, [_int]),
    "npb_hdiff_last_path": (_int, []),
    "npb_vadv_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _dbl]),
    "npb_vadv_set_mode": (_int, [_int]),
    "npb_vadv_last_path": (_int, []),
    "npb_vadv_set_trace": (_int, [_vp]),
    "npb_jacobi1d_f64": (_int, [_i64, _i64, _vp, _vp]),
    "npb_jacobi1d_f64_host": (_int, [_i64, _i64, _vp, _vp]),
    "npb_seidel2d_f64": (_int, [_i64, _i64, _vp]),
    "npb_seidel2d_f64_host": (_int, [_i64, _i64, _vp]),
    "npb_seidel2d_set_mode": (_int, [_int]),
    "npb_seidel2d_last_path": (_int, []),
    "npb_cavity_flow_f64": (_int, [_i64, _i64, _i64, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp, _dbl, _dbl]),
    "npb_cavity_flow_f64_host": (_int, [_i64, _i64, _i64, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp, _dbl, _dbl]),
    "npb_channel_flow_f64": (_int, [_i64, _i64, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp, _dbl, _dbl, _dbl, _vp]),
    "npb_channel_flow_f64_host": (_int, [_i64, _i64, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp, _dbl, _dbl, _dbl, _vp]),
    "npb_adi_f64": (_int, [_i64, _i64, _vp]),
    "npb_adi_f64_host": (_int, [_i64, _i64, _vp]),
    "npb_jacobi2d_f64_host": (_int, [_i64, _i64, _i64, _vp, _vp]),
    "npb_heat3d_f64_host": (_int, [_i64, _i64, _i64, _i64, _vp, _vp]),
    "npb_fdtd2d_f64_host": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "npb_hdiff_f64_host": (_int, [_i64, _i64, _i64, _vp, _vp, _vp]),
    "npb_vadv_f64_host": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _dbl]),
    "npb_init_jacobi2d_f64": (_int, [_i64, _i64, _i64, _i64, _vp, _vp]),
    "npb_init_heat3d_f64": (_int, [_i64, _i64, _i64, _vp, _vp]),
    "npb_init_fdtd2d_f64": (_int, [_i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
}

_NO_STATUS = {"npb_version", "npb_last_error", "npb_get_stream", "npb_launch_count", "npb_mg_count", "npb_mg_current",
              "npb_jacobi2d_tile_rows", "npb_jacobi2d_regtile_plan", "npb_fdtd2d_regtile_plan", "npb_jacobi2d_last_path", "npb_jacobi2d_last_passes", "npb_jacobi2d_pass_plan", "npb_jacobi2d_march_rows_per_chunk", "npb_jacobi2d_block_marches", "npb_seidel2d_last_path", "npb_heat3d_last_path", "npb_fdtd2d_last_path", "npb_fdtd2d_pass_plan", "npb_hdiff_last_path", "npb_vadv_last_path"}


class B200Error(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise B200Error("libnpb_b200.so not found at %s -- run `python -m npbench_b200.build`; "
                            "there is no CPU fallback" % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.cdll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, "_raw_" + name, fn)
            setattr(self, name[4:], fn if name in _NO_STATUS else self._checked(name, fn))

    def _checked(self, name, fn):
        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise B200Error("%s failed (%d): %s" % (name, rc, self.cdll.npb_last_error().decode()))
            return rc
        call.__name__ = name
        return call


_LIB = None


def lib() -> _Lib:
    """Load (once) and return the bound library."""
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def init(device: int = -1):
    """Select the GPU of this process (idempotent; -1 = keep / device 0)."""
    L = lib()
    L.init(int(device))
    return L
