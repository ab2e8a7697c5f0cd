"""Column-sharded arrays for hdiff / vadv on several GPUs driven by ONE process.

NPBench's harness is a single process and its reference kernels are single device; hdiff and vadv need no run-time
exchange (split along I with a fixed overlap of 4 `in_field` rows / 1 `wcon` row at copy-in), so the plugin can use
every GPU of the box without torchrun or NCCL: `NPB_B200_GPUS=N` makes B200Framework.setup_str scatter the array
arguments (`scatter`), the `<bench>_b200.py` functions call `npb_{hdiff,vadv}_f64_mg`, and copy_back_func gathers.
A device index may repeat in `devices` (shards sharing one GPU): that is how a one-GPU box tests this path.
"""
import ctypes
import os
from typing import List, Sequence

import numpy as np

from . import _lib
from .device_array import DeviceArray

# extra leading rows a shard of this argument carries beyond its owned rows (hdiff_numpy.py:7-28, vadv_numpy.py:16,33-34)
OVERLAP = {("hdiff", "in_field"): 4, ("vadv", "wcon"): 1}
_DEVICES: List[int] = []


def configure(devices: Sequence[int]) -> List[int]:
    """Make device slot k drive devices[k] (slot 0 = the device npbench_b200.init selected).  Idempotent."""
    global _DEVICES
    devices = [int(d) for d in devices]
    if devices != _DEVICES:
        arr = (ctypes.c_int * len(devices))(*devices)
        _lib.lib().mg_init(len(devices), arr)
        _DEVICES = devices
    return _DEVICES


def devices_from_env() -> List[int]:
    """NPB_B200_GPUS = N (devices 0..N-1, wrapped onto the visible ones) or a comma list of device indices."""
    spec = os.environ.get("NPB_B200_GPUS", "1").strip()
    first = int(os.environ.get("NPB_B200_DEVICE", "0"))
    if "," in spec:
        return [int(x) for x in spec.split(",")]
    n = max(1, int(spec))
    _lib.lib().init(first)
    visible = _visible_devices()
    return [(first + k) % visible for k in range(n)]


def _visible_devices() -> int:
    try:
        import torch
        return max(1, torch.cuda.device_count())
    except Exception:
        return 1


def bounds(n: int, nshards: int) -> List[int]:
    """i_lo[0..nshards]: shard s owns rows [i_lo[s], i_lo[s+1]) (npb_shard_bounds)."""
    L = _lib.lib()
    out = [0]
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    for s in range(nshards):
        L.shard_bounds(int(n), int(nshards), s, ctypes.byref(lo), ctypes.byref(hi))
        out.append(int(hi.value))
    return out


class ShardedArray:
    """An (I, ...) float64 array split along axis 0 over device slots; shard s holds rows
    [i_lo[s], i_lo[s+1] + overlap) on slot s."""
    __slots__ = ("shape", "dtype", "shards", "i_lo", "overlap")

    def __init__(self, shape, shards, i_lo, overlap):
        self.shape, self.dtype = tuple(int(x) for x in shape), np.dtype(np.float64)
        self.shards, self.i_lo, self.overlap = list(shards), list(i_lo), int(overlap)

    @property
    def nshards(self):
        return len(self.shards)

    @classmethod
    def from_host(cls, a, nshards: int, overlap: int = 0) -> "ShardedArray":
        a = DeviceArray._host_f64(a)
        L = _lib.lib()
        rows = a.shape[0] - overlap                       # owned rows (in_field has I + 4, wcon I + 1)
        i_lo = bounds(rows, nshards)
        shards = []
        try:
            for s in range(nshards):
                L.mg_select(s)
                part = a[i_lo[s]:i_lo[s + 1] + overlap]
                d = DeviceArray(part.shape)
                if d.nbytes:
                    L.h2d(d.ptr, part.ctypes.data, d.nbytes)
                shards.append(d)
        finally:
            L.mg_select(0)
        L.sync()                                          # `a` may be a temporary; also what setup_str promises
        return cls(a.shape, shards, i_lo, overlap)

    def to_host(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        L = _lib.lib()
        try:
            for s, d in enumerate(self.shards):
                L.mg_select(s)
                own = self.i_lo[s + 1] - self.i_lo[s] + (self.overlap if s == self.nshards - 1 else 0)
                if own:
                    part = out[self.i_lo[s]:self.i_lo[s] + own]
                    L.d2h(part.ctypes.data, d.ptr, part.nbytes)
        finally:
            L.mg_select(0)
        L.sync()
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_host()
        return a if dtype is None else a.astype(dtype)


def scatter(a, bench: str, arg: str):
    """What B200Framework.setup_str calls per array argument when NPB_B200_GPUS > 1."""
    n = len(configure(devices_from_env()))
    return ShardedArray.from_host(a, n, OVERLAP.get((bench, arg), 0))


def _ptrs(arrays):
    return (ctypes.c_void_p * len(arrays))(*[d.ptr for d in arrays])


def _slots(n):
    return (ctypes.c_int * n)(*range(n))


def hdiff_mg(in_field: ShardedArray, out_field: ShardedArray, coeff: ShardedArray) -> None:
    I, J, K = out_field.shape
    n = out_field.nshards
    if in_field.shape != (I + 4, J + 4, K) or coeff.shape != (I, J, K) or in_field.overlap != 4:
        raise ValueError("expected in_field (I+4,J+4,K) scattered with a 4-row overlap, out_field/coeff (I,J,K)")
    if not (in_field.i_lo == out_field.i_lo == coeff.i_lo):
        raise ValueError("arguments are sharded differently")
    i_lo = (ctypes.c_int64 * (n + 1))(*out_field.i_lo)
    _lib.lib().hdiff_f64_mg(n, _slots(n), I, J, K, _ptrs(in_field.shards), _ptrs(out_field.shards), _ptrs(coeff.shards), i_lo)


def vadv_mg(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage) -> None:
    I, J, K = utens_stage.shape
    n = utens_stage.nshards
    if wcon.shape != (I + 1, J, K) or wcon.overlap != 1:
        raise ValueError("wcon must be (I+1, J, K) scattered with a 1-row overlap")
    for a in (u_stage, u_pos, utens):
        if a.shape != (I, J, K) or a.i_lo != utens_stage.i_lo:
            raise ValueError("u_stage, u_pos, utens must match utens_stage and be sharded alike")
    if K < 2:
        raise IndexError("vadv needs K >= 2 (the reference indexes level k+1 at k = 0)")
    i_lo = (ctypes.c_int64 * (n + 1))(*utens_stage.i_lo)
    _lib.lib().vadv_f64_mg(n, _slots(n), I, J, K, _ptrs(utens_stage.shards), _ptrs(u_stage.shards), _ptrs(wcon.shards),
                           _ptrs(u_pos.shards), _ptrs(utens.shards), float(dtr_stage), i_lo)
