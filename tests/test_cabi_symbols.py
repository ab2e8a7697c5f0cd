"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/npb_b200.h declares, the ctypes prototypes cover exactly that set, and
the product path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "npb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(npb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_header_symbol():
    from npbench_b200 import build
    path = build.build()
    cdll = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(cdll, s), "libnpb_b200.so does not export %s" % s


def test_ctypes_prototypes_match_header():
    from npbench_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_symbols()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import npbench_b200 as nb
    A = np.zeros((8, 8)); B = np.zeros((8, 8))
    with pytest.raises(nb.B200Error, match="no CUDA device"):
        nb.jacobi_2d(3, A, B)
    with pytest.raises(nb.B200Error):
        nb.DeviceArray.from_host(A)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "npbench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_argument_validation_mirrors_numpy_errors():
    import npbench_b200 as nb
    with pytest.raises(ValueError):
        nb.jacobi_2d(3, np.zeros((4, 4)), np.zeros((4, 5)))
    with pytest.raises(ValueError):
        nb.hdiff(np.zeros((8, 8, 3)), np.zeros((4, 4, 3)), np.zeros((4, 5, 3)))
    with pytest.raises(IndexError):
        nb.vadv(*(np.zeros((2, 2, 1)),) * 2, np.zeros((3, 2, 1)), *(np.zeros((2, 2, 1)),) * 2, 0.15)
    with pytest.raises(TypeError):
        nb.heat_3d(2, np.zeros((4, 4, 4), dtype=np.float32), np.zeros((4, 4, 4), dtype=np.float32))


def test_fdtd2d_pass_plan_host_logic():
    """npb_fdtd2d_pass_plan is pure host logic (no device call): the marching time loop must cover TMAX exactly,
    with at most five steps per pass, evenly spread, and an even number of passes whenever TMAX >= 2 (the result then
    ends in the caller's arrays without a copy-back); the per-step loop is TMAX passes of one step."""
    import ctypes

    import numpy as np

    from npbench_b200 import _lib

    L = _lib.lib()
    buf = np.zeros(4096, dtype=np.int32)
    ptr = ctypes.c_void_p(buf.ctypes.data)
    assert L.fdtd2d_pass_plan(0, 1, ptr, buf.size) == 0
    for tmax in list(range(1, 70)) + [150, 499, 500, 1000]:
        n = L.fdtd2d_pass_plan(tmax, 1, ptr, buf.size)
        plan = buf[:n].tolist()
        assert sum(plan) == tmax and min(plan) >= 1 and max(plan) <= 5
        assert max(plan) - min(plan) <= 1
        assert n == 1 if tmax == 1 else n % 2 == 0
        assert n <= (tmax + 4) // 5 + 1
        m = L.fdtd2d_pass_plan(tmax, 0, ptr, buf.size)
        assert m == tmax and buf[:m].tolist() == [1] * tmax
    # cap smaller than the plan: count is still returned, only `cap` entries are written
    buf[:] = -1
    assert L.fdtd2d_pass_plan(50, 1, ptr, 3) == 10 and buf[:4].tolist() == [5, 5, 5, -1]


def test_jacobi2d_pass_plan_host_logic_matches_the_slab_driver():
    """npb_jacobi2d_pass_plan is pure host logic: the passes of the single-device call in the marching regime are the
    slab driver's (distributed.jacobi_plan / jacobi_plan_dual) -- same sizes, same order -- so that N = 1 and N > 1 run
    the same plan, and they cover 2 (TSTEPS - 1) sweeps with odd passes of at most 7 sweeps."""
    import ctypes

    import numpy as np

    from npbench_b200 import _lib
    from npbench_b200 import distributed as D

    L = _lib.lib()
    buf = np.zeros(4096, dtype=np.int32)
    ptr = ctypes.c_void_p(buf.ctypes.data)
    assert L.jacobi2d_pass_plan(1, 1, ptr, buf.size) == 0
    for ts in list(range(2, 80)) + [500, 1000]:
        S = 2 * (ts - 1)
        n = L.jacobi2d_pass_plan(ts, 0, ptr, buf.size)
        assert buf[:n].tolist() == D.jacobi_plan(S), ts
        n = L.jacobi2d_pass_plan(ts, 1, ptr, buf.size)
        plan = buf[:n].tolist()
        assert sum(plan) == S and all(p % 2 == 1 and 1 <= p <= 7 for p in plan), ts
        if S >= 4:
            assert plan == D.jacobi_plan_dual(S) and n % 2 == 0 and plan[-1] >= 3, ts
        else:
            assert plan == D.jacobi_plan(S), ts           # two sweeps: nothing to gain from a scratch grid
    buf[:] = -1
    assert L.jacobi2d_pass_plan(21, 1, ptr, 2) == 6 and buf[:3].tolist() == [5, 7, -1]


def test_jacobi2d_march_chunk_rule_host_logic():
    """npb_jacobi2d_march_rows_per_chunk (pure host logic): short chunks (~256 warps per SM over a launch, at least 96
    rows, 64 on small grids), never more rows than the launch has, never more than 65535 chunks."""
    from npbench_b200 import _lib

    L = _lib.lib()
    f = L.jacobi2d_march_rows_per_chunk
    assert f(2, 100, 100, 148) == 0 and f(7, 0, 100, 148) == 0
    assert f(7, 10238, 81920, 148) == 197          # the bench slab: 52 chunks of 197 rows
    assert f(7, 16382, 16384, 148) == 96           # 16384^2: the 96-row floor
    assert f(7, 4094, 4096, 148) == 64             # 4096^2: too few strips for 96-row chunks
    assert f(5, 40, 100000, 148) == 40             # a boundary range of the slab driver: one chunk
    for ns in (1, 3, 5, 7):
        for rows in (1, 63, 64, 200, 5000, 10 ** 6, 7 * 10 ** 6, 5 * 10 ** 7):
            for nj in (8, 130, 4096, 81920, 10 ** 6):
                for sms in (1, 132, 148):
                    rc = f(ns, rows, nj, sms)
                    assert 1 <= rc <= rows
                    assert (rows + rc - 1) // rc <= 65535
                    assert rc >= min(rows, 64)
