"""The sharded drivers with the real CUDA engine.

The driver's round-end GPU test runs on ONE GPU, so two (or three) ranks are emulated by
threads inside one process: each thread runs the unmodified per-rank driver on its own slab
with a loop-back exchanger (device-to-device copies through a mailbox, ordered with CUDA
events) in place of NCCL.  This exercises exactly what NCCL cannot change: the kernels' tile /
row-range launches, ghost-row semantics and the boundary-first / interior-overlap ordering.
With >= 2 GPUs the same check runs under real NCCL via torchrun (bench.py --gpus N and
tools/check_multigpu.py).
"""
import threading

import numpy as np
import pytest
import torch

import oracle
from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


class Mailbox:
    def __init__(self):
        self.cv, self.box = threading.Condition(), {}

    def put(self, key, val):
        with self.cv:
            self.box[key] = val
            self.cv.notify_all()

    def take(self, key):
        with self.cv:
            while key not in self.box:
                if not self.cv.wait(timeout=60):
                    raise TimeoutError("halo %r never arrived" % (key,))
            return self.box.pop(key)


class _Recv:
    def __init__(self, mb, key, dst):
        self.mb, self.key, self.dst = mb, key, dst

    def wait(self):
        data, ev = self.mb.take(self.key)
        ev.wait()                      # current (compute) stream waits for the sender's copy
        self.dst.copy_(data)
        # `data` was allocated on the sender's comm stream: tell the caching allocator that this stream reads it, or the
        # block may be handed out again before the copy above has run (seen under compute-sanitizer, where the GPU
        # lags far behind the host threads)
        data.record_stream(torch.cuda.current_stream())


class LoopbackExchanger:
    """Same contract as distributed.HaloExchanger, transport = same-device copies."""

    def __init__(self, slab, mailbox):
        self.s, self.mb, self.seq = slab, mailbox, 0

    def start(self, fields):
        s, H, reqs = self.s, self.s.H, []
        self.seq += 1
        for fi, f in enumerate(fields):
            sides = []
            if s.ht:
                sides.append((s.rank - 1, f[s.ht:s.ht + H], f[0:s.ht]))
            if s.hb:
                sides.append((s.rank + 1, f[s.nloc - s.hb - H:s.nloc - s.hb], f[s.nloc - s.hb:s.nloc]))
            for peer, send, recv in sides:
                data = send.clone()
                ev = torch.cuda.Event(); ev.record()
                self.mb.put((s.rank, peer, self.seq, fi), (data, ev))
                reqs.append(_Recv(self.mb, (peer, s.rank, self.seq, fi), recv))
        return reqs


def run_ranks(size, body):
    errs = []

    def tgt(r):
        try:
            body(r)
        except BaseException as e:      # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=tgt, args=(r,)) for r in range(size)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0][1]


@pytest.fixture(scope="module")
def D():
    from npbench_b200 import distributed
    return distributed


def _dev(full, slab):
    return torch.from_numpy(np.ascontiguousarray(full[slab.row0:slab.row0 + slab.nloc])).cuda()


@pytest.mark.parametrize("march", [False, True], ids=["blocked", "marching-dual"])
@pytest.mark.parametrize("size", [2, 3])
def test_jacobi_sharded_on_gpu(D, size, march):
    """march: marching passes forced on the small slabs; the driver then runs the scratch-slab plan whose closing
    pass stores the last two states (npb_jacobi2d_block2_f64)."""
    import npbench_b200 as nb
    rng = np.random.default_rng(1)
    ni, nj, ts = 333, 270, 12
    A, B = rng.random((ni, nj)), rng.random((ni, nj))
    wA, wB = A.copy(), B.copy()
    oracle.jacobi_2d(ts, wA, wB)
    mb, out = Mailbox(), {}

    def body(r):
        eng = D.B200Engine(0)
        slab = D.Slab(ni, size, r, D.JACOBI_MAX_BLOCK)
        assert eng.jacobi_dual_ok(slab.nloc, nj) == march
        lA, lB = _dev(A, slab), _dev(B, slab)
        D.jacobi_2d_sharded(eng, slab, ts, lA, lB, exchanger=LoopbackExchanger(slab, mb))
        eng.synchronize()
        out[r] = (slab, slab.owned(lA).cpu().numpy(), slab.owned(lB).cpu().numpy())

    nb.init(0)
    nb.lib().jacobi2d_set_mode(3 if march else 0)
    try:
        run_ranks(size, body)
    finally:
        nb.lib().jacobi2d_set_mode(0)
    for r, (slab, a, b) in out.items():
        assert_bit_equal(a, wA[slab.lo:slab.hi], "A rank %d" % r)
        assert_bit_equal(b, wB[slab.lo:slab.hi], "B rank %d" % r)


@pytest.mark.parametrize("march,H", [(False, 4), (True, 3), (True, 4)], ids=["per-sweep", "march-H3", "march-H4"])
@pytest.mark.parametrize("size", [2, 3])
def test_heat_sharded_on_gpu(D, size, march, H):
    """march: three sweeps per pass (heat3d_march_kernel) over the slab's plane ranges, one exchange per pass"""
    rng = np.random.default_rng(2)
    shape, ts = (61, 20, 33), 9
    A, B = rng.random(shape), rng.random(shape)
    wA, wB = A.copy(), B.copy()
    oracle.heat_3d(ts, wA, wB)
    mb, out = Mailbox(), {}

    def body(r):
        eng = D.B200Engine(0)
        slab = D.Slab(shape[0], size, r, H)
        lA, lB = _dev(A, slab), _dev(B, slab)
        D.heat_3d_sharded(eng, slab, ts, lA, lB, exchanger=LoopbackExchanger(slab, mb), march=march)
        eng.synchronize()
        out[r] = (slab, slab.owned(lA).cpu().numpy(), slab.owned(lB).cpu().numpy())

    run_ranks(size, body)
    for r, (slab, a, b) in out.items():
        assert_bit_equal(a, wA[slab.lo:slab.hi], "A rank %d" % r)
        assert_bit_equal(b, wB[slab.lo:slab.hi], "B rank %d" % r)


@pytest.mark.parametrize("march,H", [(False, 4), (True, 4), (True, 5), (True, 2)], ids=["per-step", "march-H4", "march-H5", "march-H2"])
@pytest.mark.parametrize("size,tm", [(2, 11), (3, 6)])
def test_fdtd_sharded_on_gpu(D, size, tm, march, H):
    """march: min(H, 5) steps per pass (fdtd2d_march_kernel) over the slab's row ranges, one exchange per pass"""
    rng = np.random.default_rng(3)
    nx, ny = 90, 301
    f = [rng.random((nx, ny)) for _ in range(3)]
    fict = rng.random(tm)
    w = [x.copy() for x in f]
    oracle.fdtd_2d(tm, w[0], w[1], w[2], fict)
    mb, out = Mailbox(), {}

    def body(r):
        eng = D.B200Engine(0)
        slab = D.Slab(nx, size, r, H)
        l = [_dev(x, slab) for x in f]
        D.fdtd_2d_sharded(eng, slab, tm, l[0], l[1], l[2], fict, exchanger=LoopbackExchanger(slab, mb), march=march)
        eng.synchronize()
        out[r] = (slab, [slab.owned(x).cpu().numpy() for x in l])

    run_ranks(size, body)
    for r, (slab, got) in out.items():
        for name, g, want in zip(("ex", "ey", "hz"), got, w):
            assert_bit_equal(g, want[slab.lo:slab.hi], "%s rank %d" % (name, r))


def test_march_entry_points_on_row_ranges(D):
    """npb_fdtd2d_march_f64 / npb_heat3d_march_f64 on a whole grid, written range by range (boundary ranges first, like
    the sharded driver): equal to ns steps / 3 sweeps of the oracle."""
    eng = D.B200Engine(0)
    rng = np.random.default_rng(17)
    nx, ny, ns = 150, 260, 5
    f = [rng.random((nx, ny)) for _ in range(3)]
    fict = rng.random(ns)
    src = [torch.from_numpy(x.copy()).cuda() for x in f]
    dst = [torch.full_like(x, float("nan")) for x in src]
    fd = eng.fict_on_device(fict)
    for lo, hi in ((0, 7), (140, 150), (7, 64), (64, 140)):
        eng.fdtd_march(ns, nx, 0, src, dst, fd, 0, lo, hi)
    eng.synchronize()
    oracle.fdtd_2d(ns, f[0], f[1], f[2], fict)
    for name, g, w in zip(("ex", "ey", "hz"), dst, f):
        assert_bit_equal(g.cpu().numpy(), w, name)
    shape = (40, 30, 70)
    A, B = rng.random(shape), rng.random(shape)
    dA, dB = torch.from_numpy(A.copy()).cuda(), torch.from_numpy(B.copy()).cuda()
    for lo, hi in ((1, 4), (33, 39), (4, 20), (20, 33)):
        eng.heat_march(dA, dB, lo, hi)
    eng.synchronize()
    A0 = A.copy()
    oracle.heat_3d_sweeps(3, A, B)          # the oracle ping-pongs: A ends as state 2, B as state 3
    assert_bit_equal(dB.cpu().numpy(), B, "state 3 in dst"); assert_bit_equal(dA.cpu().numpy(), A0, "src untouched")


def test_halo_free_shards_on_gpu(D):
    eng = D.B200Engine(0)
    I, J, K = 37, 24, 60
    inf, outf, coeff = oracle.init_hdiff(I, J, K)
    want = outf.copy(); oracle.hdiff(inf, want, coeff)
    parts = []
    for r in range(3):
        lo, hi = D.hdiff_shard(I, 3, r)
        o = eng.empty(hi - lo, J, K)
        eng.hdiff(torch.from_numpy(inf[lo:hi + 4].copy()).cuda(), o, torch.from_numpy(coeff[lo:hi].copy()).cuda())
        parts.append(o.cpu().numpy())
    assert_bit_equal(np.concatenate(parts), want, "hdiff shards")
    dtr, us, u, w, up, ut = oracle.init_vadv(I, J, K)
    want = us.copy(); oracle.vadv(want, u, w, up, ut, dtr)
    parts = []
    for r in range(3):
        lo, hi = D.vadv_shard(I, 3, r)
        t = [torch.from_numpy(x[lo:hi].copy()).cuda() for x in (us, u)] + [torch.from_numpy(w[lo:hi + 1].copy()).cuda()] \
            + [torch.from_numpy(x[lo:hi].copy()).cuda() for x in (up, ut)]
        eng.vadv(*t, dtr)
        parts.append(t[0].cpu().numpy())
    assert_bit_equal(np.concatenate(parts), want, "vadv shards")
