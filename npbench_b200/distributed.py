"""Slab-sharded multi-GPU drivers: one process per GPU, halo exchange over NCCL/NVLink.

The reference is single-device (no NCCL/MPI anywhere); this module is
the B200 scale-out of its stencil family, launched with torchrun (one rank per GPU):

  * jacobi_2d, heat_3d, fdtd_2d shard along the slowest axis (rows / i-planes are contiguous
    in memory, so a halo is one contiguous block).  Every rank keeps `H` ghost rows on each
    interior side and advances up to `H` sweeps between exchanges (the ghost zone absorbs the
    contamination that creeps in one row per sweep), so the number of messages is
    sweeps / H, not sweeps.  Exchanges are torch.distributed P2P batches (ncclSend/ncclRecv
    inside one group) issued on a side stream as soon as the boundary rows of the last sweep
    of a round are final; the interior of that sweep overlaps the transfer.
  * hdiff and vadv shard along I with a fixed overlap (4 input rows / 1 wcon row) and need
    no run-time exchange at all.

The numerical work is delegated to an *engine*: `B200Engine` launches the CUDA kernels of
libnpb_b200.so on torch CUDA tensors (torch is plumbing here: memory, streams, NCCL);
tests substitute a CPU engine so that the decomposition/halo logic is covered with the
gloo backend on machines without GPUs.  Results are bit-identical to the single-device
kernels for any number of ranks.
"""
import threading
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

JACOBI_MAX_BLOCK = 7   # == NPB_JACOBI2D_MAX_BLOCK
FDTD_GHOST = 5         # ghost rows of the fdtd_2d slabs = steps per marching pass (FM_MAX_STEPS) = steps between exchanges
HEAT_GHOST = 3         # ghost planes of the heat_3d slabs = sweeps per marching pass = sweeps between exchanges


# --------------------------------------------------------------------------- partition
def slab_bounds(n: int, size: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of an n-row grid owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Slab:
    """Geometry of one rank's row slab with `H` ghost rows on interior sides."""

    def __init__(self, n_global: int, size: int, rank: int, H: int):
        self.n_global, self.size, self.rank, self.H = n_global, size, rank, H
        self.lo, self.hi = slab_bounds(n_global, size, rank)
        if size > 1 and (self.hi - self.lo) < H:
            raise ValueError("slab of %d rows is thinner than the ghost depth %d" % (self.hi - self.lo, H))
        self.ht = H if rank > 0 else 0            # ghost rows above
        self.hb = H if rank < size - 1 else 0     # ghost rows below
        self.row0 = self.lo - self.ht             # global index of local row 0
        self.nloc = (self.hi - self.lo) + self.ht + self.hb

    def owned(self, t: torch.Tensor) -> torch.Tensor:
        return t[self.ht:self.nloc - self.hb]


def jacobi_plan(total_sweeps: int, max_block: int = JACOBI_MAX_BLOCK) -> List[int]:
    """Split S sweeps into odd-sized blocked passes, the last one a single sweep (see
    csrc/jacobi2d.cu: npb_jacobi2d_f64 -- the same plan, so results and traffic match)."""
    if total_sweeps <= 0:
        return []
    m = total_sweeps - 1
    n = (m + max_block - 1) // max_block
    if n % 2 == 0:
        n += 1
    extra, cap, plan = (m - n) // 2, (max_block - 1) // 2, []
    for p in range(n):
        left = n - p
        take = min(cap, (extra + left - 1) // left)
        extra -= take
        plan.append(1 + 2 * take)
    return plan + [1]


def jacobi_plan_dual(total_sweeps: int, max_block: int = JACOBI_MAX_BLOCK) -> List[int]:
    """Split S (even) sweeps into an EVEN number of odd-sized passes, smaller ones first; the last pass stores
    the states S and S - 1 (csrc/jacobi2d.cu: the scratch-grid plan of npb_jacobi2d_f64 -- the same sizes)."""
    assert total_sweeps >= 2 and total_sweeps % 2 == 0
    k = (total_sweeps + max_block - 1) // max_block
    if k % 2:
        k += 1
    k = max(k, 2)
    pairs, cap, plan = (total_sweeps - k) // 2, (max_block - 1) // 2, []
    for p in range(k):
        take = min(cap, pairs // (k - p))
        pairs -= take
        plan.append(1 + 2 * take)
    return plan


# --------------------------------------------------------------------------- communication
class HaloExchanger:
    """Exchange H boundary rows of row-major slabs with rank-1 / rank+1."""

    def __init__(self, slab: Slab, group=None):
        self.slab, self.group = slab, group

    def start(self, fields: Sequence[torch.Tensor]):
        """Post sends of the outermost owned rows and receives into the ghost rows.
        Returns the list of in-flight requests (empty on a single rank)."""
        s, H, ops = self.slab, self.slab.H, []
        for f in fields:
            if s.ht:
                ops.append(dist.P2POp(dist.isend, f[s.ht:s.ht + H], s.rank - 1, self.group))
                ops.append(dist.P2POp(dist.irecv, f[0:s.ht], s.rank - 1, self.group))
            if s.hb:
                ops.append(dist.P2POp(dist.isend, f[s.nloc - s.hb - H:s.nloc - s.hb], s.rank + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, f[s.nloc - s.hb:s.nloc], s.rank + 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs) -> None:
        for r in reqs:
            r.wait()


# --------------------------------------------------------------------------- engines
class B200Engine:
    """Launches libnpb_b200.so kernels on torch CUDA tensors; two streams for overlap."""

    _launch_lock = threading.Lock()   # the library's current stream is process global

    def __init__(self, device: int):
        from . import _lib
        self.torch_device = torch.device("cuda", device)
        torch.cuda.set_device(self.torch_device)
        self.lib = _lib.init(device)
        # a dedicated compute stream shared by torch and the library (handle 0 -- torch's legacy
        # default stream -- means "library stream" to npb_set_stream, so never pass that one)
        self.compute = torch.cuda.Stream()
        torch.cuda.set_stream(self.compute)
        self.comm = torch.cuda.Stream()
        self.lib.set_stream(self.compute.cuda_stream)
        self.tile_rows = self.lib.jacobi2d_tile_rows()

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, dtype=torch.float64, device=self.torch_device)

    # ---- kernels (all asynchronous on the compute stream)
    def _launch(self, fn, *args):
        # (re)bind the library to this engine's compute stream and enqueue, atomically
        with B200Engine._launch_lock:
            self.lib.set_stream(self.compute.cuda_stream)
            fn(*args)

    def jacobi_block(self, nsteps, src, dst, t_lo, t_hi, dst2=None):
        if dst2 is None:
            self._launch(self.lib.jacobi2d_block_f64, nsteps, src.shape[0], src.shape[1], src.data_ptr(),
                         dst.data_ptr(), t_lo, t_hi)
        else:       # the pass also stores the state before its last sweep (marching regime only)
            self._launch(self.lib.jacobi2d_block2_f64, nsteps, src.shape[0], src.shape[1], src.data_ptr(),
                         dst.data_ptr(), dst2.data_ptr(), t_lo, t_hi)

    def jacobi_dual_ok(self, nrows, ncols) -> bool:
        """True if a pass over an (nrows, ncols) slab can store two states (jacobi_block(..., dst2=...))."""
        return bool(self.lib.jacobi2d_block_marches(nrows, ncols))

    def heat_sweep(self, src, dst, i_lo, i_hi):
        n0, n1, n2 = src.shape
        self._launch(self.lib.heat3d_sweep_f64, n0, n1, n2, src.data_ptr(), dst.data_ptr(), i_lo, i_hi)

    def heat_march(self, src, dst, i_lo, i_hi):
        """three sweeps src -> dst over the output planes [i_lo, i_hi) (heat3d_march_kernel)"""
        n0, n1, n2 = src.shape
        self._launch(self.lib.heat3d_march_f64, n0, n1, n2, src.data_ptr(), dst.data_ptr(), i_lo, i_hi)

    def fict_on_device(self, fict):
        return torch.tensor([float(x) for x in fict], dtype=torch.float64, device=self.torch_device)

    def fdtd_march(self, ns, nx_global, row0, src, dst, fict_dev, t, r_lo, r_hi):
        """ns (2..5) steps src -> dst over the output rows [r_lo, r_hi) (fdtd2d_march_kernel); fict_dev[t] = _fict_[t]"""
        nrows, ny = src[0].shape
        self._launch(self.lib.fdtd2d_march_f64, int(ns), nx_global, row0, nrows, ny, src[0].data_ptr(), src[1].data_ptr(),
                     src[2].data_ptr(), dst[0].data_ptr(), dst[1].data_ptr(), dst[2].data_ptr(),
                     fict_dev.data_ptr() + 8 * int(t), r_lo, r_hi)

    def fdtd_step(self, nx_global, row0, src, dst, fict_t, r_lo, r_hi):
        nrows, ny = src[0].shape
        self._launch(self.lib.fdtd2d_step_f64, nx_global, row0, nrows, ny, src[0].data_ptr(), src[1].data_ptr(),
                     src[2].data_ptr(), dst[0].data_ptr(), dst[1].data_ptr(), dst[2].data_ptr(),
                     float(fict_t), r_lo, r_hi)

    def hdiff(self, inf, out, coeff):
        I, J, K = out.shape
        self._launch(self.lib.hdiff_f64, I, J, K, inf.data_ptr(), out.data_ptr(), coeff.data_ptr())

    def vadv(self, us, u, w, up, ut, dtr):
        I, J, K = us.shape
        self._launch(self.lib.vadv_f64, I, J, K, us.data_ptr(), u.data_ptr(), w.data_ptr(), up.data_ptr(),
                     ut.data_ptr(), float(dtr))

    def copy(self, dst, src):
        dst.copy_(src)

    # ---- stream choreography
    def boundary_done(self):
        ev = torch.cuda.Event()
        ev.record(self.compute)
        return ev

    def start_exchange(self, exchanger: HaloExchanger, fields, after):
        """Issue the P2P batch on the comm stream once event `after` has fired."""
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(after)
            return exchanger.start(fields)

    def finish_exchange(self, reqs):
        # req.wait() makes the *current* (compute) stream wait for the NCCL work
        HaloExchanger.finish(reqs)

    def synchronize(self):
        torch.cuda.synchronize(self.torch_device)


# --------------------------------------------------------------------------- drivers
def _boundary_tile_ranges(slab: Slab, tile_rows: int) -> Tuple[int, int, int]:
    """Tile-row split for the blocked jacobi pass on a local slab: tile rows [0, tb) and
    [te, ntr) hold every ghost row and every row that is sent; [tb, te) is the interior."""
    ntr = (slab.nloc - 2 + tile_rows - 1) // tile_rows
    tb = 0
    if slab.ht:
        tb = min(ntr, (slab.ht + slab.H - 2) // tile_rows + 1)          # covers rows 1 .. ht+H-1
    te = ntr
    if slab.hb:
        te = max(tb, (slab.nloc - slab.hb - slab.H - 1) // tile_rows)   # first tile with row nloc-hb-H
    return tb, te, ntr


def jacobi_2d_sharded(engine, slab: Slab, TSTEPS: int, A: torch.Tensor, B: torch.Tensor, group=None,
                      exchanger=None) -> None:
    """kernel(TSTEPS, A, B) of jacobi_2d_numpy.py:4-10 on a row slab.  A, B are the local
    (slab.nloc, ncols) arrays INCLUDING ghost rows, initialised consistently with the global
    grid (ghost rows hold the neighbour's rows).  On return the owned rows of A and B equal
    the corresponding rows of the single-device result, bit for bit."""
    assert slab.H >= JACOBI_MAX_BLOCK or slab.size == 1
    ex = exchanger or HaloExchanger(slab, group)
    tb, te, ntr = _boundary_tile_ranges(slab, engine.tile_rows)
    total = 2 * (TSTEPS - 1)

    def one_pass(n, src, dst, dst2, exchange):
        if slab.size == 1 or not exchange:
            engine.jacobi_block(n, src, dst, 0, ntr, dst2)
        else:
            if tb > 0:
                engine.jacobi_block(n, src, dst, 0, tb, dst2)
            if te < ntr:
                engine.jacobi_block(n, src, dst, te, ntr, dst2)
            reqs = engine.start_exchange(ex, [dst], engine.boundary_done())
            if te > tb:
                engine.jacobi_block(n, src, dst, tb, te, dst2)     # overlaps the halo transfer
            engine.finish_exchange(reqs)

    # Every rank must take the same decision: the thinnest slab of the partition decides whether the passes can store
    # two states (marching regime).  Then an even number of passes A -> B -> ... -> A -> W -> (A, B) covers all sweeps:
    # W, a scratch slab that stands in for B (it carries B's border ring), feeds the closing pass, which stores state
    # S into A and state S - 1 into B -- the separate single sweep B -> A and its exchange disappear.
    thinnest = min(Slab(slab.n_global, slab.size, r, slab.H).nloc for r in range(slab.size))
    if total >= 4 and getattr(engine, "jacobi_dual_ok", None) and engine.jacobi_dual_ok(thinnest, A.shape[1]):
        plan = jacobi_plan_dual(total)
        W = engine.empty(*A.shape)
        for sl in ((slice(None), slice(0, 1)), (slice(None), slice(-1, None)), (slice(0, 1), slice(None)),
                   (slice(-1, None), slice(None))):
            engine.copy(W[sl], B[sl])
        k = len(plan)
        for p, n in enumerate(plan):
            last = p == k - 1
            src = W if last else (B if p % 2 else A)
            dst = A if last else (W if p == k - 2 else (A if p % 2 else B))
            one_pass(n, src, dst, B if last else None, exchange=not last)
        return
    src, dst = A, B
    for n in jacobi_plan(total):
        one_pass(n, src, dst, None, exchange=True)
        src, dst = dst, src


def heat_plan(total_sweeps: int) -> List[int]:
    """Passes of three sweeps (heat3d_march_kernel) and single sweeps: an even number of passes, the last one a
    single sweep, so that state S ends in A and state S - 1 in B (run_march in csrc/heat3d_march.cuh: same plan)."""
    if total_sweeps <= 0:
        return []
    m = total_sweeps - 1
    n = (m + 2) // 3
    if n % 2 == 0:
        n += 1
    triples = max(0, (m - n) // 2)
    return [3] * triples + [1] * (n - triples) + [1]


def _edge_ranges(slab: Slab, n: int, lo_edge: int, hi_edge: int) -> Tuple[int, int]:
    """[b_lo, b_hi): the interior of the local rows [lo_edge, hi_edge); rows below b_lo feed the upward send, rows
    from b_hi on the downward send (each H rows, next to the ghost rows)."""
    H = slab.H
    b_lo = min(hi_edge, slab.ht + H) if slab.ht else lo_edge
    b_hi = max(b_lo, n - slab.hb - H) if slab.hb else hi_edge
    return b_lo, b_hi


def heat_3d_sharded(engine, slab: Slab, TSTEPS: int, A: torch.Tensor, B: torch.Tensor, group=None,
                    exchanger=None, march=None) -> None:
    """kernel(TSTEPS, A, B) of heat_3d_numpy.py:4-20 on an i-plane slab (ghost depth slab.H).  With a ghost depth of
    at least 3 and an engine that has it, the sweeps run as marching passes (three sweeps per pass over memory,
    heat3d_march_kernel) with one halo exchange per pass; otherwise one sweep per launch, one exchange per H sweeps."""
    ex = exchanger or HaloExchanger(slab, group)
    total = 2 * (TSTEPS - 1)
    n, H = slab.nloc, slab.H
    if march is None:
        # every rank must take the same decision (the passes carry the exchanges): judge by the thinnest slab
        thinnest = slab.n_global // slab.size + (H if slab.size > 1 else 0)
        march = hasattr(engine, "heat_march") and (H >= 3 or slab.size == 1) and thinnest >= 8
    src, dst = A, B
    if march:
        for k in heat_plan(total):
            run = (lambda lo, hi: engine.heat_march(src, dst, lo, hi)) if k == 3 else \
                  (lambda lo, hi: engine.heat_sweep(src, dst, lo, hi))
            if slab.size == 1:
                run(1, n - 1)
            else:
                # ghost planes are not computed: the exchange fills them
                first, last = max(1, slab.ht), min(n - 1, n - slab.hb)
                b_lo, b_hi = _edge_ranges(slab, n, first, last)
                if b_lo > first:
                    run(first, b_lo)
                if b_hi < last:
                    run(b_hi, last)
                reqs = engine.start_exchange(ex, [dst], engine.boundary_done())
                if b_hi > b_lo:
                    run(b_lo, b_hi)                 # overlaps the halo transfer
                engine.finish_exchange(reqs)
            src, dst = dst, src
        return
    done = 0
    while done < total:
        k = min(H, total - done) if slab.size > 1 else total - done
        for s in range(k):
            last = (s == k - 1) and slab.size > 1
            if not last:
                engine.heat_sweep(src, dst, 1, n - 1)
            else:
                b_lo = (slab.ht + H) if slab.ht else 1            # planes [1, b_lo) feed the upward send
                b_hi = (n - slab.hb - H) if slab.hb else n - 1    # planes [b_hi, n-1) feed the downward send
                b_lo = min(b_lo, n - 1); b_hi = max(b_hi, b_lo)
                if b_lo > 1:
                    engine.heat_sweep(src, dst, 1, b_lo)
                if b_hi < n - 1:
                    engine.heat_sweep(src, dst, b_hi, n - 1)
                reqs = engine.start_exchange(ex, [dst], engine.boundary_done())
                if b_hi > b_lo:
                    engine.heat_sweep(src, dst, b_lo, b_hi)
                engine.finish_exchange(reqs)
            src, dst = dst, src
        done += k
    # The array that was NOT written last holds state total-1 with ghost rows one round
    # stale -- irrelevant: only owned rows are results.


def fdtd_2d_sharded(engine, slab: Slab, TMAX: int, ex_: torch.Tensor, ey: torch.Tensor, hz: torch.Tensor,
                    fict: Sequence[float], group=None, exchanger=None, march=None) -> None:
    """kernel(TMAX, ex, ey, hz, _fict_) of fdtd_2d_numpy.py:4-11 on a row slab.  `fict` is a host sequence (the
    reference indexes _fict_[t] on the host too).  With an engine that has it, the steps run as marching passes of up
    to min(H, 5) steps per pass over memory (fdtd2d_march_kernel) with one halo exchange per pass; otherwise one
    launch per step, one exchange per H steps."""
    exch = exchanger or HaloExchanger(slab, group)
    user = [ex_, ey, hz]
    work = [engine.empty(*f.shape) for f in user]
    src, dst, t, H, n = user, work, 0, slab.H, slab.nloc
    if march is None:
        march = hasattr(engine, "fdtd_march") and slab.n_global // slab.size >= 2 and (H >= 2 or slab.size == 1)
    if march:
        fict_dev = engine.fict_on_device(fict)
        cap = FDTD_GHOST if slab.size == 1 else min(H, FDTD_GHOST)
        while t < TMAX:
            k = min(cap, TMAX - t)
            if k >= 2:
                run = lambda lo, hi: engine.fdtd_march(k, slab.n_global, slab.row0, src, dst, fict_dev, t, lo, hi)
            else:
                run = lambda lo, hi: engine.fdtd_step(slab.n_global, slab.row0, src, dst, float(fict[t]), lo, hi)
            if slab.size == 1:
                run(0, n)
            else:
                first, last = slab.ht, n - slab.hb         # ghost rows are not computed: the exchange fills them
                b_lo, b_hi = _edge_ranges(slab, n, first, last)
                if b_lo > first:
                    run(first, b_lo)
                if b_hi < last:
                    run(b_hi, last)
                reqs = engine.start_exchange(exch, dst, engine.boundary_done())
                if b_hi > b_lo:
                    run(b_lo, b_hi)                        # overlaps the halo transfer
                engine.finish_exchange(reqs)
            src, dst = dst, src
            t += k
    else:
        while t < TMAX:
            k = min(H, TMAX - t) if slab.size > 1 else TMAX - t
            for s in range(k):
                last = (s == k - 1) and slab.size > 1
                f_t = float(fict[t + s])
                if not last:
                    engine.fdtd_step(slab.n_global, slab.row0, src, dst, f_t, 0, n)
                else:
                    b_lo = min(n, slab.ht + H) if slab.ht else 0
                    b_hi = max(b_lo, n - slab.hb - H) if slab.hb else n
                    if b_lo > 0:
                        engine.fdtd_step(slab.n_global, slab.row0, src, dst, f_t, 0, b_lo)
                    if b_hi < n:
                        engine.fdtd_step(slab.n_global, slab.row0, src, dst, f_t, b_hi, n)
                    reqs = engine.start_exchange(exch, dst, engine.boundary_done())
                    if b_hi > b_lo:
                        engine.fdtd_step(slab.n_global, slab.row0, src, dst, f_t, b_lo, b_hi)
                    engine.finish_exchange(reqs)
                src, dst = dst, src
            t += k
    if src is not user:          # an odd number of passes leaves the result in the workspace
        for u, w in zip(user, src):
            engine.copy(u, w)


# --------------------------------------------------------------------------- halo-free shards
def hdiff_shard(I: int, size: int, rank: int) -> Tuple[int, int]:
    """Output rows [lo, hi) of rank; it needs in_field[lo : hi + 4] (fixed 4-row overlap,
    hdiff_numpy.py:7-28) and coeff/out_field[lo:hi]."""
    return slab_bounds(I, size, rank)


def vadv_shard(I: int, size: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank; wcon needs rows [lo : hi + 1] (vadv_numpy.py:16, 33-34)."""
    return slab_bounds(I, size, rank)
