// jacobi2d_regtile.cuh -- on-chip resident jacobi_2d for grids that fit in the SMs' registers (NPBench presets
// S / M / L): ONE cooperative launch runs all 2 * (TSTEPS - 1) sweeps of kernel(TSTEPS, A, B),
// npbench/benchmarks/polybench/jacobi_2d/jacobi_2d_numpy.py:4-10.
//
// Why this shape.  The blocked passes (jacobi2d_block_kernel) pay a kernel launch, a tile load and a tile store per
// 7 sweeps: 700^2 took 58 graph nodes of 15 us.  The grid is 3.9 MB; the register files of 148 SMs hold 37 MB.  So:
//
//   * The interior is cut into PI x PJ tiles, one CTA (= one SM) each.  A CTA keeps its tile plus a T-deep halo ring
//     -- its REGION, (NW * RB) rows x (32 * CB) columns -- IN REGISTERS for the whole time loop: warp w owns RB
//     consecutive region rows, lane l owns CB consecutive columns of them (lanes along the contiguous axis).
//     Of a cell's four neighbours the lateral ones come from the thread's own registers or from the adjacent lane by
//     warp shuffle, the vertical ones from its own registers or, across a warp edge, from shared memory, where every
//     thread publishes the first and last row of its block once per sweep (double buffered by sweep parity: one
//     __syncthreads per sweep).  No tile loads, no index arithmetic, 5 FP64 operations per cell.
//   * Halos are exchanged only every T sweeps: between exchanges the region is updated as a whole, so cells within q
//     of a region edge that faces another tile hold garbage after q sweeps of a block -- the tile itself sits T cells
//     inside and never sees it (the classic redundant-halo scheme, here without any shared-memory tile).  After the
//     T-th sweep every thread sends the cells of its tile that lie within T of a tile edge straight from its
//     registers into the inboxes of the neighbour tiles that hold them as halo (a thread's block has at most three:
//     vertical, horizontal, diagonal), and the threads that hold halo cells poll them straight into their registers
//     (sentinel protocol of inbox.cuh: relaxed gpu-scope stores / loads on the value itself, re-arm after
//     consumption, ring of SLOTS exchange slots, a gpu-scope fence before the sends of every FENCE_EVERY-th
//     exchange).  One exchange costs ~2 us (tools/pingpong.cu: 0.43 us for two CTAs, 1.7 us when 144 exchange at
//     once) -- several sweeps of arithmetic -- so T trades redundant rim work against exposed round trips; the
//     host picks it per grid from a cost model (pick_regtile_config in jacobi2d.cu).
//   * The constant border ring: even states carry A's border, odd states B's (jacobi_2d_numpy.py never writes them).
//     Region cells on the ring or outside the grid are "fixed": after every sweep their registers are reloaded from a
//     per-parity shared copy, so neighbours see exactly the border value.  Only threads that own such cells pay.
//   * Inboxes stay armed between calls (every cell that is sent is consumed and re-armed, the last block sends
//     nothing); the launch is cooperative only to guarantee that all CTAs are co-resident.
//
// Arithmetic: 0.2 * ((((c + left) + right) + down) + up), NumPy's order, -fmad=false (as jacobi2d_block_kernel).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "inbox.cuh"

namespace j2rt {

constexpr int SLOTS = 8;          // inbox ring depth, in exchanges
constexpr int FENCE_EVERY = 4;    // gpu-scope fence cadence, in exchanges (2 * cadence <= SLOTS, see inbox.cuh)

struct Params {
    int ni, nj;
    int PI, PJ;                   // tiles along i and j
    int T;                        // sweeps per halo exchange = halo depth
    int nsweeps;                  // 2 * (TSTEPS - 1)
    int NW;                       // warps per CTA; region = (NW * RB) x (32 * CB)
    double *A, *B;
    unsigned long long *inbox;    // [PI * PJ][SLOTS][region cells]
};

// CB consecutive doubles at a 16-byte aligned shared address (CB is 2 or 4)
template <int CB>
__device__ __forceinline__ void lds_row(const double *a, double (&v)[CB]) {
#pragma unroll
    for (int b = 0; b < CB; b += 2) {
        const double2 t = *reinterpret_cast<const double2 *>(a + b);
        v[b] = t.x; v[b + 1] = t.y;
    }
}
template <int CB>
__device__ __forceinline__ void sts_row(double *a, const double (&v)[CB]) {
#pragma unroll
    for (int b = 0; b < CB; b += 2) *reinterpret_cast<double2 *>(a + b) = make_double2(v[b], v[b + 1]);
}

__global__ void jacobi2d_inbox_arm_kernel(unsigned long long *box, size_t n) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x)
        box[w] = HR_SENTINEL;
}

// One sweep of the thread's RB x CB cells: o = state s - 1 (registers), v = state s.  All shuffles and both shared
// loads first, then the five operations of NumPy's order in STAGES over the independent cells: a warp issues in
// order, so the dependent chain of one cell must be interleaved with the other cells'.
template <int RB, int CB>
__device__ __forceinline__ void sweep_cells(const double (&o)[RB][CB], double (&v)[RB][CB], const double *pub_me, int pub_tb, int rp) {
    constexpr int RC = 32 * CB;
    double up[CB], dn[CB], lft[RB], rgt[RB];
#pragma unroll
    for (int a = 0; a < RB; ++a) {
        lft[a] = __shfl_up_sync(0xffffffffu, o[a][CB - 1], 1);
        rgt[a] = __shfl_down_sync(0xffffffffu, o[a][0], 1);
    }
    lds_row<CB>(pub_me - RC + pub_tb + rp, up);                       // bottom row of warp w - 1
    lds_row<CB>(pub_me + RC + rp, dn);                                // top row of warp w + 1
#pragma unroll
    for (int b = CB - 1; b >= 0; --b)                                 // c + left (columns with an own left neighbour first)
#pragma unroll
        for (int a = 0; a < RB; ++a) v[a][b] = o[a][b] + (b ? o[a][b - 1] : lft[a]);
#pragma unroll
    for (int b = 0; b < CB; ++b)                                      // + right
#pragma unroll
        for (int a = 0; a < RB; ++a) v[a][b] = v[a][b] + ((b < CB - 1) ? o[a][b + 1] : rgt[a]);
#pragma unroll
    for (int a = 0; a < RB; ++a)                                      // + down (row i + 1)
#pragma unroll
        for (int b = 0; b < CB; ++b) v[a][b] = v[a][b] + ((a < RB - 1) ? o[a + 1][b] : dn[b]);
#pragma unroll
    for (int a = RB - 1; a >= 0; --a)                                 // + up (row i - 1)
#pragma unroll
        for (int b = 0; b < CB; ++b) v[a][b] = v[a][b] + (a ? o[a - 1][b] : up[b]);
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) v[a][b] = 0.2 * v[a][b];
}
// cells on the constant border ring (or outside the grid) take the value of the state's parity
template <int RB, int CB>
__device__ __forceinline__ void fix_cells(double (&v)[RB][CB], unsigned fixed, const double *fx) {
    constexpr int RC = 32 * CB;
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b)
            if ((fixed >> (a * CB + b)) & 1u) v[a][b] = fx[a * RC + b];
}

// Per-thread roles of the cold paths (exchange, fixed cells, final stores) live in shared memory, not in registers:
// the sweep loop is latency bound on the FP64 pipe, and ptxas only interleaves the RB * CB independent chains of a
// thread when it has registers to spare (with the role masks, tile bounds and inbox pointers live across the loop it
// serialised the cells two at a time: 0.46 us per sweep for ONE warp).
struct Desc {
    unsigned rows;      // bits 0-7 own rows, 8-15 own rows within T of the tile's top edge (neighbour above), 16-23 bottom
    unsigned cols;      // bits 0-7 own columns, 8-15 near the left edge, 16-23 near the right edge
    unsigned halo;      // cells (bit a * CB + b) received from a neighbour tile at every exchange
    unsigned fixed;     // cells on the constant border ring or outside the grid
    // sends: up to three neighbour tiles (vertical, horizontal, diagonal -- a tile is at least 2 T + RB - 1 cells
    // high and 2 T + CB - 1 wide where it has neighbours, so a thread's block never touches two opposite edges)
    int off[3];         // inbox word offset of the block's cell (0, 0) in that neighbour's slot 0 (may be negative)
    unsigned msk[3];    // cells that go there
};
struct Geo { int ilo, ihi, jlo, jhi; };

template <int RB, int CB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) jacobi2d_regtile_kernel(Params p) {
    extern __shared__ __align__(16) double sm[];
    __shared__ Geo s_geo;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NW = p.NW, RR = NW * RB, T = p.T;
    constexpr int RC = 32 * CB;
    const int cells = RR * RC;
    double *const fx0 = sm, *const fx1 = sm + cells;                  // fixed cells of even / odd states
    double *const pub = sm + 2 * cells;                               // [parity][top | bottom][NW + 2][RC]
    const int pub_tb = (NW + 2) * RC, pub_par = 2 * pub_tb;
    Desc *const s_desc = reinterpret_cast<Desc *>(pub + 2 * pub_par); // [threads]
    unsigned roles = 0;                                               // 1 fixed cells, 2 halo cells, 4 sends, 8 owns tile cells
    double o[RB][CB];                                                 // own cells, state s - 1
    {
        const int ni = p.ni, nj = p.nj;
        const int ti = blockIdx.x / p.PJ, tj = blockIdx.x % p.PJ;
        int ilo, ihi, jlo, jhi;
        tile_bounds(ni - 2, p.PI, ti, ilo, ihi);
        tile_bounds(nj - 2, p.PJ, tj, jlo, jhi);
        if (tid == 0) {
            s_geo = Geo{ilo, ihi, jlo, jhi};
        }
        // region (r, c) <-> global (ilo - T + r, jlo - T + c); this thread: rows w * RB + a, columns lane * CB + b
        const int gi0 = ilo - T + w * RB, gj0 = jlo - T + lane * CB;
        // ---- shared copies of the fixed cells (border ring of A / of B; 0 outside the grid), zeroed row exchange area
        for (int x = tid; x < cells; x += blockDim.x) {
            const int r = x / RC, c = x - r * RC;
            const int gi = ilo - T + r, gj = jlo - T + c;
            const bool inside = gi >= 0 && gi < ni && gj >= 0 && gj < nj;
            const bool interior = gi >= 1 && gi <= ni - 2 && gj >= 1 && gj <= nj - 2;
            const bool ring = inside && !interior;
            fx0[x] = ring ? __ldg(p.A + (long long)gi * nj + gj) : 0.0;
            fx1[x] = ring ? __ldg(p.B + (long long)gi * nj + gj) : 0.0;
        }
        for (int x = tid; x < 2 * pub_par; x += blockDim.x) pub[x] = 0.0;
        // ---- roles of this thread's cells
        Desc d{0, 0, 0, 0, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int gi = gi0 + a;
            if (gi >= ilo && gi < ihi) {
                d.rows |= 1u << a;
                if (gi - ilo < T && ti > 0) d.rows |= 256u << a;
                if (gi >= ihi - T && ti < p.PI - 1) d.rows |= 65536u << a;
            }
        }
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const int gj = gj0 + b;
            if (gj >= jlo && gj < jhi) {
                d.cols |= 1u << b;
                if (gj - jlo < T && tj > 0) d.cols |= 256u << b;
                if (gj >= jhi - T && tj < p.PJ - 1) d.cols |= 65536u << b;
            }
        }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                const int gi = gi0 + a, gj = gj0 + b;
                const bool inside = gi >= 0 && gi < ni && gj >= 0 && gj < nj;
                const bool interior = gi >= 1 && gi <= ni - 2 && gj >= 1 && gj <= nj - 2;
                const bool own = ((d.rows >> a) & 1u) && ((d.cols >> b) & 1u);
                const bool need = gi >= ilo - T && gi < ihi + T && gj >= jlo - T && gj < jhi + T;
                if (!interior) d.fixed |= 1u << (a * CB + b);
                if (interior && !own && need) d.halo |= 1u << (a * CB + b);
                o[a][b] = inside ? __ldg(p.A + (long long)gi * nj + gj) : 0.0;
            }
        {
            const unsigned r_top = (d.rows >> 8) & 255u, r_bot = (d.rows >> 16) & 255u, r_own = d.rows & 255u;
            const unsigned c_lft = (d.cols >> 8) & 255u, c_rgt = (d.cols >> 16) & 255u, c_own = d.cols & 255u;
            const int di = r_top ? -1 : (r_bot ? 1 : 0), dj = c_lft ? -1 : (c_rgt ? 1 : 0);
            const unsigned r_edge = r_top | r_bot, c_edge = c_lft | c_rgt;
            int ilo_n = ilo, jlo_n = jlo, hi;
            if (di) tile_bounds(ni - 2, p.PI, ti + di, ilo_n, hi);
            if (dj) tile_bounds(nj - 2, p.PJ, tj + dj, jlo_n, hi);
            const long long box_words = (long long)SLOTS * cells;
            auto cellmask = [](unsigned rs, unsigned cs) {
                unsigned m = 0;
#pragma unroll
                for (int a = 0; a < RB; ++a)
#pragma unroll
                    for (int b = 0; b < CB; ++b)
                        if (((rs >> a) & 1u) && ((cs >> b) & 1u)) m |= 1u << (a * CB + b);
                return m;
            };
            auto offset = [&](int ddi, int ddj) {
                const long long nb = (long long)(ti + ddi) * p.PJ + (tj + ddj);
                return (int)(nb * box_words + (long long)(gi0 - (ddi ? ilo_n : ilo) + T) * RC + (gj0 - (ddj ? jlo_n : jlo) + T));
            };
            d.msk[0] = cellmask(r_edge, c_own); d.off[0] = d.msk[0] ? offset(di, 0) : 0;       // vertical neighbour
            d.msk[1] = cellmask(r_own, c_edge); d.off[1] = d.msk[1] ? offset(0, dj) : 0;       // horizontal
            d.msk[2] = cellmask(r_edge, c_edge); d.off[2] = d.msk[2] ? offset(di, dj) : 0;     // diagonal
        }
        s_desc[tid] = d;
        const bool sender = (d.msk[0] | d.msk[1] | d.msk[2]) != 0u;
        roles = (d.fixed ? 1u : 0u) | (d.halo ? 2u : 0u) | (sender ? 4u : 0u) | (((d.rows & 255u) && (d.cols & 255u)) ? 8u : 0u);
    }
    __syncthreads();

    // shared addresses of this thread's row-exchange cells (parity 0): own top row; + pub_tb: own bottom row;
    // - RC + pub_tb: bottom row of warp w - 1; + RC: top row of warp w + 1
    double *const pub_me = pub + (w + 1) * RC + lane * CB;
    sts_row<CB>(pub_me, o[0]);
    sts_row<CB>(pub_me + pub_tb, o[RB - 1]);
    // A tile's first exchange cannot complete before its neighbours have read their initial halos, so nobody
    // overwrites A (last sweep) under a neighbour that is still loading -- unless there is no exchange at all
    if (p.nsweeps <= T && gridDim.x > 1) cooperative_groups::this_grid().sync();
    __syncthreads();

    const int nsweeps = p.nsweeps;
    const double *const fx_me = fx0 + (w * RB) * RC + lane * CB;      // + cells: odd states
    int s = 0;                                                        // sweeps done
    int nx = 0;                                                       // exchanges done
    while (s < nsweeps) {
        // ---- one block: up to T sweeps, then an exchange -- or, in the last block, the two final states leave the chip.
        //      All but the block's last sweep (last two in the last block) run in a loop without any of those checks.
        const int nb = min(T, nsweeps - s);
        const bool last_block = (s + nb == nsweeps);
        const int lean = last_block ? max(nb - 2, 0) : nb - 1;
        for (int q = 0; q < lean; ++q) {
            ++s;
            double v[RB][CB];
            sweep_cells<RB, CB>(o, v, pub_me, pub_tb, ((s - 1) & 1) * pub_par);
            if (roles & 1u) fix_cells<RB, CB>(v, s_desc[tid].fixed, fx_me + (s & 1) * cells);
            sts_row<CB>(pub_me + (s & 1) * pub_par, v[0]);
            sts_row<CB>(pub_me + (s & 1) * pub_par + pub_tb, v[RB - 1]);
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int b = 0; b < CB; ++b) o[a][b] = v[a][b];
            __syncthreads();
        }
        for (int q = lean; q < nb; ++q) {
            ++s;
            // A gpu-scope fence before the sends of every FENCE_EVERY-th exchange orders this CTA's earlier re-arms
            // before those sends, hence (the neighbour fences likewise after receiving them) before the neighbour's
            // next writes to the same cells SLOTS exchanges later: 2 * FENCE_EVERY <= SLOTS.  It sits a block of
            // sweeps after the re-arm stores were issued -- they have long completed -- not right behind them.
            if (!last_block && ((nx + 1) % FENCE_EVERY) == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            double v[RB][CB];
            sweep_cells<RB, CB>(o, v, pub_me, pub_tb, ((s - 1) & 1) * pub_par);
            if (!last_block) {
                // ---- halo exchange: tile cells within T of a tile edge go to the neighbours that hold them as halo
                ++nx;
                const unsigned slot_sz = (unsigned)cells, box_sz = (unsigned)SLOTS * slot_sz;
                const unsigned out_off = (unsigned)(nx % SLOTS) * slot_sz;
                if (roles & 6u) {
                    const Desc d = s_desc[tid];
                    unsigned long long *const box = p.inbox;
                    if (roles & 4u) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const unsigned m = d.msk[k];
                            if (!m) continue;
                            unsigned long long *const dst = box + ((long long)d.off[k] + out_off);
#pragma unroll
                            for (int a = 0; a < RB; ++a)
#pragma unroll
                                for (int b = 0; b < CB; ++b)
                                    if ((m >> (a * CB + b)) & 1u) st_relaxed_f64((double *)(dst + a * RC + b), v[a][b]);
                        }
                    }
                    if (roles & 2u) {
                        unsigned long long *const qb = box + (size_t)blockIdx.x * box_sz + out_off + (unsigned)(w * RB) * RC + lane * CB;
                        unsigned pend = d.halo;
                        // spin on ONE cell (the thread's last) until the neighbour's stores start to land: every
                        // polling thread of every CTA spinning on all its cells floods the L2 that has to deliver them
                        {
                            const int hb = 31 - __clz(pend);
                            unsigned long long *const q1 = qb + (hb / CB) * RC + (hb % CB);
                            while (ld_relaxed_u64(q1) == HR_SENTINEL) {}
                        }
                        do {
#pragma unroll
                            for (int a = 0; a < RB; ++a)
#pragma unroll
                                for (int b = 0; b < CB; ++b)
                                    if ((pend >> (a * CB + b)) & 1u)
                                        v[a][b] = __longlong_as_double((long long)ld_relaxed_u64(qb + a * RC + b));
#pragma unroll
                            for (int a = 0; a < RB; ++a)
#pragma unroll
                                for (int b = 0; b < CB; ++b)
                                    if (((pend >> (a * CB + b)) & 1u) && (unsigned long long)__double_as_longlong(v[a][b]) != HR_SENTINEL) {
                                        st_relaxed_u64(qb + a * RC + b, HR_SENTINEL);  // re-arm for exchange + SLOTS
                                        pend &= ~(1u << (a * CB + b));
                                    }
                        } while (pend);
                    }
                }
            }
            if (roles & 1u) fix_cells<RB, CB>(v, s_desc[tid].fixed, fx_me + (s & 1) * cells);
            if (s >= nsweeps - 1 && (roles & 8u)) {
                // ---- the last two states leave the chip: state S - 1 (odd) is B's, state S is A's (jacobi_2d_numpy.py:8-10)
                const Desc d = s_desc[tid];
                const int gi0 = s_geo.ilo - T + w * RB, gj0 = s_geo.jlo - T + lane * CB;
                double *const gp = ((s & 1) ? p.B : p.A) + (long long)gi0 * p.nj + gj0;
#pragma unroll
                for (int a = 0; a < RB; ++a)
#pragma unroll
                    for (int b = 0; b < CB; ++b)
                        if (((d.rows >> a) & 1u) && ((d.cols >> b) & 1u)) gp[(long long)a * p.nj + b] = v[a][b];
            }
            sts_row<CB>(pub_me + (s & 1) * pub_par, v[0]);
            sts_row<CB>(pub_me + (s & 1) * pub_par + pub_tb, v[RB - 1]);
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int b = 0; b < CB; ++b) o[a][b] = v[a][b];
            __syncthreads();
        }
    }
}

}  // namespace j2rt
