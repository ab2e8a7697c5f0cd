// fdtd2d_regtile.cuh -- on-chip resident fdtd_2d for grids that fit in the SMs' registers (NPBench presets S / M / L):
// ONE cooperative launch runs all TMAX steps of kernel(TMAX, ex, ey, hz, _fict_),
// npbench/benchmarks/polybench/fdtd_2d/fdtd_2d_numpy.py:4-11.
//
// Same scheme as jacobi2d_regtile.cuh (read its header first): the grid is cut into PI x PJ tiles, one CTA each; a
// CTA keeps its tile plus a T-deep halo ring -- (NW * RB) rows x (32 * CB) columns of ALL THREE FIELDS -- in
// registers for the whole time loop; warp w owns RB region rows, lane l owns CB columns of them.  Halos are exchanged
// every T steps through sentinel-armed L2 inboxes (inbox.cuh); between exchanges garbage creeps in from the region
// edges one cell per step and never reaches the tile.
//
// One time step, per thread, with ONE __syncthreads (fdtd_2d_numpy.py:7-11 order):
//   ey -= 0.5 * (hz - hz[i-1])      hz of the row above: own registers, or the bottom hz row the warp above published
//   ex -= 0.5 * (hz - hz[j-1])      hz of the left column: own registers, or the left lane's by shuffle
//   hz -= 0.7 * (ex[j+1] - ex + ey[i+1] - ey)   with the NEW ex / ey:
//        ex[j+1]: own new value, or the right lane's NEW first column by a second shuffle (warp-synchronous, no barrier)
//        ey[i+1]: own new value, or -- across the warp edge -- recomputed from the OLD top rows (ey and hz) that the warp
//                 below published in the previous step: ey_old[i+1] - 0.5 * (hz_old[i+1] - hz_old[i]) is exactly what
//                 that warp computes (row i + 1 >= 1 is never the _fict_ row)
// so every thread publishes three rows per step (new hz top, new hz bottom, new ey top) and loads three.
// Boundary rules: ey row 0 = _fict_[t]; ex column 0, hz's last row and last column never change; region cells
// outside the grid keep whatever they hold (nothing valid reads them).  Only threads that own such cells pay.
//
// Arithmetic in NumPy order, one rounding per operation (-fmad=false), as fdtd2d_step_kernel.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "inbox.cuh"

namespace f2rt {

constexpr int SLOTS = 8;          // inbox ring depth, in exchanges
constexpr int FENCE_EVERY = 4;    // gpu-scope fence cadence, in exchanges (2 * cadence <= SLOTS, see inbox.cuh)

struct Params {
    int nx, ny;
    int PI, PJ;                   // tiles along i and j
    int T;                        // steps per halo exchange = halo depth
    int tmax;
    int NW;                       // warps per CTA; region = (NW * RB) x (32 * CB)
    double *ex, *ey, *hz;
    const double *fict;
    unsigned long long *inbox;    // [PI * PJ][SLOTS][3 fields][region cells]
};

// grid cells [0, n) cut into `parts` nearly equal ranges
__device__ __host__ __forceinline__ void cell_range(int n, int parts, int t, int &lo, int &hi) {
    const int base = n / parts, rem = n % parts;
    lo = t * base + (t < rem ? t : rem);
    hi = lo + base + (t < rem ? 1 : 0);
}

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

__global__ void fdtd2d_inbox_arm_kernel(unsigned long long *box, size_t n) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x)
        box[w] = HR_SENTINEL;
}

// roles of a thread's cells in the cold paths, kept in shared memory (registers are for the field state)
struct Desc {
    unsigned rows, cols;   // as in jacobi2d_regtile.cuh: bits 0-7 own, 8-15 near the top / left edge, 16-23 bottom / right
    unsigned halo;         // cells (bit a * CB + b) received at every exchange
    unsigned keep_e;       // bits 0-15: ex keeps its value (column 0 or outside the grid); bits 16-31: ey likewise (outside)
    unsigned fict_hz;      // bits 0-15: ey = _fict_[t] (row 0); bits 16-31: hz keeps its value (last row / column, outside)
    int off[3];            // sends: inbox word offset of the block's cell (0, 0), field 0, slot 0, in up to 3 neighbours
    unsigned msk[3];
};

// One time step of the thread's RB x CB cells, in place (fdtd_2d_numpy.py:7-11 order).  BORDER: the warp owns cells
// with a boundary rule -- keep_e bits 0-15 ex keeps its value, 16-31 ey keeps; fict_hz bits 0-15 ey = *fict, 16-31 hz keeps.
template <int RB, int CB, bool BORDER>
__device__ __forceinline__ void step_fields(double (&ex)[RB][CB], double (&ey)[RB][CB], double (&hz)[RB][CB], const double *pub_me,
                                            int pub_row, int rp, unsigned keep_e, unsigned fict_hz, const double *fict) {
    constexpr int RC = 32 * CB;
    double hzl[RB], hz_up[CB], hz_dn[CB], ey_dn[CB], eyd[CB], exr[RB], t[RB][CB];
#pragma unroll
    for (int a = 0; a < RB; ++a) hzl[a] = __shfl_up_sync(0xffffffffu, hz[a][CB - 1], 1);
    lds_row<CB>(pub_me - RC + pub_row + rp, hz_up);                   // bottom hz row of warp w - 1
    lds_row<CB>(pub_me + RC + rp, hz_dn);                             // top hz row of warp w + 1
    lds_row<CB>(pub_me + RC + 2 * pub_row + rp, ey_dn);               // top ey row of warp w + 1
    double f = 0.0;
    if (BORDER) f = __ldg(fict);
    // :8  ey[1:, :] -= 0.5 * (hz[1:, :] - hz[:-1, :])   (:7 ey[0, :] = _fict_[t])
#pragma unroll
    for (int a = RB - 1; a >= 0; --a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = hz[a][b] - (a ? hz[a - 1][b] : hz_up[b]);
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = 0.5 * t[a][b];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const int bit = a * CB + b;
            if (!BORDER || !(((keep_e >> 16) | fict_hz) >> bit & 1u)) ey[a][b] = ey[a][b] - t[a][b];
            if (BORDER && ((fict_hz >> bit) & 1u)) ey[a][b] = f;
        }
    // the new ey of the row below the block, recomputed from the old rows the warp below published
#pragma unroll
    for (int b = 0; b < CB; ++b) eyd[b] = ey_dn[b] - 0.5 * (hz_dn[b] - hz[RB - 1][b]);
    // :9  ex[:, 1:] -= 0.5 * (hz[:, 1:] - hz[:, :-1])
#pragma unroll
    for (int b = CB - 1; b >= 0; --b)
#pragma unroll
        for (int a = 0; a < RB; ++a) t[a][b] = hz[a][b] - (b ? hz[a][b - 1] : hzl[a]);
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = 0.5 * t[a][b];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b)
            if (!BORDER || !((keep_e >> (a * CB + b)) & 1u)) ex[a][b] = ex[a][b] - t[a][b];
    // :10-11  hz[:-1, :-1] -= 0.7 * (ex[:-1, 1:] - ex[:-1, :-1] + ey[1:, :-1] - ey[:-1, :-1]), new ex / ey
#pragma unroll
    for (int a = 0; a < RB; ++a) exr[a] = __shfl_down_sync(0xffffffffu, ex[a][0], 1);
#pragma unroll
    for (int b = 0; b < CB; ++b)
#pragma unroll
        for (int a = 0; a < RB; ++a) t[a][b] = ((b < CB - 1) ? ex[a][b + 1] : exr[a]) - ex[a][b];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = t[a][b] + ((a < RB - 1) ? ey[a + 1][b] : eyd[b]);
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = t[a][b] - ey[a][b];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b) t[a][b] = 0.7 * t[a][b];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b)
            if (!BORDER || !((fict_hz >> (16 + a * CB + b)) & 1u)) hz[a][b] = hz[a][b] - t[a][b];
}

// halo cells of the three fields, polled straight into the registers and re-armed.  The loads of ALL pending cells of
// all three fields are issued before the first test: one L2 round trip per polling round, not one per field (three
// separate loops cost ~5 us more per exchange).
template <int RB, int CB>
__device__ __forceinline__ void poll_fields(double (&ex)[RB][CB], double (&ey)[RB][CB], double (&hz)[RB][CB],
                                            unsigned long long *q, int cells, unsigned halo) {
    constexpr int RC = 32 * CB;
    unsigned px = halo, py = halo, pz = halo;
    do {
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                const unsigned bit = 1u << (a * CB + b);
                if (px & bit) ex[a][b] = __longlong_as_double((long long)ld_relaxed_u64(q + a * RC + b));
                if (py & bit) ey[a][b] = __longlong_as_double((long long)ld_relaxed_u64(q + cells + a * RC + b));
                if (pz & bit) hz[a][b] = __longlong_as_double((long long)ld_relaxed_u64(q + 2 * cells + a * RC + b));
            }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                const unsigned bit = 1u << (a * CB + b);
                if ((px & bit) && (unsigned long long)__double_as_longlong(ex[a][b]) != HR_SENTINEL) {
                    st_relaxed_u64(q + a * RC + b, HR_SENTINEL);      // re-arm for exchange + SLOTS
                    px &= ~bit;
                }
                if ((py & bit) && (unsigned long long)__double_as_longlong(ey[a][b]) != HR_SENTINEL) {
                    st_relaxed_u64(q + cells + a * RC + b, HR_SENTINEL);
                    py &= ~bit;
                }
                if ((pz & bit) && (unsigned long long)__double_as_longlong(hz[a][b]) != HR_SENTINEL) {
                    st_relaxed_u64(q + 2 * cells + a * RC + b, HR_SENTINEL);
                    pz &= ~bit;
                }
            }
    } while (px | py | pz);
}

template <int RB, int CB, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) fdtd2d_regtile_kernel(Params p) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_geo[4];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NW = p.NW, RR = NW * RB, T = p.T;
    constexpr int RC = 32 * CB;
    const int cells = RR * RC;
    // row exchange area: [parity][hz top | hz bottom | ey top][NW + 2][RC]
    double *const pub = sm;
    const int pub_row = (NW + 2) * RC, pub_par = 3 * pub_row;
    Desc *const s_desc = reinterpret_cast<Desc *>(pub + 2 * pub_par);
    unsigned roles = 0;                                               // 1 boundary cells, 2 halo cells, 4 sends, 8 owns tile cells
    double ex[RB][CB], ey[RB][CB], hz[RB][CB];
    {
        const int nx = p.nx, ny = p.ny;
        const int ti = blockIdx.x / p.PJ, tj = blockIdx.x % p.PJ;
        int ilo, ihi, jlo, jhi;
        cell_range(nx, p.PI, ti, ilo, ihi);
        cell_range(ny, p.PJ, tj, jlo, jhi);
        if (tid == 0) { s_geo[0] = ilo; s_geo[1] = ihi; s_geo[2] = jlo; s_geo[3] = jhi; }
        const int gi0 = ilo - T + w * RB, gj0 = jlo - T + lane * CB;
        for (int x = tid; x < 2 * pub_par; x += blockDim.x) pub[x] = 0.0;
        Desc d{0, 0, 0, 0, 0, {0, 0, 0}, {0, 0, 0}};
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
                const unsigned bit = 1u << (a * CB + b);
                const bool inside = gi >= 0 && gi < nx && gj >= 0 && gj < ny;
                const bool own = ((d.rows >> a) & 1u) && ((d.cols >> b) & 1u);
                const bool need = gi >= ilo - T && gi < ihi + T && gj >= jlo - T && gj < jhi + T;
                if (inside && !own && need) d.halo |= bit;
                if (!inside || gj == 0) d.keep_e |= bit;
                if (!inside) d.keep_e |= bit << 16;
                if (inside && gi == 0) d.fict_hz |= bit;
                if (!inside || gi == nx - 1 || gj == ny - 1) d.fict_hz |= bit << 16;
                const long long g = (long long)gi * ny + gj;
                ex[a][b] = inside ? __ldg(p.ex + g) : 0.0;
                ey[a][b] = inside ? __ldg(p.ey + g) : 0.0;
                hz[a][b] = inside ? __ldg(p.hz + g) : 0.0;
            }
        {
            const unsigned r_top = (d.rows >> 8) & 255u, r_bot = (d.rows >> 16) & 255u, r_own = d.rows & 255u;
            const unsigned c_lft = (d.cols >> 8) & 255u, c_rgt = (d.cols >> 16) & 255u, c_own = d.cols & 255u;
            const int di = r_top ? -1 : (r_bot ? 1 : 0), dj = c_lft ? -1 : (c_rgt ? 1 : 0);
            const unsigned r_edge = r_top | r_bot, c_edge = c_lft | c_rgt;
            int ilo_n = ilo, jlo_n = jlo, hi;
            if (di) cell_range(nx, p.PI, ti + di, ilo_n, hi);
            if (dj) cell_range(ny, p.PJ, tj + dj, jlo_n, hi);
            const long long box_words = (long long)SLOTS * 3 * cells;
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
        roles = ((d.keep_e | d.fict_hz) ? 1u : 0u) | (d.halo ? 2u : 0u) | (sender ? 4u : 0u) |
                (((d.rows & 255u) && (d.cols & 255u)) ? 8u : 0u);
    }
    __syncthreads();
    // shared addresses of this thread's row-exchange cells (parity 0, kind 0 = hz top rows): own row; - RC: warp above; + RC: below
    double *const pub_me = pub + (w + 1) * RC + lane * CB;
    sts_row<CB>(pub_me, hz[0]);
    sts_row<CB>(pub_me + pub_row, hz[RB - 1]);
    sts_row<CB>(pub_me + 2 * pub_row, ey[0]);
    // Nobody overwrites the fields (after the last step) under a neighbour that is still loading its initial halos:
    // a tile's first exchange needs its neighbours' sends -- unless there is no exchange at all
    if (p.tmax <= T && gridDim.x > 1) cooperative_groups::this_grid().sync();
    __syncthreads();

    const int tmax = p.tmax;
    // warps without boundary cells run the step without the predicated boundary rules
    const bool border_warp = __any_sync(0xffffffffu, roles & 1u) != 0;
    unsigned keep_e = 0, fict_hz = 0;
    if (border_warp) { keep_e = s_desc[tid].keep_e; fict_hz = s_desc[tid].fict_hz; }
    int next_x = (T < tmax) ? T : 0;                                  // step after which the next exchange happens
    int nx_done = 0;
    for (int s = 1; s <= tmax; ++s) {
        const int rp = ((s - 1) & 1) * pub_par, wp = (s & 1) * pub_par;
        if (s == next_x && ((nx_done + 1) % FENCE_EVERY) == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");   // see jacobi2d_regtile.cuh
        if (border_warp) step_fields<RB, CB, true>(ex, ey, hz, pub_me, pub_row, rp, keep_e, fict_hz, p.fict + (s - 1));
        else step_fields<RB, CB, false>(ex, ey, hz, pub_me, pub_row, rp, 0u, 0u, nullptr);
        if (s == next_x) {
            // ---- halo exchange of all three fields (jacobi2d_regtile.cuh protocol; slot = [field][region cell])
            ++nx_done;
            const unsigned slot_sz = 3u * (unsigned)cells, box_sz = (unsigned)SLOTS * slot_sz;
            const unsigned out_off = (unsigned)(nx_done % SLOTS) * slot_sz;
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
                                if ((m >> (a * CB + b)) & 1u) {
                                    st_relaxed_f64((double *)(dst + a * RC + b), ex[a][b]);
                                    st_relaxed_f64((double *)(dst + cells + a * RC + b), ey[a][b]);
                                    st_relaxed_f64((double *)(dst + 2 * cells + a * RC + b), hz[a][b]);
                                }
                    }
                }
                if (roles & 2u) {
                    unsigned long long *const qb = box + (size_t)blockIdx.x * box_sz + out_off + (unsigned)(w * RB) * RC + lane * CB;
                    // spin on ONE cell (the last field of the thread's last cell), then collect
                    {
                        const int hb = 31 - __clz(d.halo);
                        unsigned long long *const q1 = qb + 2 * cells + (hb / CB) * RC + (hb % CB);
                        while (ld_relaxed_u64(q1) == HR_SENTINEL) {}
                    }
                    poll_fields<RB, CB>(ex, ey, hz, qb, cells, d.halo);
                }
            }
            next_x = (s + T < tmax) ? s + T : 0;
        }
        if (s == tmax && (roles & 8u)) {
            // ---- the final fields leave the chip
            const Desc d = s_desc[tid];
            const int gi0 = s_geo[0] - T + w * RB, gj0 = s_geo[2] - T + lane * CB;
            const long long g0 = (long long)gi0 * p.ny + gj0;
#pragma unroll
            for (int a = 0; a < RB; ++a)
#pragma unroll
                for (int b = 0; b < CB; ++b)
                    if (((d.rows >> a) & 1u) && ((d.cols >> b) & 1u)) {
                        const long long g = g0 + (long long)a * p.ny + b;
                        p.ex[g] = ex[a][b]; p.ey[g] = ey[a][b]; p.hz[g] = hz[a][b];
                    }
        }
        sts_row<CB>(pub_me + wp, hz[0]);
        sts_row<CB>(pub_me + wp + pub_row, hz[RB - 1]);
        sts_row<CB>(pub_me + wp + 2 * pub_row, ey[0]);
        __syncthreads();
    }
}

}  // namespace f2rt
