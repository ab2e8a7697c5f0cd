// vadv.cu -- COSMO vertical advection: tridiagonal assembly + Thomas solve per
// (i,j) column, fused in one kernel (sm_100a).
//
// Replaces vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage),
// npbench/benchmarks/weather_stencils/vadv/vadv_numpy.py:9-78.
//
// Layout problem: arrays are (I,J,K) with K -- the recurrence axis --
// contiguous, so "one column per thread" has a lane stride of K*8 bytes, while
// the Thomas recurrence is a serial IEEE-divide chain (~150 cycles per level).
// Design: a column group (NC <= 32 consecutive columns = one contiguous
// NC*K*8-byte chunk per array) goes through four phases over a shared-memory
// tile [K][NCP] (NCP odd => conflict-free for lanes-along-k and for
// lanes-along-column accesses):
//   A  lanes along k, coalesced global reads of all five inputs; everything
//      that does not depend on the recurrence is computed here: a_k (= acol =
//      as) and the full right-hand side dcol_k before the Thomas step; stored
//      transposed into the tile;
//   B  lanes along columns: Thomas forward sweep, tile <- ccol_k, dcol_k;
//   C  lanes along columns: back-substitution, tile <- datacol_k;
//   D  lanes along k: utens_stage = dtr*(datacol - u_pos), coalesced store.
// ccol/dcol never leave the SM (the reference allocates them as full (I,J,K)
// temporaries, vadv_numpy.py:11-12).
//
// Scheduling (measured, profiles/): A and D are HBM phases that need many warps
// in flight; B+C is a latency chain that needs exactly one warp per group and
// keeps one SM sub-partition's FP64 pipe busy.  So ONE persistent CTA per SM
// owns up to three tiles and is warp specialised: warps 13..15 (three different
// sub-partitions, highest issue priority) are the solvers of tiles 0..2; warps 0..12 are movers that
// run phase D of a solved tile and phase A of its next group, tile after tile.
// Tiles hand over through mbarriers (loaded[t], solved[t]); no __syncthreads in
// the steady state, and HBM traffic of one tile overlaps the solves of the
// others.
//
// Identities used (exact in binary64): BET_M == BET_P == 0.5, hence
// as == acol and cs == ccol-before-division; and gcv_k*0.5 == -(gav_{k+1}*0.5)
// because 0.25*w and -0.25*w differ only in sign.  Compiled with -fmad=false;
// evaluation order as in oracle/stencil_oracle.c: npb_oracle_vadv.
#include "common.cuh"

namespace {

constexpr int VA_TILES = 3;                              // tiles (column groups in flight) per CTA
constexpr int VA_WARPS = 16;
constexpr int VA_THREADS = 32 * VA_WARPS;
constexpr int VA_MOVERS = VA_WARPS - VA_TILES;           // warps 0..12 move, warps 13..15 solve
constexpr int VA_U = 2;                                  // phase-A items per mover batch (11 loads each)
constexpr int VD_U = 6;                                  // phase-D items per mover batch (1 load each)

struct VadvParams {
    long long ncols;      // I*J
    long long ngroups;
    int K;
    int J;
    int NC;               // columns per group
    int NCP;              // padded (odd) tile row stride
    int ntiles;           // tiles actually used (<= VA_TILES)
    double dtr;
    double *utens_stage;
    const double *u_stage, *wcon, *u_pos, *utens;
    unsigned long long *trace;   // profiling aid: [ngroups][8] globaltimer stamps, or NULL
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// `backoff_ns` > 0: sleep between polls so that waiting warps do not steal issue slots from
// the solver warps that share their sub-partition.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity, unsigned backoff_ns = 0) {
    unsigned done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (backoff_ns) __nanosleep(backoff_ns);
    }
}

// ---------------- phase A (mover warps, lanes along k) -----------------------
// Work items = (column, 32-level block), dealt cyclically to the mover warps and
// processed in batches of VA_U; all loads of a batch are issued before the first use.  Neighbour levels
// use clamped indices and selects instead of branches.
__device__ __forceinline__ void phase_a(const VadvParams &p, double *tA, double *tD, long long col0, int nc,
                                        int mover, int lane) {
    const int K = p.K, NCP = p.NCP;
    const double dtr = p.dtr;
    const long long JK = (long long)p.J * K;
    const int nkb = (K + 31) >> 5;
    const int items = nc * nkb;
    // item `it` belongs to mover (it % VA_MOVERS) in phase A and in phase D alike
    for (int it0 = mover; it0 < items; it0 += VA_MOVERS * VA_U) {
        double r_um[VA_U], r_uc[VA_U], r_un[VA_U], r_wc[VA_U], r_wn[VA_U], r_d0a[VA_U], r_ut[VA_U], r_uo[VA_U];
        int r_k[VA_U], r_cc[VA_U];
#pragma unroll
        for (int u = 0; u < VA_U; ++u) {
            const int it = min(it0 + u * VA_MOVERS, items - 1);
            const int cc = it / nkb;
            const int k = min(((it - cc * nkb) << 5) + lane, K - 1);
            const int km = max(k - 1, 0), kp = min(k + 1, K - 1);
            const long long base = (col0 + cc) * K;
            r_k[u] = k; r_cc[u] = cc;
            r_um[u] = __ldg(p.u_stage + base + km);
            r_uc[u] = __ldg(p.u_stage + base + k);
            r_un[u] = __ldg(p.u_stage + base + kp);
            // wcon[i+1,j,k] + wcon[i,j,k]   (vadv_numpy.py:16,33,34,56)
            r_wc[u] = __ldg(p.wcon + base + JK + k) + __ldg(p.wcon + base + k);
            r_wn[u] = __ldg(p.wcon + base + JK + kp) + __ldg(p.wcon + base + kp);
            r_d0a[u] = ldg_stream(p.u_pos + base + k);
            r_ut[u] = ldg_stream(p.utens + base + k);
            r_uo[u] = ldg_stream(p.utens_stage + base + k);
        }
#pragma unroll
        for (int u = 0; u < VA_U; ++u) {
            const int k = r_k[u];
            // :24-25 / :47-48 / :63-64
            const double d0 = (dtr * r_d0a[u] + r_ut[u]) + r_uo[u];
            // :33,36,39   gav = -0.25*w_k ; as = acol = gav*0.5      (unused at k = 0)
            const double a = (-0.25 * r_wc[u]) * 0.5;
            // :16-17 / :34,37   gcv = 0.25*w_{k+1} ; cs = gcv*0.5     (unused at k = K-1)
            const double cs = (0.25 * r_wn[u]) * 0.5;
            const double t_lo = (-a) * (r_um[u] - r_uc[u]);
            const double t_hi = cs * (r_un[u] - r_uc[u]);
            // :23 (-cs)*(...) == -(cs*(...)) exactly ; :44-46 ; :62
            const double corr = (k == 0) ? -t_hi : ((k < K - 1) ? (t_lo - t_hi) : t_lo);
            const int itu = it0 + u * VA_MOVERS;
            if (itu < items && ((itu - r_cc[u] * nkb) << 5) + lane < K) {
                tA[k * NCP + r_cc[u]] = a;
                tD[k * NCP + r_cc[u]] = d0 + corr;
            }
        }
    }
}

// ---------------- phases B + C (one solver warp, lanes along columns) --------
__device__ __forceinline__ void phase_bc(const VadvParams &p, double *tA, double *tD, int nc, int lane) {
    const int K = p.K, NCP = p.NCP;
    const double dtr = p.dtr;
    if (lane >= nc) return;
    double *cA = tA + lane;
    double *cD = tD + lane;
    // k = 0 : vadv_numpy.py:19-30   ccol = gcv*BET_P == -a_1
    double a_cur = cA[NCP];                              // a_1
    double a_nxt = cA[min(2, K - 1) * NCP];              // a_2
    double dc_cur = cD[min(1, K - 1) * NCP];             // dc_1
    double ccv = -a_cur;
    double bcol = dtr - ccv;
    double divided = 1.0 / bcol;
    double c_prev = ccv * divided;
    double d_prev = cD[0] * divided;
    cA[0] = c_prev;
    cD[0] = d_prev;
    // 1 <= k <= K-2 : :32-53.  Operands of level k+1 are fetched before the stores of
    // level k, so their shared-memory latency hides under the divide chain.
    for (int k = 1; k < K - 1; ++k) {
        const double a_pf = cA[min(k + 2, K - 1) * NCP];
        const double dc_pf = cD[(k + 1) * NCP];
        const double a = a_cur;
        ccv = -a_nxt;
        bcol = (dtr - a) - ccv;
        divided = 1.0 / (bcol - c_prev * a);
        c_prev = ccv * divided;
        d_prev = (dc_cur - d_prev * a) * divided;
        cA[k * NCP] = c_prev;
        cD[k * NCP] = d_prev;
        a_cur = a_nxt; a_nxt = a_pf; dc_cur = dc_pf;
    }
    {   // k = K-1 : :55-68
        const double a = a_cur;
        bcol = dtr - a;
        divided = 1.0 / (bcol - c_prev * a);
        d_prev = (dc_cur - d_prev * a) * divided;
        cD[(K - 1) * NCP] = d_prev;       // datacol_{K-1} = dcol_{K-1}  (:70-73)
    }
    // back-substitution : :75-78
    double x = d_prev;
    double c_k = cA[(K - 2) * NCP], d_k = cD[(K - 2) * NCP];
    for (int k = K - 2; k >= 0; --k) {
        const int kn = max(k - 1, 0);
        const double c_pf = cA[kn * NCP], d_pf = cD[kn * NCP];
        x = d_k - c_k * x;
        cD[k * NCP] = x;
        c_k = c_pf; d_k = d_pf;
    }
}

// ---------------- phase D (mover warps, lanes along k) -----------------------
// Same item -> warp map as phase A: a mover warp reads in D exactly the tile cells it
// overwrites in the following phase A of the same tile, so D -> A needs no barrier.
__device__ __forceinline__ void phase_d(const VadvParams &p, const double *tD, long long col0, int nc, int mover,
                                        int lane) {
    const int K = p.K, NCP = p.NCP;
    const double dtr = p.dtr;
    const int nkb = (K + 31) >> 5;
    const int items = nc * nkb;
    for (int it0 = mover; it0 < items; it0 += VA_MOVERS * VD_U) {
        double r_up[VD_U];
#pragma unroll
        for (int u = 0; u < VD_U; ++u) {
            const int it = min(it0 + u * VA_MOVERS, items - 1);
            const int cc = it / nkb;
            const int k = min(((it - cc * nkb) << 5) + lane, K - 1);
            r_up[u] = __ldg(p.u_pos + (col0 + cc) * K + k);
        }
#pragma unroll
        for (int u = 0; u < VD_U; ++u) {
            const int it = it0 + u * VA_MOVERS;
            const int cc = min(it, items - 1) / nkb;
            const int k = ((it - cc * nkb) << 5) + lane;
            if (it < items && k < K)
                stg_stream(p.utens_stage + (col0 + cc) * K + k, dtr * (tD[k * NCP + cc] - r_up[u]));   // :73, :78
        }
    }
}

__device__ __forceinline__ void stamp(const VadvParams &p, long long g, int slot, bool who) {
    if (p.trace && who) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.trace[g * 8 + slot] = t;
    }
}

__global__ void __launch_bounds__(VA_THREADS, 1)
vadv_pipeline_kernel(VadvParams p) {
    extern __shared__ double tiles[];
    __shared__ __align__(8) unsigned long long loaded[VA_TILES], solved[VA_TILES];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t tile_sz = (size_t)2 * p.K * p.NCP;          // doubles per tile: a/ccol plane + dcol plane
    if (threadIdx.x == 0) {
        for (int t = 0; t < VA_TILES; ++t) { mbar_init(&loaded[t], VA_MOVERS); mbar_init(&solved[t], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tile t of CTA b works on groups  b*ntiles + t + n * (ntiles * gridDim.x),  n = 0, 1, ...
    const long long stride = (long long)p.ntiles * gridDim.x;
    const long long first = (long long)blockIdx.x * p.ntiles;

    // Measured (tools/vadv_trace.py): a solve takes 12-13 us when its sub-partition is otherwise
    // idle and ~21 us when mover warps share it (FP64 pipe / issue slots); giving the movers only
    // sub-partition 0 fixes the solve but starves phases A/D (4 movers).  Next step: feed phase A
    // with TMA bulk copies so that few mover warps suffice.
    if (warp >= VA_MOVERS) {
        // ------------------------------- solver of tile `warp - VA_MOVERS` ---------
        const int t = warp - VA_MOVERS;
        if (t >= p.ntiles) return;
        double *tA = tiles + t * tile_sz, *tD = tA + (size_t)p.K * p.NCP;
        unsigned n = 0;
        for (long long g = first + t; g < p.ngroups; g += stride, ++n) {
            const int nc = (int)min((long long)p.NC, p.ncols - g * p.NC);
            mbar_wait(&loaded[t], n & 1u);
            stamp(p, g, 2, lane == 0);
            phase_bc(p, tA, tD, nc, lane);
            __syncwarp();
            stamp(p, g, 3, lane == 0);
            if (lane == 0) mbar_arrive(&solved[t]);
        }
        return;
    }

    // ----------------------------------- movers ------------------------------------
    const int mover = warp;
    for (int t = 0; t < p.ntiles; ++t) {                      // prologue: first group of every tile
        const long long g = first + t;
        if (g >= p.ngroups) break;
        double *tA = tiles + t * tile_sz, *tD = tA + (size_t)p.K * p.NCP;
        stamp(p, g, 0, mover == 0 && lane == 0);
        phase_a(p, tA, tD, g * p.NC, (int)min((long long)p.NC, p.ncols - g * p.NC), mover, lane);
        __syncwarp();
        stamp(p, g, 1, mover == 0 && lane == 0);
        if (lane == 0) mbar_arrive(&loaded[t]);
    }
    for (unsigned n = 0;; ++n) {
        bool any = false;
        for (int t = 0; t < p.ntiles; ++t) {
            const long long g = first + t + (long long)n * stride;
            if (g >= p.ngroups) continue;
            any = true;
            double *tA = tiles + t * tile_sz, *tD = tA + (size_t)p.K * p.NCP;
            mbar_wait(&solved[t], n & 1u, 256);
            stamp(p, g, 4, mover == 0 && lane == 0);
            phase_d(p, tD, g * p.NC, (int)min((long long)p.NC, p.ncols - g * p.NC), mover, lane);
            stamp(p, g, 5, mover == 0 && lane == 0);
            const long long g2 = g + stride;
            if (g2 < p.ngroups) {
                stamp(p, g2, 0, mover == 0 && lane == 0);
                phase_a(p, tA, tD, g2 * p.NC, (int)min((long long)p.NC, p.ncols - g2 * p.NC), mover, lane);
                __syncwarp();
                stamp(p, g2, 1, mover == 0 && lane == 0);
                if (lane == 0) mbar_arrive(&loaded[t]);
            }
        }
        if (!any) break;
    }
}

// ---------------------------------------------------------------------------
// TMA-fed variant (even K, opt-in via npb_vadv_set_mode(2)): same tiles, same solvers, but phase A no longer
// issues global loads from the mover warps.  Measured on the kernel above
// (tools/vadv_trace.py): (1) per-SM LDG traffic saturates at ~33 GB/s (outstanding
// L1 miss lines), which makes the movers the bottleneck; (2) a solve takes 12-13 us
// when its SM sub-partition is otherwise idle but ~21 us when mover warps share
// it.  So here one producer thread streams the raw inputs of CS columns at a
// time into a double-buffered staging area with cp.async.bulk (TMA), the three
// solver warps sit alone on sub-partitions 1..3, and the seven mover warps --
// all on sub-partition 0 -- only do shared-memory work: staging -> (a_k, dcol_k)
// -> tile, and tile -> utens_stage.
// ---------------------------------------------------------------------------
constexpr int VT_CS = 4;                     // columns per staging chunk
constexpr int VT_STAGES = 2;
constexpr int VT_MOVERS = 7;                 // warps 0,4,...,24 (sub-partition 0)
constexpr int VT_WARPS = 32;                 // warp 28 = producer; warps 1,2,3 = solvers; others exit
constexpr int VT_PRODUCER = 4 * VT_MOVERS;
constexpr int VT_THREADS = 32 * VT_WARPS;
constexpr int VT_DU = 4;                     // phase-D items per batch

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// phase A from staging: chunk of `ccols` columns starting at group-local column c0
__device__ __forceinline__ void phase_a_staged(const VadvParams &p, const double *stg, double *tA, double *tD,
                                               int c0, int ccols, int mover, int lane) {
    const int K = p.K, NCP = p.NCP;
    const double dtr = p.dtr;
    const int nkb = (K + 31) >> 5;
    const int cstride = VT_CS * K;               // doubles per array in a stage
    const double *s_uo = stg, *s_us = stg + cstride, *s_w0 = stg + 2 * cstride, *s_w1 = stg + 3 * cstride,
                 *s_up = stg + 4 * cstride, *s_ut = stg + 5 * cstride;
    for (int it = mover; it < ccols * nkb; it += VT_MOVERS) {
        const int cl = it / nkb;
        const int k = ((it - cl * nkb) << 5) + lane;
        if (k >= K) continue;
        const int km = max(k - 1, 0), kp = min(k + 1, K - 1);
        const int o = cl * K;
        const double u_c = s_us[o + k];
        const double wc = s_w1[o + k] + s_w0[o + k];          // wcon[i+1,j,k] + wcon[i,j,k]
        const double wn = s_w1[o + kp] + s_w0[o + kp];
        const double d0 = (dtr * s_up[o + k] + s_ut[o + k]) + s_uo[o + k];
        const double a = (-0.25 * wc) * 0.5;
        const double cs = (0.25 * wn) * 0.5;
        const double t_lo = (-a) * (s_us[o + km] - u_c);
        const double t_hi = cs * (s_us[o + kp] - u_c);
        const double corr = (k == 0) ? -t_hi : ((k < K - 1) ? (t_lo - t_hi) : t_lo);
        tA[k * NCP + c0 + cl] = a;
        tD[k * NCP + c0 + cl] = d0 + corr;
    }
}

__device__ __forceinline__ void phase_d4(const VadvParams &p, const double *tD, long long col0, int nc, int mover,
                                         int lane) {
    const int K = p.K, NCP = p.NCP;
    const double dtr = p.dtr;
    const int nkb = (K + 31) >> 5;
    const int items = nc * nkb;
    for (int it0 = mover; it0 < items; it0 += VT_MOVERS * VT_DU) {
        double r_up[VT_DU];
#pragma unroll
        for (int u = 0; u < VT_DU; ++u) {
            const int it = min(it0 + u * VT_MOVERS, items - 1);
            const int cc = it / nkb;
            const int k = min(((it - cc * nkb) << 5) + lane, K - 1);
            r_up[u] = __ldg(p.u_pos + (col0 + cc) * K + k);
        }
#pragma unroll
        for (int u = 0; u < VT_DU; ++u) {
            const int it = it0 + u * VT_MOVERS;
            const int cc = min(it, items - 1) / nkb;
            const int k = ((it - cc * nkb) << 5) + lane;
            if (it < items && k < K)
                stg_stream(p.utens_stage + (col0 + cc) * K + k, dtr * (tD[k * NCP + cc] - r_up[u]));
        }
    }
}

__global__ void __launch_bounds__(VT_THREADS, 1)
vadv_tma_kernel(VadvParams p) {
    extern __shared__ __align__(128) double smem_all[];
    __shared__ __align__(8) unsigned long long loaded[VA_TILES], solved[VA_TILES], full_bar[VT_STAGES],
        empty_bar[VT_STAGES];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int K = p.K;
    const size_t stage_sz = (size_t)6 * VT_CS * K;            // doubles per stage
    double *stages = smem_all;
    double *tiles = smem_all + VT_STAGES * stage_sz;
    const size_t tile_sz = (size_t)2 * K * p.NCP;
    if (threadIdx.x == 0) {
        for (int t = 0; t < VA_TILES; ++t) { mbar_init(&loaded[t], VT_MOVERS); mbar_init(&solved[t], 1); }
        for (int s = 0; s < VT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], VT_MOVERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long stride = (long long)p.ntiles * gridDim.x;
    const long long first = (long long)blockIdx.x * p.ntiles;
    const long long JK = (long long)p.J * K;

    if (warp >= 1 && warp <= 3) {
        // ------------------------------- solver of tile `warp - 1` (alone on its sub-partition)
        const int t = warp - 1;
        if (t >= p.ntiles) return;
        double *tA = tiles + t * tile_sz, *tD = tA + (size_t)K * p.NCP;
        unsigned n = 0;
        for (long long g = first + t; g < p.ngroups; g += stride, ++n) {
            const int nc = (int)min((long long)p.NC, p.ncols - g * p.NC);
            mbar_wait(&loaded[t], n & 1u, 64);
            stamp(p, g, 2, lane == 0);
            phase_bc(p, tA, tD, nc, lane);
            __syncwarp();
            stamp(p, g, 3, lane == 0);
            if (lane == 0) mbar_arrive(&solved[t]);
        }
        return;
    }

    // The A-jobs are visited in the same order by the producer and by the movers:
    //   prologue: (tile t, group first+t) for t = 0..ntiles-1;
    //   then for n = 0,1,..: for t: if group g(t,n) exists and g(t,n+1) exists -> A-job (t, g(t,n+1)).
    if (warp == VT_PRODUCER) {
        // ------------------------------- producer: one thread streams chunks with TMA
        if (lane != 0) return;
        unsigned job = 0;
        auto feed = [&](long long g) {
            const int nc = (int)min((long long)p.NC, p.ncols - g * p.NC);
            for (int c0 = 0; c0 < nc; c0 += VT_CS, ++job) {
                const int cc = min(VT_CS, nc - c0);
                const int st = job % VT_STAGES;
                mbar_wait(&empty_bar[st], ((job / VT_STAGES) & 1u) ^ 1u, 32);
                const unsigned bytes = (unsigned)(cc * K) * 8u;
                double *dst = stages + st * stage_sz;
                const long long base = (g * p.NC + c0) * K;
                mbar_arrive_expect_tx(&full_bar[st], 6u * bytes);
                bulk_g2s(dst + 0 * VT_CS * K, p.utens_stage + base, bytes, &full_bar[st]);
                bulk_g2s(dst + 1 * VT_CS * K, p.u_stage + base, bytes, &full_bar[st]);
                bulk_g2s(dst + 2 * VT_CS * K, p.wcon + base, bytes, &full_bar[st]);
                bulk_g2s(dst + 3 * VT_CS * K, p.wcon + base + JK, bytes, &full_bar[st]);
                bulk_g2s(dst + 4 * VT_CS * K, p.u_pos + base, bytes, &full_bar[st]);
                bulk_g2s(dst + 5 * VT_CS * K, p.utens + base, bytes, &full_bar[st]);
            }
        };
        for (int t = 0; t < p.ntiles; ++t)
            if (first + t < p.ngroups) feed(first + t);
        for (unsigned n = 0;; ++n) {
            bool any = false;
            for (int t = 0; t < p.ntiles; ++t) {
                const long long g = first + t + (long long)n * stride;
                if (g >= p.ngroups) continue;
                any = true;
                if (g + stride < p.ngroups) feed(g + stride);
            }
            if (!any) break;
        }
        return;
    }
    if ((warp & 3) != 0 || warp >= VT_PRODUCER) return;       // warps without a role

    // ----------------------------------- movers (warps 0,4,..: sub-partition 0) ------
    const int mover = warp >> 2;
    unsigned job = 0;
    auto load_tile = [&](int t, long long g) {
        double *tA = tiles + t * tile_sz, *tD = tA + (size_t)K * p.NCP;
        const int nc = (int)min((long long)p.NC, p.ncols - g * p.NC);
        stamp(p, g, 0, mover == 0 && lane == 0);
        for (int c0 = 0; c0 < nc; c0 += VT_CS, ++job) {
            const int st = job % VT_STAGES;
            mbar_wait(&full_bar[st], (job / VT_STAGES) & 1u, 32);
            phase_a_staged(p, stages + st * stage_sz, tA, tD, c0, min(VT_CS, nc - c0), mover, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
        stamp(p, g, 1, mover == 0 && lane == 0);
        if (lane == 0) mbar_arrive(&loaded[t]);
    };
    for (int t = 0; t < p.ntiles; ++t)
        if (first + t < p.ngroups) load_tile(t, first + t);
    for (unsigned n = 0;; ++n) {
        bool any = false;
        for (int t = 0; t < p.ntiles; ++t) {
            const long long g = first + t + (long long)n * stride;
            if (g >= p.ngroups) continue;
            any = true;
            const double *tD = tiles + t * tile_sz + (size_t)K * p.NCP;
            mbar_wait(&solved[t], n & 1u, 128);
            stamp(p, g, 4, mover == 0 && lane == 0);
            phase_d4(p, tD, g * p.NC, (int)min((long long)p.NC, p.ncols - g * p.NC), mover, lane);
            stamp(p, g, 5, mover == 0 && lane == 0);
            if (g + stride < p.ngroups) load_tile(t, g + stride);
        }
        if (!any) break;
    }
}

// Geometry of the TMA variant: 3 tiles + 2 staging buffers; false if the groups would be too narrow.
bool pick_geometry_tma(int K, size_t smem_per_block, int *NC, int *NCP, size_t *bytes) {
    if (K & 1) return false;                                  // bulk copies need 16-byte aligned columns
    const size_t staging = (size_t)VT_STAGES * 6 * VT_CS * K * sizeof(double);
    for (int nc = 31; nc >= 12; --nc) {
        if ((nc & 1) == 0) continue;                          // odd width: no padding column wasted
        const size_t b = staging + (size_t)VA_TILES * 2 * K * nc * sizeof(double);
        if (b + 512 <= smem_per_block) { *NC = nc; *NCP = nc; *bytes = b; return true; }
    }
    return false;
}

// Geometry: as many tiles (<= 3) as possible with the widest column groups that fit.
bool pick_geometry(int K, size_t smem_per_block, int *ntiles, int *NC, int *NCP, size_t *bytes) {
    for (int nt = VA_TILES; nt >= 1; --nt)
        for (int nc = 32; nc >= (nt > 1 ? 12 : 1); --nc) {
            const int ncp = nc | 1;
            const size_t b = (size_t)nt * 2 * K * ncp * sizeof(double);
            if (b + 256 <= smem_per_block) { *ntiles = nt; *NC = nc; *NCP = ncp; *bytes = b; return true; }
        }
    return false;
}

unsigned long long *g_trace = nullptr;
int g_vadv_mode = 0;     // 0 dispatch (streaming solver when eligible), 1 tile kernel, 2 TMA-fed tile kernel,
                         // 3/4/5 streaming solver variants 1/2/3 (vadv_stream.cu)
int g_vadv_last = 0;     // 1 tile kernel, 2 TMA-fed tile kernel, 3 streaming solver

}  // namespace

namespace npb {
int vadv_stream_launch(int variant, int64_t I, int64_t J, int64_t K, double *utens_stage, const double *u_stage,
                       const double *wcon, const double *u_pos, const double *utens, double dtr, unsigned long long *trace);
}

// profiling aid: device buffer of ngroups*8 u64 receiving per-group phase timestamps (NULL = off)
// 0/1: LDG pipeline kernel (default, fastest measured); 2: TMA-fed kernel for even K with enough columns
// (solves run undisturbed at 13 us instead of 21 us, but the staged phase A is slower: 309 vs 220 us at `paper`)
extern "C" int npb_vadv_set_mode(int mode) { g_vadv_mode = mode; return 0; }
extern "C" int npb_vadv_last_path(void) { return g_vadv_last; }
extern "C" int npb_vadv_set_trace(void *dev_buf) { g_trace = (unsigned long long *)dev_buf; return 0; }

extern "C" int npb_vadv_f64(int64_t I, int64_t J, int64_t K, double *utens_stage,
                            const double *u_stage, const double *wcon, const double *u_pos,
                            const double *utens, double dtr_stage) {
    NPB_REQUIRE_INIT();
    NPB_ARG(I >= 0 && J >= 0, "npb_vadv_f64", "negative extent");
    NPB_ARG(K >= 2, "npb_vadv_f64", "K must be >= 2 (the reference indexes level k+1 at k=0)");
    if (I == 0 || J == 0) return 0;
    NPB_ARG(K < (1 << 20), "npb_vadv_f64", "K too large");
    if (g_vadv_mode == 0 || g_vadv_mode >= 3) {
        const int rc = npb::vadv_stream_launch(g_vadv_mode == 0 ? 0 : g_vadv_mode - 2, I, J, K, utens_stage, u_stage,
                                               wcon, u_pos, utens, dtr_stage, g_trace);
        if (rc < 0) return npb::fail_cuda("vadv stream kernel", cudaGetLastError());
        if (rc == 1) { g_vadv_last = 3; return 0; }
    }
    int ntiles = 0, NC = 0, NCP = 0;
    size_t bytes = 0;
    const int sms_ = npb::st().sm_count;
    bool use_tma = (g_vadv_mode == 2) && pick_geometry_tma((int)K, npb::st().smem_optin, &NC, &NCP, &bytes) &&
                   (I * J + NC - 1) / NC >= (long long)VA_TILES * sms_;       // enough groups for 3 tiles per SM
    if (use_tma) ntiles = VA_TILES;
    else
    NPB_ARG(pick_geometry((int)K, npb::st().smem_optin, &ntiles, &NC, &NCP, &bytes), "npb_vadv_f64",
            "K too large for the shared-memory column tile");
    g_vadv_last = use_tma ? 2 : 1;
    static size_t cfg[NPB_MAX_DEVICES] = {0}, cfg_tma[NPB_MAX_DEVICES] = {0};      // per device
    size_t &configured = cfg[npb::cur_device()], &configured_tma = cfg_tma[npb::cur_device()];
    if (!use_tma && bytes > configured) {
        NPB_CUDA(cudaFuncSetAttribute(vadv_pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    if (use_tma && bytes > configured_tma) {
        NPB_CUDA(cudaFuncSetAttribute(vadv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured_tma = bytes;
    }
    VadvParams p;
    p.ncols = I * J; p.K = (int)K; p.J = (int)J; p.NC = NC; p.NCP = NCP; p.dtr = dtr_stage;
    p.utens_stage = utens_stage; p.u_stage = u_stage; p.wcon = wcon; p.u_pos = u_pos; p.utens = utens;
    p.ngroups = (p.ncols + NC - 1) / NC;
    NPB_ARG(p.ngroups < (1LL << 31), "npb_vadv_f64", "too many columns");
    // few groups: fewer tiles per CTA so that every SM gets work
    const int sms = npb::st().sm_count;
    while (ntiles > 1 && p.ngroups < (long long)ntiles * sms) --ntiles;
    p.ntiles = ntiles;
    p.trace = g_trace;
    long long grid = (p.ngroups + ntiles - 1) / ntiles;
    if (grid > sms) grid = sms;
    if (use_tma) vadv_tma_kernel<<<(unsigned)grid, VT_THREADS, bytes, npb::st().stream>>>(p);
    else vadv_pipeline_kernel<<<(unsigned)grid, VA_THREADS, bytes, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("vadv kernel");
    npb::count_launch();
    return 0;
}
