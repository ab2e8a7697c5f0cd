// vadv.cu -- COSMO vertical advection: tridiagonal assembly + Thomas solve per
// (i,j) column, fused in one kernel (sm_100a).
//
// Replaces vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage),
// npbench/benchmarks/weather_stencils/vadv/vadv_numpy.py:9-78.
//
// Layout problem: arrays are (I,J,K) with K -- the recurrence axis --
// contiguous, so "one column per thread" has a lane stride of K*8 bytes.
// Design: one small CTA (4 warps) owns NC <= 32 consecutive columns (one contiguous
// NC*K*8-byte chunk per array) and runs four phases over a shared-memory tile
// [K][NCP] (NCP odd => conflict-free both for lanes-along-k and
// lanes-along-column accesses):
//   A  lanes along k, coalesced global reads of all five inputs; everything
//      that does not depend on the recurrence is computed here:
//      a_k (= acol = as) and the full right-hand side dcol_k before the
//      Thomas step; stored transposed into the tile;
//   B  lanes along columns: Thomas forward sweep (the serial divide chain),
//      overwriting the tile with ccol_k, dcol_k;
//   C  lanes along columns: back-substitution, tile <- datacol_k;
//   D  lanes along k: utens_stage = dtr*(datacol - u_pos), coalesced store.
// All HBM traffic is coalesced; ccol/dcol never leave the SM (the reference
// allocates them as full (I,J,K) temporaries, vadv_numpy.py:11-12).
// All warps of the CTA take part in the coalesced phases A and D (memory-level
// parallelism), the first NC threads run the serial phases B and C.  Several
// CTAs per SM are resident (as many as shared memory allows), so phase A/D
// traffic of one overlaps the latency-bound B/C of the others.
//
// Identities used (exact in binary64): BET_M == BET_P == 0.5, hence
// as == acol and cs == ccol-before-division; and gcv_k*0.5 == -(gav_{k+1}*0.5)
// because 0.25*w and -0.25*w differ only in sign.  Compiled with -fmad=false;
// evaluation order as in oracle/stencil_oracle.c: npb_oracle_vadv.
#include "common.cuh"

namespace {

constexpr int VA_WARPS = 4;                 // warps per CTA: all of them load/store (phases A, D),
constexpr int VA_THREADS = 32 * VA_WARPS;   // the first NC threads run the per-column solve (B, C)
constexpr int VA_U = 6;    // phase-A work items loaded per batch
constexpr int VD_U = 10;   // phase-D work items loaded per batch

struct VadvParams {
    long long ncols;      // I*J
    int K;
    int J;
    int NC;               // columns per CTA
    int NCP;              // padded (odd) tile row stride
    double dtr;
    double *utens_stage;
    const double *u_stage, *wcon, *u_pos, *utens;
};

__global__ void __launch_bounds__(VA_THREADS, 4)
vadv_warp_kernel(VadvParams p) {
    extern __shared__ double tile[];
    const int K = p.K, NCP = p.NCP;
    double *tA = tile;                     // a_k   -> ccol_k
    double *tD = tile + (size_t)K * NCP;   // dc_k  -> dcol_k -> datacol_k
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long col0 = (long long)blockIdx.x * p.NC;
    const int nc = (int)min((long long)p.NC, p.ncols - col0);
    const double dtr = p.dtr;
    const long long JK = (long long)p.J * K;

    // ---------------- phase A: lanes along k ------------------------------
    // Work items = (column, 32-level block); VA_U items are loaded as one batch so that
    // ~10*VA_U independent loads per lane are in flight (the warp is alone in its CTA, so
    // memory-level parallelism has to come from the instruction stream).  Neighbour
    // levels use clamped indices and selects instead of branches.
    const int nkb = (K + 31) >> 5;
    const int items = nc * nkb;
    for (int it0 = warp * VA_U; it0 < items; it0 += VA_WARPS * VA_U) {
        double r_um[VA_U], r_uc[VA_U], r_un[VA_U], r_wc[VA_U], r_wn[VA_U], r_d0a[VA_U], r_ut[VA_U], r_uo[VA_U];
        int r_k[VA_U], r_cc[VA_U];
#pragma unroll
        for (int u = 0; u < VA_U; ++u) {
            const int it = min(it0 + u, items - 1);
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
            if (it0 + u < items && ((it0 + u - r_cc[u] * nkb) << 5) + lane < K) {
                tA[k * NCP + r_cc[u]] = a;
                tD[k * NCP + r_cc[u]] = d0 + corr;
            }
        }
    }
    __syncthreads();

    // ---------------- phase B + C: lanes along columns --------------------
    if (threadIdx.x < nc) {
        double *cA = tA + threadIdx.x;
        double *cD = tD + threadIdx.x;
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
        double c_k = (K >= 2) ? cA[(K - 2) * NCP] : 0.0, d_k = cD[(K - 2) * NCP];
        for (int k = K - 2; k >= 0; --k) {
            const int kn = max(k - 1, 0);
            const double c_pf = cA[kn * NCP], d_pf = cD[kn * NCP];
            x = d_k - c_k * x;
            cD[k * NCP] = x;
            c_k = c_pf; d_k = d_pf;
        }
    }
    __syncthreads();

    // ---------------- phase D: lanes along k ------------------------------
    for (int it0 = warp * VD_U; it0 < items; it0 += VA_WARPS * VD_U) {
        double r_up[VD_U];
#pragma unroll
        for (int u = 0; u < VD_U; ++u) {
            const int it = min(it0 + u, items - 1);
            const int cc = it / nkb;
            const int k = min(((it - cc * nkb) << 5) + lane, K - 1);
            r_up[u] = __ldg(p.u_pos + (col0 + cc) * K + k);
        }
#pragma unroll
        for (int u = 0; u < VD_U; ++u) {
            const int it = it0 + u;
            const int cc = min(it, items - 1) / nkb;
            const int k = ((it - cc * nkb) << 5) + lane;
            if (it < items && k < K)
                stg_stream(p.utens_stage + (col0 + cc) * K + k, dtr * (tD[k * NCP + cc] - r_up[u]));   // :73, :78
        }
    }
}

// Pick NC (columns per CTA).  The forward sweep is a serial divide chain per
// column; one warp per SM sub-partition already keeps the FP64 pipe ~85% busy, so the goal
// is to maximise columns resident on up to 4 warps per SM (more CTAs only help overlap the
// load/store phases), with full-width warps preferred when shared memory allows.
void pick_geometry(int K, size_t smem_per_sm, size_t smem_per_block, int *NC, int *NCP, size_t *bytes) {
    long best_score = -1;
    *NC = 0;
    for (int nc = 32; nc >= 8; --nc) {
        const int ncp = nc | 1;
        const size_t b = (size_t)2 * K * ncp * sizeof(double);
        if (b > smem_per_block) continue;
        long ctas = (long)(smem_per_sm / (b + 1024));
        if (ctas > 8) ctas = 8;
        const long score = (ctas < 4 ? ctas : 4) * nc * 16 + ctas;   // columns on <=4 warps, then more CTAs
        if (score > best_score) { best_score = score; *NC = nc; *NCP = ncp; *bytes = b; }
    }
}

}  // namespace

extern "C" int npb_vadv_f64(int64_t I, int64_t J, int64_t K, double *utens_stage,
                            const double *u_stage, const double *wcon, const double *u_pos,
                            const double *utens, double dtr_stage) {
    NPB_REQUIRE_INIT();
    NPB_ARG(I >= 0 && J >= 0, "npb_vadv_f64", "negative extent");
    NPB_ARG(K >= 2, "npb_vadv_f64", "K must be >= 2 (the reference indexes level k+1 at k=0)");
    if (I == 0 || J == 0) return 0;
    NPB_ARG(K < (1 << 20), "npb_vadv_f64", "K too large");
    int NC = 0, NCP = 0;
    size_t bytes = 0;
    const size_t per_block = npb::st().smem_optin;
    pick_geometry((int)K, per_block + 1024, per_block, &NC, &NCP, &bytes);
    NPB_ARG(NC > 0, "npb_vadv_f64", "K too large for the shared-memory column tile");
    static size_t configured = 0;
    if (bytes > configured) {
        NPB_CUDA(cudaFuncSetAttribute(vadv_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)per_block));
        configured = per_block;
    }
    VadvParams p;
    p.ncols = I * J; p.K = (int)K; p.J = (int)J; p.NC = NC; p.NCP = NCP; p.dtr = dtr_stage;
    p.utens_stage = utens_stage; p.u_stage = u_stage; p.wcon = wcon; p.u_pos = u_pos; p.utens = utens;
    const long long nblk = (p.ncols + NC - 1) / NC;
    NPB_ARG(nblk < (1LL << 31), "npb_vadv_f64", "too many columns");
    vadv_warp_kernel<<<(unsigned)nblk, VA_THREADS, bytes, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("vadv_warp_kernel");
    npb::count_launch();
    return 0;
}
