// vadv.cu -- COSMO vertical advection: tridiagonal assembly + Thomas solve per
// (i,j) column, fused in one kernel (sm_100a).
//
// Replaces vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage),
// npbench/benchmarks/weather_stencils/vadv/vadv_numpy.py:9-78.
//
// Layout problem: arrays are (I,J,K) with K -- the recurrence axis --
// contiguous, so "one column per thread" has a lane stride of K*8 bytes.
// Design: one single-warp CTA owns NC consecutive columns (one contiguous
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
// Several single-warp CTAs per SM are resident (as many as shared memory
// allows), so phase A/D traffic of one overlaps the latency-bound B/C of
// the others.  No __syncthreads: one warp per CTA, __syncwarp between phases.
//
// Identities used (exact in binary64): BET_M == BET_P == 0.5, hence
// as == acol and cs == ccol-before-division; and gcv_k*0.5 == -(gav_{k+1}*0.5)
// because 0.25*w and -0.25*w differ only in sign.  Compiled with -fmad=false;
// evaluation order as in oracle/stencil_oracle.c: npb_oracle_vadv.
#include "common.cuh"

namespace {

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

__global__ void __launch_bounds__(32)
vadv_warp_kernel(VadvParams p) {
    extern __shared__ double tile[];
    const int K = p.K, NCP = p.NCP;
    double *tA = tile;                     // a_k   -> ccol_k
    double *tD = tile + (size_t)K * NCP;   // dc_k  -> dcol_k -> datacol_k
    const int lane = threadIdx.x;
    const long long col0 = (long long)blockIdx.x * p.NC;
    const int nc = (int)min((long long)p.NC, p.ncols - col0);
    const double dtr = p.dtr;
    const long long JK = (long long)p.J * K;

    // ---------------- phase A: lanes along k ------------------------------
    for (int cc = 0; cc < nc; ++cc) {
        const long long base = (col0 + cc) * K;
        const double *us = p.u_stage + base;
        const double *w0 = p.wcon + base;          // wcon[i  , j, :]
        const double *w1 = p.wcon + base + JK;     // wcon[i+1, j, :]
        const double *up = p.u_pos + base;
        const double *ut = p.utens + base;
        const double *uo = p.utens_stage + base;
        for (int k = lane; k < K; k += 32) {
            const double u_c = __ldg(us + k);
            const double upk = ldg_stream(up + k);
            const double utk = ldg_stream(ut + k);
            const double uok = ldg_stream(uo + k);
            // vadv_numpy.py:24-25 / 47-48 / 63-64
            const double d0 = (dtr * upk + utk) + uok;
            double a = 0.0, corr;
            if (k == 0) {
                // :16-23   gcv = 0.25*(wcon[i+1,k+1]+wcon[i,k+1]); cs = gcv*BET_M
                const double gcv = 0.25 * (__ldg(w1 + 1) + __ldg(w0 + 1));
                const double cs = gcv * 0.5;
                corr = (-cs) * (__ldg(us + 1) - u_c);
            } else {
                // :33,36,39 / :56-58   gav = -0.25*(wcon[i+1,k]+wcon[i,k]); as = acol = gav*0.5
                const double gav = -0.25 * (__ldg(w1 + k) + __ldg(w0 + k));
                a = gav * 0.5;
                const double t_lo = (-a) * (__ldg(us + k - 1) - u_c);
                if (k < K - 1) {
                    // :34,37,44-46
                    const double gcv = 0.25 * (__ldg(w1 + k + 1) + __ldg(w0 + k + 1));
                    const double cs = gcv * 0.5;
                    corr = t_lo - (cs * (__ldg(us + k + 1) - u_c));
                } else {
                    corr = t_lo;   // :62
                }
            }
            tA[k * NCP + cc] = a;
            tD[k * NCP + cc] = d0 + corr;
        }
    }
    __syncwarp();

    // ---------------- phase B + C: lanes along columns --------------------
    if (lane < nc) {
        double *cA = tA + lane;
        double *cD = tD + lane;
        // k = 0 : vadv_numpy.py:19-30   ccol = gcv*BET_P == -a_1
        double a_next = cA[NCP];
        double ccv = -a_next;
        double bcol = dtr - ccv;
        double divided = 1.0 / bcol;
        double c_prev = ccv * divided;
        double d_prev = cD[0] * divided;
        cA[0] = c_prev;
        cD[0] = d_prev;
        // 1 <= k <= K-2 : :32-53
        for (int k = 1; k < K - 1; ++k) {
            const double a = a_next;
            a_next = cA[(k + 1) * NCP];
            ccv = -a_next;
            bcol = (dtr - a) - ccv;
            divided = 1.0 / (bcol - c_prev * a);
            const double dc = cD[k * NCP];
            c_prev = ccv * divided;
            d_prev = (dc - d_prev * a) * divided;
            cA[k * NCP] = c_prev;
            cD[k * NCP] = d_prev;
        }
        {   // k = K-1 : :55-68
            const int k = K - 1;
            const double a = a_next;
            bcol = dtr - a;
            divided = 1.0 / (bcol - c_prev * a);
            d_prev = (cD[k * NCP] - d_prev * a) * divided;
            cD[k * NCP] = d_prev;       // datacol_{K-1} = dcol_{K-1}  (:70-73)
        }
        // back-substitution : :75-78
        double x = d_prev;
        for (int k = K - 2; k >= 0; --k) {
            x = cD[k * NCP] - cA[k * NCP] * x;
            cD[k * NCP] = x;
        }
    }
    __syncwarp();

    // ---------------- phase D: lanes along k ------------------------------
    for (int cc = 0; cc < nc; ++cc) {
        const long long base = (col0 + cc) * K;
        const double *up = p.u_pos + base;
        double *uo = p.utens_stage + base;
        for (int k = lane; k < K; k += 32)
            stg_stream(uo + k, dtr * (tD[k * NCP + cc] - __ldg(up + k)));   // :73, :78
    }
}

// Pick NC (columns per single-warp CTA) maximising resident columns per SM.
void pick_geometry(int K, size_t smem_per_sm, size_t smem_per_block, int *NC, int *NCP, size_t *bytes) {
    long best_cols = -1;
    for (int nc = 32; nc >= 8; --nc) {
        const int ncp = nc | 1;
        const size_t b = (size_t)2 * K * ncp * sizeof(double);
        if (b > smem_per_block) continue;
        long ctas = (long)(smem_per_sm / (b + 1024));
        if (ctas > 32) ctas = 32;
        const long cols = ctas * nc;
        if (cols > best_cols) { best_cols = cols; *NC = nc; *NCP = ncp; *bytes = b; }
    }
    if (best_cols < 0) { *NC = 0; }
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
    vadv_warp_kernel<<<(unsigned)nblk, 32, bytes, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("vadv_warp_kernel");
    npb::count_launch();
    return 0;
}
