// adi.cu -- Polybench alternating-direction implicit solver (widening row), sm_100a.
//
// Replaces kernel(TSTEPS, N, u), npbench/benchmarks/polybench/adi/adi_numpy.py:6-54.  Per time step:
//   column sweep (:28-38)  for every column i: Thomas recurrence along j over u[j, i-1..i+1] -> v[j, i]
//   row sweep    (:40-52)  for every row i:    Thomas recurrence along j over v[i-1..i+1, j] -> u[i, j]
// N-2 independent chains of N-2 steps, each step one IEEE division on the chain
//   q_j = (rhs_j - a*q_{j-1}) / (a*p_{j-1} + b),   p_j = -c / (a*p_{j-1} + b)
// (p and the denominators do not depend on the data: computed once per call by adi_coef_kernel).
// One thread per chain, one warp per CTA so that each chain-carrying warp has an SM sub-partition to
// itself; two launches per time step (the sweeps are separated by a transpose-like dependency), the
// whole time loop replayed as one CUDA graph.  v is stored TRANSPOSED ([column][row]) by the column
// sweep, so both sweeps are the same kernel: operands read coalesced (fetched 8 steps ahead of the chain:
// an L2 round trip is ~4 chain steps), results written with stride N (stores do not stall a chain);
// q of a whole chain and the chain-independent p stay in shared memory for the back-substitution.
//
// Arithmetic order as in oracle/stencil_oracle.c: npb_oracle_adi; -fmad=false, IEEE division.
#include "common.cuh"

namespace {

struct AdiCoef { double a, b, c, d, e, f, k1, k2; };      // adi_numpy.py:20-25; k1 = 1+2d, k2 = 1+2a

constexpr int AD_D = 8;       // rows of operands fetched ahead of the chain (one L2 latency ~ 4 chain steps)

// One directional sweep for chain i = 1 + blockIdx.x*32 + lane.
//   COLUMN: operands u[j][i-1..i+1] (src = u, row-major), result v[j][i] stored transposed: dst[i*N + j]
//   ROW:    operands v[i-1..i+1][j] = src[j*N + i-1..i+1] (src = transposed v), result u[i][j]: dst[i*N + j]
// so both sweeps read src[j*N + i +- 1] (coalesced over the chains of a warp) and write dst[i*N + j].
// q of the whole chain stays in shared memory ([step][lane], conflict free), p -- identical for every chain
// -- once per CTA; when the chain is too long for shared memory both go to the global work arrays instead.
// p_j = -c / (a*p_{j-1} + b) and the denominators den_j = a*p_{j-1} + b do not depend on the data, on the chain
// or on the time step: one thread per sweep direction computes them once per call (same operations, same
// bits), which takes one of the two divisions off every chain step.  coef = [den(1..N-2) | p(1..N-2)] x 2.
__global__ void adi_coef_kernel(int N, double a, double b, double c, double d, double e, double f, double *coef) {
    const int dir = threadIdx.x;                               // 0 column sweep (a, b, c), 1 row sweep (d, e, f)
    if (dir > 1) return;
    const double ca = dir ? d : a, cb = dir ? e : b, cc = dir ? f : c;
    double *den = coef + (size_t)dir * 2 * (N - 2), *pp = den + (N - 2);
    double p = 0.0;
    for (int j = 1; j <= N - 2; ++j) {
        const double dn = ca * p + cb;
        p = (-cc) / dn;                                        // adi_numpy.py:32 / :45
        den[j - 1] = dn; pp[j - 1] = p;
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(32)
adi_sweep_kernel(int N, double ca, const double *__restrict__ coef, double k0, double k1, double k2,
                 const double *__restrict__ src, double *__restrict__ dst, double *__restrict__ Qg) {
    extern __shared__ double sm[];                             // den[N-2], p[N-2], then (SMEM) Q[(N-2)][32]
    const int lane = threadIdx.x;
    const int i = min(1 + (int)blockIdx.x * 32 + lane, N - 2); // surplus lanes shadow the last chain (no stores)
    const bool live = 1 + (int)blockIdx.x * 32 + lane <= N - 2;
    double *dens = sm, *Ps = sm + (N - 2), *Qs = sm + 2 * (size_t)(N - 2) + lane;
    double *Qc = Qg + i;                                       // global fallback: [step][chain]
    for (int w = lane; w < 2 * (N - 2); w += 32) sm[w] = coef[w];
    __syncwarp();
    double q = 1.0;                                            // q[i,0] = v[0,i] = 1.0 / u[i,0] = 1.0   (:28-30 / :41-43)
    if (live) dst[(long long)i * N + 0] = 1.0;
    double om[AD_D], oc[AD_D], op[AD_D];
#pragma unroll
    for (int r = 0; r < AD_D; ++r) {
        const long long o = (long long)min(1 + r, N - 2) * N + i;
        om[r] = src[o - 1]; oc[r] = src[o]; op[r] = src[o + 1];
    }
    for (int j0 = 1; j0 <= N - 2; j0 += AD_D) {
        double nm[AD_D], nc[AD_D], np[AD_D];
#pragma unroll
        for (int r = 0; r < AD_D; ++r) {                       // the next block's operands, a block ahead of the chain
            const long long o = (long long)min(j0 + AD_D + r, N - 2) * N + i;
            nm[r] = src[o - 1]; nc[r] = src[o]; np[r] = src[o + 1];
        }
#pragma unroll
        for (int r = 0; r < AD_D; ++r) {
            const int j = j0 + r;
            if (j <= N - 2) {
                q = (((k0 * om[r] + k1 * oc[r]) - k2 * op[r]) - ca * q) / dens[j - 1];   // :33-36 / :46-49
                if (SMEM) Qs[(size_t)(j - 1) * 32] = q;
                else Qc[(long long)j * N] = q;
            }
        }
#pragma unroll
        for (int r = 0; r < AD_D; ++r) { om[r] = nm[r]; oc[r] = nc[r]; op[r] = np[r]; }
    }
    if (SMEM) __syncwarp();
    double x = 1.0;                                            // v[N-1,i] = 1.0 / u[i,N-1] = 1.0  (:37 / :50)
    if (live) dst[(long long)i * N + N - 1] = 1.0;
    for (int j = N - 2; j >= 1; --j) {                         // :38-39 / :51-52
        const double pj = Ps[j - 1], qj = SMEM ? Qs[(size_t)(j - 1) * 32] : Qc[(long long)j * N];
        x = pj * x + qj;
        if (live) dst[(long long)i * N + j] = x;
    }
}

}  // namespace

extern "C" int npb_adi_f64(int64_t tsteps, int64_t n, double *u) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n >= 0 && n < (1 << 15), "npb_adi_f64", "extent out of range");
    NPB_ARG(tsteps >= 1, "npb_adi_f64", "TSTEPS must be >= 1 (the reference divides by it: adi_numpy.py:14)");
    if (n < 3) return 0;                          // no interior chains
    const double N = (double)n;
    const double DX = 1.0 / N, DY = 1.0 / N, DT = 1.0 / (double)tsteps;
    const double B1 = 2.0, B2 = 1.0;
    const double mul1 = (B1 * DT) / (DX * DX), mul2 = (B2 * DT) / (DY * DY);
    AdiCoef k;
    k.a = (-mul1) / 2.0; k.b = 1.0 + mul2; k.c = k.a; k.d = (-mul2) / 2.0; k.e = 1.0 + mul2; k.f = k.d;
    k.k1 = 1.0 + 2.0 * k.d; k.k2 = 1.0 + 2.0 * k.a;
    const size_t cells = (size_t)n * (size_t)n;
    const size_t smem_q = ((size_t)(n - 2) * 34) * sizeof(double), smem_c = ((size_t)(n - 2) * 2) * sizeof(double);
    const bool in_smem = smem_q + 1024 <= npb::st().smem_optin;
    const size_t smem = in_smem ? smem_q : smem_c;
    NPB_ARG(smem + 1024 <= npb::st().smem_optin, "npb_adi_f64", "N too large for the coefficient tables in shared memory");
    const size_t ncoef = 4 * (size_t)(n - 2);
    double *ws = (double *)npb::workspace(5, ((in_smem ? 1 : 2) * cells + ncoef) * sizeof(double));
    NPB_ARG(ws != nullptr, "npb_adi_f64", "out of device memory for the work arrays");
    double *vt = ws, *coef = ws + cells, *Q = in_smem ? ws : ws + cells + ncoef;
    static size_t configured = 0, configured_g = 0;
    if (in_smem && smem > configured) {
        NPB_CUDA(cudaFuncSetAttribute(adi_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (!in_smem && smem > configured_g) {
        NPB_CUDA(cudaFuncSetAttribute(adi_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_g = smem;
    }
    const unsigned grid = (unsigned)((n - 2 + 31) / 32);
    npb::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.kind = 7; key.dims[0] = tsteps; key.dims[1] = n;
    key.ptrs[0] = u; key.ptrs[1] = ws;
    const bool use_graph = tsteps >= 2;
    if (use_graph && npb::graph_replay(key)) return 0;
    const bool capturing = use_graph && npb::graph_begin();
    int rc = 0;
    adi_coef_kernel<<<1, 32, 0, npb::st().stream>>>((int)n, k.a, k.b, k.c, k.d, k.e, k.f, coef);
    npb::count_launch();
    const double *coef_col = coef, *coef_row = coef + 2 * (n - 2);
    for (int64_t t = 1; t <= tsteps && !rc; ++t) {
        // column sweep: q = ((((-d)*u[j,i-1] + (1+2d)*u[j,i]) - f*u[j,i+1]) - a*q) / (a*p + b), p = -c / (a*p + b)
        // row sweep:    q = ((((-a)*v[i-1,j] + (1+2a)*v[i,j]) - c*v[i+1,j]) - d*q) / (d*p + e), p = -f / (d*p + e)
        if (in_smem) {
            adi_sweep_kernel<true><<<grid, 32, smem, npb::st().stream>>>((int)n, k.a, coef_col, -k.d, k.k1, k.f, u, vt, Q);
            adi_sweep_kernel<true><<<grid, 32, smem, npb::st().stream>>>((int)n, k.d, coef_row, -k.a, k.k2, k.c, vt, u, Q);
        } else {
            adi_sweep_kernel<false><<<grid, 32, smem, npb::st().stream>>>((int)n, k.a, coef_col, -k.d, k.k1, k.f, u, vt, Q);
            adi_sweep_kernel<false><<<grid, 32, smem, npb::st().stream>>>((int)n, k.d, coef_row, -k.a, k.k2, k.c, vt, u, Q);
        }
        if (cudaGetLastError() != cudaSuccess) rc = npb::fail("npb_adi_f64", "kernel launch failed");
        npb::count_launch(2);
    }
    if (capturing) {
        const int rc2 = npb::graph_end_and_launch(key, rc);
        if (!rc) rc = rc2;
    }
    return rc;
}

extern "C" int npb_adi_f64_host(int64_t tsteps, int64_t n, double *u) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n >= 0, "npb_adi_f64_host", "negative extent");
    const size_t bytes = (size_t)n * (size_t)n * sizeof(double);
    if (!bytes) return 0;
    void *d = nullptr;
    int rc = npb_malloc(bytes, &d);
    if (rc) return rc;
    rc = npb_h2d(d, u, bytes);
    if (!rc) rc = npb_adi_f64(tsteps, n, (double *)d);
    if (!rc) rc = npb_d2h(u, d, bytes);
    if (!rc) rc = npb_sync();
    npb_free(d);
    return rc;
}
