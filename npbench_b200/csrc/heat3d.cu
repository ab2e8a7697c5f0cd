// heat3d.cu -- 7-point 3-D heat equation sweeps (sm_100a).
//
// Replaces kernel(TSTEPS, A, B),
// npbench/benchmarks/polybench/heat_3d/heat_3d_numpy.py:4-20.
//
// v1 design: one launch per sweep.  The (j,k) plane is flattened to one
// contiguous axis (border cells included, their threads idle), so accesses are
// unit-stride for any extent; each thread owns one (j,k) column and marches
// along i over a chunk of planes with a 3-plane register window; the four
// in-plane neighbours come through L1.  At the NPBench presets the whole grid
// is L2 resident (70^3 * 8 B = 2.7 MB) and the sweep is latency/launch bound.
//
// Arithmetic (NumPy order, one rounding per op; -fmad=false):
//   ((0.125*((A[i+1]-2c)+A[i-1]) + 0.125*((A[j+1]-2c)+A[j-1]))
//     + 0.125*((A[k+1]-2c)+A[k-1])) + c
#include "common.cuh"

namespace {

constexpr int H3_THREADS = 256;

__global__ void __launch_bounds__(H3_THREADS)
heat3d_sweep_kernel(int n1, int n2, long long plane, const double *__restrict__ src,
                    double *__restrict__ dst, long long i_lo, long long i_hi, int planes_per_chunk) {
    const long long c = (long long)blockIdx.x * H3_THREADS + threadIdx.x;   // flat (j,k)
    if (c >= plane) return;
    const int j = (int)(c / n2), k = (int)(c % n2);
    if (j < 1 || j > n1 - 2 || k < 1 || k > n2 - 2) return;
    const long long ia = i_lo + (long long)blockIdx.y * planes_per_chunk;
    const long long ib = min(i_hi, ia + planes_per_chunk);
    if (ia >= ib) return;
    const double *p = src + ia * plane + c;
    double *o = dst + ia * plane + c;
    double up = __ldg(p - plane);
    double ce = __ldg(p);
    for (long long i = ia; i < ib; ++i) {
        const double dn = __ldg(p + plane);
        const double jm = __ldg(p - n2), jp = __ldg(p + n2);
        const double km = __ldg(p - 1), kp = __ldg(p + 1);
        const double c2 = 2.0 * ce;
        const double t1 = 0.125 * ((dn - c2) + up);
        const double t2 = 0.125 * ((jp - c2) + jm);
        const double t3 = 0.125 * ((kp - c2) + km);
        *o = ((t1 + t2) + t3) + ce;
        up = ce; ce = dn;
        p += plane; o += plane;
    }
}

int launch_sweep(int64_t n0, int64_t n1, int64_t n2, const double *src, double *dst, int64_t i_lo,
                 int64_t i_hi) {
    if (i_lo < 1) i_lo = 1;
    if (i_hi < 0 || i_hi > n0 - 1) i_hi = n0 - 1;
    if (i_lo >= i_hi) return 0;
    const long long plane = (long long)n1 * n2;
    const long long col_blocks = (plane + H3_THREADS - 1) / H3_THREADS;
    const long long planes = i_hi - i_lo;
    long long ppc = 32;   // planes per chunk: shrink until the machine is filled ~4x
    while (ppc > 2 && col_blocks * ((planes + ppc - 1) / ppc) < 4LL * npb::st().sm_count) ppc >>= 1;
    long long chunks = (planes + ppc - 1) / ppc;
    while (chunks > 65535) { ppc <<= 1; chunks = (planes + ppc - 1) / ppc; }
    dim3 grid((unsigned)col_blocks, (unsigned)chunks);
    heat3d_sweep_kernel<<<grid, H3_THREADS, 0, npb::st().stream>>>(
        (int)n1, (int)n2, plane, src, dst, (long long)i_lo, (long long)i_hi, (int)ppc);
    NPB_CHECK_LAUNCH("heat3d_sweep_kernel");
    npb::count_launch();
    return 0;
}

}  // namespace

extern "C" int npb_heat3d_sweep_f64(int64_t n0, int64_t n1, int64_t n2, const double *src,
                                    double *dst, int64_t i_lo, int64_t i_hi) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n0 >= 0 && n1 >= 0 && n2 >= 0, "npb_heat3d_sweep_f64", "negative extent");
    NPB_ARG(n1 < (1LL << 31) && n2 < (1LL << 31) && n1 * n2 < (1LL << 40), "npb_heat3d_sweep_f64",
            "plane too large");
    if (n0 < 3 || n1 < 3 || n2 < 3) return 0;
    return launch_sweep(n0, n1, n2, src, dst, i_lo, i_hi);
}

extern "C" int npb_heat3d_f64(int64_t tsteps, int64_t n0, int64_t n1, int64_t n2, double *A,
                              double *B) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n0 >= 0 && n1 >= 0 && n2 >= 0, "npb_heat3d_f64", "negative extent");
    NPB_ARG(n1 < (1LL << 31) && n2 < (1LL << 31) && n1 * n2 < (1LL << 40), "npb_heat3d_f64",
            "plane too large");
    if (tsteps <= 1 || n0 < 3 || n1 < 3 || n2 < 3) return 0;
    for (int64_t t = 1; t < tsteps; ++t) {   // heat_3d_numpy.py:6
        int rc = launch_sweep(n0, n1, n2, A, B, 1, n0 - 1);
        if (rc) return rc;
        rc = launch_sweep(n0, n1, n2, B, A, 1, n0 - 1);
        if (rc) return rc;
    }
    return 0;
}
