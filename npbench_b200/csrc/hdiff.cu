// hdiff.cu -- COSMO horizontal diffusion, fused single pass (sm_100a).
//
// Replaces hdiff(in_field, out_field, coeff),
// npbench/benchmarks/weather_stencils/hdiff/hdiff_numpy.py:5-29: the reference
// materialises lap_field, flx_field and fly_field as full-size temporaries and
// makes 21 ufunc passes; here Laplacian, both flux limiters and the output
// stage are fused in registers -- `in` and `coeff` are read once, `out` is
// written once, nothing else touches HBM.
//
// Layout: in (I+4, J+4, K), out/coeff (I, J, K), K contiguous.  (j,k) is
// flattened to one contiguous axis: output column c = j*K + k reads the input
// row at c + 2K + dq*K for dq in -2..2, so every access is unit-stride across
// lanes for any K.  Each thread owns one flattened column and marches along i
// with a rolling register window (13-point footprint; 5 new loads per output);
// the Laplacians of column q and the x-flux are carried from row to row.
//
// Arithmetic follows NumPy's evaluation order exactly (see oracle/
// stencil_oracle.c: npb_oracle_hdiff) and the TU is compiled with -fmad=false.
#include "common.cuh"

namespace {

constexpr int HD_THREADS = 256;

__device__ __forceinline__ double lap5(double c, double ip, double im, double jp, double jm) {
    // hdiff_numpy.py:7-9   4*in[c] - (in[i+1] + in[i-1] + in[j+1] + in[j-1])
    return 4.0 * c - (((ip + im) + jp) + jm);
}

__device__ __forceinline__ double limit(double r, double din) {
    // hdiff_numpy.py:12-17 / 20-25   np.where(res * d_in > 0, 0, res)
    return (r * din > 0.0) ? 0.0 : r;
}

// rows_per_chunk output rows per thread; grid.y = number of chunks.
__global__ void __launch_bounds__(HD_THREADS)
hdiff_march_kernel(int I, int JK, int K, long long in_pitch,   // in_pitch = (J+4)*K
                   const double *__restrict__ in, double *__restrict__ out,
                   const double *__restrict__ coeff, int rows_per_chunk) {
    const int c = blockIdx.x * HD_THREADS + threadIdx.x;
    if (c >= JK) return;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int i1 = min(I, i0 + rows_per_chunk);
    if (i0 >= i1) return;

    // pointer to in[(i0+2), q, k] : centre of the first output row
    const double *pc = in + (long long)(i0 + 2) * in_pitch + c + 2 * (long long)K;
    const long long P = in_pitch;

    // window: column q rows p-2..p+2; columns q-1/q+1 rows p-1..p+1; q-2/q+2 row p
    double c_m2 = __ldg(pc - 2 * P), c_m1 = __ldg(pc - P), c_0 = __ldg(pc), c_p1 = __ldg(pc + P),
           c_p2 = __ldg(pc + 2 * P);
    double l_m1 = __ldg(pc - P - K), l_0 = __ldg(pc - K), l_p1 = __ldg(pc + P - K);
    double r_m1 = __ldg(pc - P + K), r_0 = __ldg(pc + K), r_p1 = __ldg(pc + P + K);
    double ll_0 = __ldg(pc - 2 * K), rr_0 = __ldg(pc + 2 * K);

    double lap_m = lap5(c_m1, c_0, c_m2, r_m1, l_m1);   // lap(p-1, q)
    double lap_c = lap5(c_0, c_p1, c_m1, r_0, l_0);     // lap(p,   q)
    double flx_m = limit(lap_c - lap_m, c_0 - c_m1);    // flux between (p-1,q) and (p,q)

    const double *pco = coeff + (long long)i0 * JK + c;
    double *po = out + (long long)i0 * JK + c;

    for (int i = i0; i < i1; ++i) {
        // prefetch the 5 new values of the next row's window (rows exist up to I+3)
        double n_c_p2 = 0.0, n_l_p1 = 0.0, n_r_p1 = 0.0, n_ll = 0.0, n_rr = 0.0;
        const bool more = (i + 1 < i1);
        if (more) {
            n_c_p2 = __ldg(pc + 3 * P);
            n_l_p1 = __ldg(pc + 2 * P - K);
            n_r_p1 = __ldg(pc + 2 * P + K);
            n_ll = __ldg(pc + P - 2 * K);
            n_rr = __ldg(pc + P + 2 * K);
        }
        const double cf = ldg_stream(pco);

        const double lap_p = lap5(c_p1, c_p2, c_0, r_p1, l_p1);   // lap(p+1, q)
        const double lap_r = lap5(r_0, r_p1, r_m1, rr_0, c_0);    // lap(p, q+1)
        const double lap_l = lap5(l_0, l_p1, l_m1, c_0, ll_0);    // lap(p, q-1)

        const double flx_c = limit(lap_p - lap_c, c_p1 - c_0);
        const double fly_c = limit(lap_r - lap_c, r_0 - c_0);
        const double fly_m = limit(lap_c - lap_l, c_0 - l_0);

        // hdiff_numpy.py:27-29
        const double res = c_0 - cf * (((flx_c - flx_m) + fly_c) - fly_m);
        stg_stream(po, res);

        // roll the window one row down
        c_m2 = c_m1; c_m1 = c_0; c_0 = c_p1; c_p1 = c_p2; c_p2 = n_c_p2;
        l_m1 = l_0; l_0 = l_p1; l_p1 = n_l_p1;
        r_m1 = r_0; r_0 = r_p1; r_p1 = n_r_p1;
        ll_0 = n_ll; rr_0 = n_rr;
        lap_m = lap_c; lap_c = lap_p; flx_m = flx_c;
        pc += P; pco += JK; po += JK;
    }
}

}  // namespace

extern "C" int npb_hdiff_f64(int64_t I, int64_t J, int64_t K, const double *in_field,
                             double *out_field, const double *coeff) {
    NPB_REQUIRE_INIT();
    NPB_ARG(I >= 0 && J >= 0 && K >= 0, "npb_hdiff_f64", "negative extent");
    if (I == 0 || J == 0 || K == 0) return 0;
    NPB_ARG(J * K < (1LL << 31) && I < (1LL << 31), "npb_hdiff_f64", "plane too large for 32-bit column index");
    const int JK = (int)(J * K);
    // enough chunks along i to fill the machine a few times over, but long
    // enough marches to amortise the 13-load window prologue
    const int col_blocks = (JK + HD_THREADS - 1) / HD_THREADS;
    int rows = 16;
    while (rows > 4 && (long long)col_blocks * ((I + rows - 1) / rows) < 4LL * npb::st().sm_count) rows >>= 1;
    dim3 grid(col_blocks, (unsigned)((I + rows - 1) / rows));
    hdiff_march_kernel<<<grid, HD_THREADS, 0, npb::st().stream>>>(
        (int)I, JK, (int)K, (long long)(J + 4) * K, in_field, out_field, coeff, rows);
    NPB_CHECK_LAUNCH("hdiff_march_kernel");
    npb::count_launch();
    return 0;
}
