// init_fields.cu -- device-side closed-form initialisers (sm_100a).
//
// NPBench's `initialize` functions for the three PolyBench stencils are closed
// forms over the index (np.fromfunction):
//   jacobi_2d.py:6-10   A = i*(j+2)/N,          B = i*(j+3)/N
//   heat_3d.py:6-11     A = B = (i+j+(N-k))*10/N
//   fdtd_2d.py:6-15     ex = i*(j+1)/NX, ey = i*(j+2)/NY, hz = i*(j+3)/NX, _fict_[t] = t
// They are evaluated here on the device for row slabs [row0, row0+nrows) of
// the global grid, so that the scaled multi-GPU grids (tens of GB per GPU)
// never exist on the host.  All index products are exact in binary64 (< 2^53);
// one IEEE division per element => bit-identical to np.fromfunction.
#include "common.cuh"

namespace {

__global__ void init_jacobi2d_kernel(double n, long long row0, long long nrows, long long ncols,
                                     double *__restrict__ A, double *__restrict__ B) {
    const long long total = nrows * ncols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / ncols, j = idx - r * ncols;
        const double i = (double)(row0 + r);
        A[idx] = (i * ((double)j + 2.0)) / n;
        B[idx] = (i * ((double)j + 3.0)) / n;
    }
}

__global__ void init_heat3d_kernel(long long n, long long row0, long long nrows,
                                   double *__restrict__ A, double *__restrict__ B) {
    const long long total = nrows * n * n;
    const double dn = (double)n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / (n * n), rem = idx - r * n * n;
        const long long j = rem / n, k = rem - j * n;
        const double i = (double)(row0 + r);
        const double v = (((i + (double)j) + (dn - (double)k)) * 10.0) / dn;
        A[idx] = v;
        B[idx] = v;
    }
}

__global__ void init_fdtd2d_kernel(long long tmax, double nx, double nyd, long long ny,
                                   long long row0, long long nrows, double *__restrict__ ex,
                                   double *__restrict__ ey, double *__restrict__ hz,
                                   double *__restrict__ fict) {
    const long long total = nrows * ny;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long idx = tid; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / ny, j = idx - r * ny;
        const double i = (double)(row0 + r);
        ex[idx] = (i * ((double)j + 1.0)) / nx;
        ey[idx] = (i * ((double)j + 2.0)) / nyd;
        hz[idx] = (i * ((double)j + 3.0)) / nx;
    }
    if (fict)
        for (long long t = tid; t < tmax; t += (long long)gridDim.x * blockDim.x) fict[t] = (double)t;
}

unsigned grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 32LL * npb::st().sm_count;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace

extern "C" int npb_init_jacobi2d_f64(int64_t n_global, int64_t row0, int64_t nrows, int64_t ncols,
                                     double *A, double *B) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n_global > 0 && nrows >= 0 && ncols >= 0, "npb_init_jacobi2d_f64", "bad extent");
    if (nrows * ncols == 0) return 0;
    init_jacobi2d_kernel<<<grid_for(nrows * ncols), 256, 0, npb::st().stream>>>(
        (double)n_global, row0, nrows, ncols, A, B);
    NPB_CHECK_LAUNCH("init_jacobi2d_kernel");
    npb::count_launch();
    return 0;
}

extern "C" int npb_init_heat3d_f64(int64_t n_global, int64_t row0, int64_t nrows, double *A,
                                   double *B) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n_global > 0 && nrows >= 0, "npb_init_heat3d_f64", "bad extent");
    if (nrows == 0) return 0;
    init_heat3d_kernel<<<grid_for(nrows * n_global * n_global), 256, 0, npb::st().stream>>>(
        n_global, row0, nrows, A, B);
    NPB_CHECK_LAUNCH("init_heat3d_kernel");
    npb::count_launch();
    return 0;
}

extern "C" int npb_init_fdtd2d_f64(int64_t tmax, int64_t nx_global, int64_t ny, int64_t row0,
                                   int64_t nrows, double *ex, double *ey, double *hz, double *fict) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx_global > 0 && ny > 0 && nrows >= 0, "npb_init_fdtd2d_f64", "bad extent");
    init_fdtd2d_kernel<<<grid_for(nrows * ny + 1), 256, 0, npb::st().stream>>>(
        tmax, (double)nx_global, (double)ny, ny, row0, nrows, ex, ey, hz, fict);
    NPB_CHECK_LAUNCH("init_fdtd2d_kernel");
    npb::count_launch();
    return 0;
}
