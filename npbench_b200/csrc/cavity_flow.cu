// cavity_flow.cu -- lid-driven cavity micro-app (widening row), sm_100a.
//
// Replaces cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu),
// npbench/benchmarks/cavity_flow/cavity_flow_numpy.py:46-89 (build_up_b :13-23, pressure_poisson :26-43).
// Per time step: one kernel for the source term b, `nit` Jacobi iterations of the pressure Poisson equation
// (each a full NumPy pass + four boundary assignments in the reference), one kernel for the momentum update of
// u and v with their boundary values.  Every kernel writes a COMPLETE array (interior and boundary cells), so the
// reference's `pn = p.copy()`, `un = u.copy()`, `vn = v.copy()` become ping-pong buffers; the boundary
// assignments (:40-43, :80-87; their order matters at the corners) are folded into the producing kernel: each
// boundary cell of p is the value of one specific interior cell (or 0), written by that cell's thread.
// The grids are tiny (61^2 .. 201^2), so the call is a chain of nt*(nit+2) dependent launches: captured once
// per problem and replayed as one CUDA graph.
//
// Arithmetic order as in oracle/stencil_oracle.c: npb_oracle_cavity_flow (NumPy order, Python-float scalar
// subexpressions computed on the host with the same libm calls); -fmad=false.
#include <math.h>

#include "common.cuh"

namespace {

struct CavCoef {
    double s1, c2dx, c2dy;          // 1/dt, 2*dx, 2*dy
    double dx2, dy2, den, coef;     // dx**2, dy**2, 2*(dx**2+dy**2), dx**2*dy**2/den
    double dt, dx, dy, rho, nu;
    double cpx, cpy, cdx, cdy;      // dt/(2*rho*dx), dt/(2*rho*dy), dt/dx**2, dt/dy**2
};

#define CAV_AT(a, i, j) a[(size_t)(i) * nx + (j)]

// build_up_b (:15-23): interior of b
__global__ void cavity_b_kernel(int nx, int ny, CavCoef k, const double *__restrict__ u, const double *__restrict__ v,
                                double *__restrict__ b) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = 1 + blockIdx.x * blockDim.x + threadIdx.x, i = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 2 || j > nx - 2) return;
    const double a1 = (CAV_AT(u, i, j + 1) - CAV_AT(u, i, j - 1)) / k.c2dx;
    const double a2 = (CAV_AT(v, i + 1, j) - CAV_AT(v, i - 1, j)) / k.c2dy;
    const double t1 = k.s1 * (a1 + a2);
    const double t2 = a1 * a1;
    const double t3 = 2.0 * ((((CAV_AT(u, i + 1, j) - CAV_AT(u, i - 1, j)) / k.c2dy) * (CAV_AT(v, i, j + 1) - CAV_AT(v, i, j - 1))) / k.c2dx);
    const double t4 = a2 * a2;
    CAV_AT(b, i, j) = k.rho * (((t1 - t2) - t3) - t4);
}

// one pressure iteration (:31-43): pn -> p, boundary cells included.  After the four assignments
//   p[:, -1] = p[:, -2] ; p[0, :] = p[1, :] ; p[:, 0] = p[:, 1] ; p[-1, :] = 0
// row ny-1 is 0, and every other boundary cell equals the new value of its interior neighbour (the corners
// (0,0) and (0,nx-1) that of (1,1) and (1,nx-2)).
__global__ void cavity_p_kernel(int nx, int ny, CavCoef k, const double *__restrict__ pn, double *__restrict__ p,
                                const double *__restrict__ b) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = 1 + blockIdx.x * blockDim.x + threadIdx.x, i = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 2 || j > nx - 2) return;
    const double val = ((((CAV_AT(pn, i, j + 1) + CAV_AT(pn, i, j - 1)) * k.dy2) + ((CAV_AT(pn, i + 1, j) + CAV_AT(pn, i - 1, j)) * k.dx2)) / k.den) -
                       (k.coef * CAV_AT(b, i, j));
    CAV_AT(p, i, j) = val;
    const bool west = (j == 1), east = (j == nx - 2);
    if (east) CAV_AT(p, i, nx - 1) = val;
    if (west) CAV_AT(p, i, 0) = val;
    if (i == 1) {
        CAV_AT(p, 0, j) = val;
        if (east) CAV_AT(p, 0, nx - 1) = val;
        if (west) CAV_AT(p, 0, 0) = val;
    }
    if (i == ny - 2) {
        CAV_AT(p, ny - 1, j) = 0.0;
        if (east) CAV_AT(p, ny - 1, nx - 1) = 0.0;
        if (west) CAV_AT(p, ny - 1, 0) = 0.0;
    }
}

// momentum update (:56-87): (un, vn, p) -> (u, v), boundary cells included (lid row ny-1: u = 1)
__global__ void cavity_uv_kernel(int nx, int ny, CavCoef k, const double *__restrict__ un, const double *__restrict__ vn,
                                 const double *__restrict__ p, double *__restrict__ u, double *__restrict__ v) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 1 || j > nx - 1) return;
    if (i == 0 || j == 0 || i == ny - 1 || j == nx - 1) {
        CAV_AT(u, i, j) = (i == ny - 1) ? 1.0 : 0.0;          // u[-1, :] = 1 is assigned last (:83)
        CAV_AT(v, i, j) = 0.0;
        return;
    }
    const double uc = CAV_AT(un, i, j), vc = CAV_AT(vn, i, j);
    const double adv_u = (uc * k.dt) / k.dx, adv_v = (vc * k.dt) / k.dy;
    const double lap_u = (k.cdx * ((CAV_AT(un, i, j + 1) - 2.0 * uc) + CAV_AT(un, i, j - 1))) +
                         (k.cdy * ((CAV_AT(un, i + 1, j) - 2.0 * uc) + CAV_AT(un, i - 1, j)));
    CAV_AT(u, i, j) = (((uc - adv_u * (uc - CAV_AT(un, i, j - 1))) - adv_v * (uc - CAV_AT(un, i - 1, j))) -
                       k.cpx * (CAV_AT(p, i, j + 1) - CAV_AT(p, i, j - 1))) + k.nu * lap_u;
    const double lap_v = (k.cdx * ((CAV_AT(vn, i, j + 1) - 2.0 * vc) + CAV_AT(vn, i, j - 1))) +
                         (k.cdy * ((CAV_AT(vn, i + 1, j) - 2.0 * vc) + CAV_AT(vn, i - 1, j)));
    CAV_AT(v, i, j) = (((vc - adv_u * (vc - CAV_AT(vn, i, j - 1))) - adv_v * (vc - CAV_AT(vn, i - 1, j))) -
                       k.cpy * (CAV_AT(p, i + 1, j) - CAV_AT(p, i - 1, j))) + k.nu * lap_v;
}

long long bits(double x) { long long r; memcpy(&r, &x, sizeof(r)); return r; }

}  // namespace

extern "C" int npb_cavity_flow_f64(int64_t nx, int64_t ny, int64_t nt, int64_t nit, double *u, double *v, double dt,
                                   double dx, double dy, double *p, double rho, double nu) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx >= 3 && ny >= 3 && nx < (1 << 15) && ny < (1 << 15), "npb_cavity_flow_f64", "nx and ny must be in [3, 32768)");
    NPB_ARG(nt >= 0 && nit >= 0 && nt < (1LL << 31) && nit < (1LL << 31), "npb_cavity_flow_f64", "negative step count");
    if (nt == 0) return 0;
    // Python-float subexpressions of the reference, same operations and the same libm pow (not folded to x*x)
    double (*volatile pw)(double, double) = pow;
    CavCoef k;
    k.s1 = 1.0 / dt; k.c2dx = 2.0 * dx; k.c2dy = 2.0 * dy;
    k.dx2 = pw(dx, 2.0); k.dy2 = pw(dy, 2.0);
    k.den = 2.0 * (k.dx2 + k.dy2); k.coef = (k.dx2 * k.dy2) / k.den;
    k.dt = dt; k.dx = dx; k.dy = dy; k.rho = rho; k.nu = nu;
    k.cpx = dt / ((2.0 * rho) * dx); k.cpy = dt / ((2.0 * rho) * dy);
    k.cdx = dt / k.dx2; k.cdy = dt / k.dy2;
    const size_t cells = (size_t)nx * (size_t)ny, bytes = cells * sizeof(double);
    double *ws = (double *)npb::workspace(6, 4 * bytes);
    NPB_ARG(ws != nullptr, "npb_cavity_flow_f64", "out of device memory for the work arrays");
    double *pbuf[2] = {p, ws}, *ubuf[2] = {u, ws + cells}, *vbuf[2] = {v, ws + 2 * cells}, *b = ws + 3 * cells;
    npb::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.kind = 8;
    key.dims[0] = (nx << 32) | ny; key.dims[1] = (nt << 32) | nit; key.dims[2] = bits(dt); key.dims[3] = bits(dx);
    key.ptrs[0] = u; key.ptrs[1] = v; key.ptrs[2] = p; key.ptrs[3] = ws;
    key.ptrs[4] = (const void *)bits(dy); key.ptrs[5] = (const void *)bits(rho); key.ptrs[6] = (const void *)bits(nu);
    const bool use_graph = nt * (nit + 2) >= 8;
    if (use_graph && npb::graph_replay(key)) return 0;
    const bool capturing = use_graph && npb::graph_begin();
    const dim3 blk(32, 8), grid_in((unsigned)((nx - 2 + 31) / 32), (unsigned)((ny - 2 + 7) / 8)),
        grid_all((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8));
    cudaStream_t st = npb::st().stream;
    int pc = 0, uc = 0, rc = 0;
    for (int64_t n = 0; n < nt && !rc; ++n) {
        pdl_launch(cavity_b_kernel, grid_in, blk, 0, st, (int)nx, (int)ny, k, (const double *)ubuf[uc], (const double *)vbuf[uc], b);
        for (int64_t q = 0; q < nit; ++q) {
            pdl_launch(cavity_p_kernel, grid_in, blk, 0, st, (int)nx, (int)ny, k, (const double *)pbuf[pc], pbuf[pc ^ 1], (const double *)b);
            pc ^= 1;
        }
        pdl_launch(cavity_uv_kernel, grid_all, blk, 0, st, (int)nx, (int)ny, k, (const double *)ubuf[uc], (const double *)vbuf[uc], (const double *)pbuf[pc], ubuf[uc ^ 1], vbuf[uc ^ 1]);
        uc ^= 1;
        if (cudaGetLastError() != cudaSuccess) rc = npb::fail("npb_cavity_flow_f64", "kernel launch failed");
        npb::count_launch((int)(nit + 2));
    }
    if (!rc && pc) rc = cudaMemcpyAsync(p, pbuf[1], bytes, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : npb::fail("npb_cavity_flow_f64", "copy failed");
    if (!rc && uc) {
        if (cudaMemcpyAsync(u, ubuf[1], bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(v, vbuf[1], bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = npb::fail("npb_cavity_flow_f64", "copy failed");
    }
    if (capturing) {
        const int rc2 = npb::graph_end_and_launch(key, rc);
        if (!rc) rc = rc2;
    }
    return rc;
}

extern "C" int npb_cavity_flow_f64_host(int64_t nx, int64_t ny, int64_t nt, int64_t nit, double *u, double *v, double dt,
                                        double dx, double dy, double *p, double rho, double nu) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx >= 3 && ny >= 3, "npb_cavity_flow_f64_host", "nx and ny must be >= 3");
    const size_t bytes = (size_t)nx * (size_t)ny * sizeof(double);
    void *d[3] = {nullptr, nullptr, nullptr};
    double *h[3] = {u, v, p};
    int rc = 0;
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_malloc(bytes, &d[a]);
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_h2d(d[a], h[a], bytes);
    if (!rc) rc = npb_cavity_flow_f64(nx, ny, nt, nit, (double *)d[0], (double *)d[1], dt, dx, dy, (double *)d[2], rho, nu);
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_d2h(h[a], d[a], bytes);
    if (!rc) rc = npb_sync();
    for (int a = 0; a < 3; ++a) if (d[a]) npb_free(d[a]);
    return rc;
}
