// channel_flow.cu -- periodic channel flow micro-app (widening row), sm_100a.
//
// Replaces channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F),
// npbench/benchmarks/channel_flow/channel_flow_numpy.py:74-170 (build_up_b :13-40,
// pressure_poisson_periodic :43-71).  Same scheme as cavity_flow.cu with
//   * periodic x boundaries: the reference's "Periodic BC" statements are the interior formulas with the
//     column indices wrapped, so one kernel covers all columns;
//   * walls at rows 0 / ny-1 (u = v = 0, dp/dy = 0), folded into the producing kernels;
//   * the body force F*dt on u;
//   * the convergence loop  udiff = (np.sum(u) - np.sum(un)) / np.sum(u) ; while udiff > .001.
// np.sum is NumPy's pairwise summation over the flattened array (blocks of <= 128 elements with 8 strided
// accumulators, recursive halving on multiples of 8).  It is reproduced bit for bit -- one thread per leaf block,
// one thread walking the (host-built) combination tree -- because a differently rounded sum could move the step
// at which the loop stops.  The host reads one double per time step (the new sum; sum(un) is last step's sum(u))
// and evaluates udiff with the reference's two operations.  Returns the step count like the reference.
//
// Arithmetic order as in oracle/stencil_oracle.c: npb_oracle_channel_flow; -fmad=false.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

struct ChCoef {
    double s1, c2dx, c2dy, dx2, dy2, den, coef, dt, dx, dy, rho, nu, cpx, cpy, cdx, cdy, fdt;
};

#define CH_AT(a, i, j) a[(size_t)(i) * nx + (j)]

__global__ void channel_b_kernel(int nx, int ny, ChCoef k, const double *__restrict__ u, const double *__restrict__ v,
                                 double *__restrict__ b) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 2 || j > nx - 1) return;
    const int je = (j + 1 == nx) ? 0 : j + 1, jw = (j == 0) ? nx - 1 : j - 1;
    const double a1 = (CH_AT(u, i, je) - CH_AT(u, i, jw)) / k.c2dx;
    const double a2 = (CH_AT(v, i + 1, j) - CH_AT(v, i - 1, j)) / k.c2dy;
    const double t1 = k.s1 * (a1 + a2);
    const double t2 = a1 * a1;
    const double t3 = 2.0 * ((((CH_AT(u, i + 1, j) - CH_AT(u, i - 1, j)) / k.c2dy) * (CH_AT(v, i, je) - CH_AT(v, i, jw))) / k.c2dx);
    const double t4 = a2 * a2;
    CH_AT(b, i, j) = k.rho * (((t1 - t2) - t3) - t4);
}

// one pressure iteration: pn -> p for rows 1..ny-2 (all columns, wrapped), then p[-1,:] = p[-2,:], p[0,:] = p[1,:]
__global__ void channel_p_kernel(int nx, int ny, ChCoef k, const double *__restrict__ pn, double *__restrict__ p,
                                 const double *__restrict__ b) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 2 || j > nx - 1) return;
    const int je = (j + 1 == nx) ? 0 : j + 1, jw = (j == 0) ? nx - 1 : j - 1;
    const double val = ((((CH_AT(pn, i, je) + CH_AT(pn, i, jw)) * k.dy2) + ((CH_AT(pn, i + 1, j) + CH_AT(pn, i - 1, j)) * k.dx2)) / k.den) -
                       (k.coef * CH_AT(b, i, j));
    CH_AT(p, i, j) = val;
    if (i == ny - 2) CH_AT(p, ny - 1, j) = val;
    if (i == 1) CH_AT(p, 0, j) = val;
}

__global__ void channel_uv_kernel(int nx, int ny, ChCoef k, const double *__restrict__ un, const double *__restrict__ vn,
                                  const double *__restrict__ p, double *__restrict__ u, double *__restrict__ v) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > ny - 1 || j > nx - 1) return;
    if (i == 0 || i == ny - 1) { CH_AT(u, i, j) = 0.0; CH_AT(v, i, j) = 0.0; return; }     // walls (:158-161)
    const int je = (j + 1 == nx) ? 0 : j + 1, jw = (j == 0) ? nx - 1 : j - 1;
    const double uc = CH_AT(un, i, j), vc = CH_AT(vn, i, j);
    const double adv_u = (uc * k.dt) / k.dx, adv_v = (vc * k.dt) / k.dy;
    const double lap_u = (k.cdx * ((CH_AT(un, i, je) - 2.0 * uc) + CH_AT(un, i, jw))) +
                         (k.cdy * ((CH_AT(un, i + 1, j) - 2.0 * uc) + CH_AT(un, i - 1, j)));
    CH_AT(u, i, j) = ((((uc - adv_u * (uc - CH_AT(un, i, jw))) - adv_v * (uc - CH_AT(un, i - 1, j))) -
                       k.cpx * (CH_AT(p, i, je) - CH_AT(p, i, jw))) + k.nu * lap_u) + k.fdt;
    const double lap_v = (k.cdx * ((CH_AT(vn, i, je) - 2.0 * vc) + CH_AT(vn, i, jw))) +
                         (k.cdy * ((CH_AT(vn, i + 1, j) - 2.0 * vc) + CH_AT(vn, i - 1, j)));
    CH_AT(v, i, j) = (((vc - adv_u * (vc - CH_AT(vn, i, jw))) - adv_v * (vc - CH_AT(vn, i - 1, j))) -
                      k.cpy * (CH_AT(p, i + 1, j) - CH_AT(p, i - 1, j))) + k.nu * lap_v;
}

// ---- np.sum: leaf blocks (8 <= n <= 128) and the combination tree ---------------------------------------------
__global__ void npsum_leaf_kernel(const double *__restrict__ a, const long long *__restrict__ leaf_start,
                                  const int *__restrict__ leaf_n, int nleaf, double *__restrict__ leaf_sum) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaf) return;
    const double *x = a + leaf_start[l];
    const int n = leaf_n[l];
    double res;
    if (n < 8) {
        res = 0.0;
        for (int i = 0; i < n; ++i) res += x[i];
    } else {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = x[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += x[i + j];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += x[i];
    }
    leaf_sum[l] = res;
}

// prog: post-order walk of pairwise_sum's recursion; entry >= 0: push leaf_sum[entry]; -1: pop b, pop a, push a + b
__global__ void npsum_tree_kernel(const int *__restrict__ prog, int nprog, const double *__restrict__ leaf_sum,
                                  double *__restrict__ out) {
    pdl_wait();                                   // launched with pdl_launch (common.cuh)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double stack[48];
    int sp = 0;
    for (int q = 0; q < nprog; ++q) {
        const int e = prog[q];
        if (e >= 0) stack[sp++] = leaf_sum[e];
        else { const double b = stack[--sp], a = stack[--sp]; stack[sp++] = a + b; }
    }
    *out = stack[0];
}

long long bits(double x) { long long r; memcpy(&r, &x, sizeof(r)); return r; }

void build_tree(long long start, long long n, std::vector<long long> &ls, std::vector<int> &ln, std::vector<int> &prog) {
    if (n <= 128) {
        prog.push_back((int)ls.size());
        ls.push_back(start); ln.push_back((int)n);
        return;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    build_tree(start, n2, ls, ln, prog);
    build_tree(start + n2, n - n2, ls, ln, prog);
    prog.push_back(-1);
}

}  // namespace

extern "C" int npb_channel_flow_f64(int64_t nit, int64_t nx, int64_t ny, double *u, double *v, double dt, double dx,
                                    double dy, double *p, double rho, double nu, double F, int64_t *stepcount) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx >= 3 && ny >= 3 && nx < (1 << 15) && ny < (1 << 15), "npb_channel_flow_f64", "nx and ny must be in [3, 32768)");
    NPB_ARG(nit >= 0 && nit < (1LL << 31), "npb_channel_flow_f64", "negative iteration count");
    NPB_ARG(stepcount != nullptr, "npb_channel_flow_f64", "null stepcount pointer");
    double (*volatile pw)(double, double) = pow;             // libm pow like Python's float ** (not folded to x*x)
    ChCoef k;
    k.s1 = 1.0 / dt; k.c2dx = 2.0 * dx; k.c2dy = 2.0 * dy;
    k.dx2 = pw(dx, 2.0); k.dy2 = pw(dy, 2.0);
    k.den = 2.0 * (k.dx2 + k.dy2); k.coef = (k.dx2 * k.dy2) / k.den;
    k.dt = dt; k.dx = dx; k.dy = dy; k.rho = rho; k.nu = nu;
    k.cpx = dt / ((2.0 * rho) * dx); k.cpy = dt / ((2.0 * rho) * dy);
    k.cdx = dt / k.dx2; k.cdy = dt / k.dy2; k.fdt = F * dt;
    const size_t cells = (size_t)nx * (size_t)ny, bytes = cells * sizeof(double);
    // summation tree of np.sum over `cells` elements
    std::vector<long long> ls; std::vector<int> ln, prog;
    build_tree(0, (long long)cells, ls, ln, prog);
    const size_t nleaf = ls.size(), nprog = prog.size();
    const size_t off_ls = 4 * bytes, off_ln = off_ls + nleaf * 8, off_prog = off_ln + ((nleaf * 4 + 7) & ~(size_t)7),
                 off_sum = off_prog + ((nprog * 4 + 7) & ~(size_t)7), off_out = off_sum + nleaf * 8, total = off_out + 16;
    char *ws = (char *)npb::workspace(2, total);
    NPB_ARG(ws != nullptr, "npb_channel_flow_f64", "out of device memory for the work arrays");
    cudaStream_t st = npb::st().stream;
    NPB_CUDA(cudaMemcpyAsync(ws + off_ls, ls.data(), nleaf * 8, cudaMemcpyHostToDevice, st));
    NPB_CUDA(cudaMemcpyAsync(ws + off_ln, ln.data(), nleaf * 4, cudaMemcpyHostToDevice, st));
    NPB_CUDA(cudaMemcpyAsync(ws + off_prog, prog.data(), nprog * 4, cudaMemcpyHostToDevice, st));
    NPB_CUDA(cudaStreamSynchronize(st));                     // the vectors go out of scope with this call
    double *w0 = (double *)ws;
    double *pbuf[2] = {p, w0}, *ubuf[2] = {u, w0 + cells}, *vbuf[2] = {v, w0 + 2 * cells}, *b = w0 + 3 * cells;
    const long long *d_ls = (const long long *)(ws + off_ls);
    const int *d_ln = (const int *)(ws + off_ln), *d_prog = (const int *)(ws + off_prog);
    double *d_leaf = (double *)(ws + off_sum), *d_out = (double *)(ws + off_out);
    static double *h_sum = nullptr;                          // pinned: the per-step D2H copy is part of a captured graph
    if (!h_sum) NPB_CUDA(cudaHostAlloc((void **)&h_sum, 64, cudaHostAllocDefault));
    auto enqueue_sum = [&](const double *a) -> int {        // np.sum(a) -> *h_sum (valid after a stream sync)
        pdl_launch(npsum_leaf_kernel, dim3((unsigned)((nleaf + 127) / 128)), dim3(128), 0, st, a, d_ls, d_ln, (int)nleaf, d_leaf);
        pdl_launch(npsum_tree_kernel, dim3(1), dim3(32), 0, st, d_prog, (int)nprog, (const double *)d_leaf, d_out);
        npb::count_launch(2);
        return cudaMemcpyAsync(h_sum, d_out, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess;
    };
    NPB_CUDA(cudaMemsetAsync(b, 0, bytes, st));              // b = zeros_like(u): rows 0 and ny-1 are never written
    NPB_ARG(enqueue_sum(u) == 0, "npb_channel_flow_f64", "device sum failed");
    NPB_CUDA(cudaStreamSynchronize(st));
    double sum_prev = *h_sum, sum_new = 0.0;
    const dim3 blk(32, 8), grid_rows((unsigned)((nx + 31) / 32), (unsigned)((ny - 2 + 7) / 8)),
        grid_all((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8));
    int pc = 0, uc = 0;
    double udiff = 1.0;
    int64_t steps = 0;
    // One time step = nit + 4 tiny dependent launches + the copy of the sum: captured once per buffer parity
    // (which of the ping-pong buffers is current) and replayed as a graph.
    npb::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.kind = 9;
    key.dims[1] = bits(dt); key.dims[2] = bits(dx); key.dims[3] = bits(dy);
    key.ptrs[0] = u; key.ptrs[1] = v; key.ptrs[2] = p; key.ptrs[3] = ws;
    key.ptrs[4] = (const void *)bits(rho); key.ptrs[5] = (const void *)bits(nu); key.ptrs[6] = (const void *)bits(F);
    while (udiff > .001) {                                   // :78
        key.dims[0] = ((long long)nx << 48) | ((long long)ny << 33) | ((long long)nit << 2) | (long long)(pc | (uc << 1));
        const int pc_in = pc, uc_in = uc;
        pc ^= (int)(nit & 1); uc ^= 1;
        if (!npb::graph_replay(key)) {
            const bool capturing = npb::graph_begin();
            int q_pc = pc_in;
            pdl_launch(channel_b_kernel, grid_rows, blk, 0, st, (int)nx, (int)ny, k, (const double *)ubuf[uc_in], (const double *)vbuf[uc_in], b);
            for (int64_t q = 0; q < nit; ++q) {
                pdl_launch(channel_p_kernel, grid_rows, blk, 0, st, (int)nx, (int)ny, k, (const double *)pbuf[q_pc], pbuf[q_pc ^ 1], (const double *)b);
                q_pc ^= 1;
            }
            pdl_launch(channel_uv_kernel, grid_all, blk, 0, st, (int)nx, (int)ny, k, (const double *)ubuf[uc_in], (const double *)vbuf[uc_in],
                       (const double *)pbuf[q_pc], ubuf[uc_in ^ 1], vbuf[uc_in ^ 1]);
            npb::count_launch((int)(nit + 2));
            int rc = enqueue_sum(ubuf[uc_in ^ 1]);
            if (cudaGetLastError() != cudaSuccess) rc = 1;
            if (capturing) {
                const int rc2 = npb::graph_end_and_launch(key, rc);
                if (!rc) rc = rc2;
            }
            if (rc) return npb::fail("npb_channel_flow_f64", "kernel launch failed");
        }
        NPB_CUDA(cudaStreamSynchronize(st));
        sum_new = *h_sum;
        udiff = (sum_new - sum_prev) / sum_new;              // :167  (np.sum(un) is the previous step's np.sum(u))
        sum_prev = sum_new;
        ++steps;
        NPB_ARG(steps < (1LL << 40), "npb_channel_flow_f64", "no convergence");
    }
    if (pc) NPB_CUDA(cudaMemcpyAsync(p, pbuf[1], bytes, cudaMemcpyDeviceToDevice, st));
    if (uc) {
        NPB_CUDA(cudaMemcpyAsync(u, ubuf[1], bytes, cudaMemcpyDeviceToDevice, st));
        NPB_CUDA(cudaMemcpyAsync(v, vbuf[1], bytes, cudaMemcpyDeviceToDevice, st));
    }
    *stepcount = steps;
    return 0;
}

extern "C" int npb_channel_flow_f64_host(int64_t nit, int64_t nx, int64_t ny, double *u, double *v, double dt, double dx,
                                         double dy, double *p, double rho, double nu, double F, int64_t *stepcount) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx >= 3 && ny >= 3, "npb_channel_flow_f64_host", "nx and ny must be >= 3");
    const size_t bytes = (size_t)nx * (size_t)ny * sizeof(double);
    void *d[3] = {nullptr, nullptr, nullptr};
    double *h[3] = {u, v, p};
    int rc = 0;
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_malloc(bytes, &d[a]);
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_h2d(d[a], h[a], bytes);
    if (!rc) rc = npb_channel_flow_f64(nit, nx, ny, (double *)d[0], (double *)d[1], dt, dx, dy, (double *)d[2], rho, nu, F, stepcount);
    for (int a = 0; a < 3 && !rc; ++a) rc = npb_d2h(h[a], d[a], bytes);
    if (!rc) rc = npb_sync();
    for (int a = 0; a < 3; ++a) if (d[a]) npb_free(d[a]);
    return rc;
}
