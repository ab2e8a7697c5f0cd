// seidel2d.cu -- Polybench 2-D Gauss-Seidel (widening row), sm_100a.
//
// Replaces kernel(TSTEPS, N, A), npbench/benchmarks/polybench/seidel_2d/seidel_2d_numpy.py:4-13:
//   for t in range(TSTEPS-1): for i in 1..N-2:
//       A[i,1:-1] += A[i-1,:-2] + A[i-1,1:-1] + A[i-1,2:] + A[i,2:] + A[i+1,:-2] + A[i+1,1:-1] + A[i+1,2:]
//       for j in 1..N-2: A[i,j] += A[i,j-1] ; A[i,j] /= 9.0
// i.e. with "new" = already updated in sweep t and "old" = value of sweep t-1:
//   S = ((((((new[i-1,j-1] + new[i-1,j]) + new[i-1,j+1]) + old[i,j+1]) + old[i+1,j-1]) + old[i+1,j]) + old[i+1,j+1])
//   new[i,j] = ((old[i,j] + S) + new[i,j-1]) / 9.0
//
// Not a Jacobi-type kernel: every cell depends on its west and north neighbours of the same sweep,
// so the parallelism is a wavefront.  Cell (t,i,j) only depends on cells with a smaller value of
//   h = 4t + 2i + j        [(t,i-1,j+1): h-1, (t,i,j-1): h-1, (t-1,i+1,j+1): h-1, (t-1,i,j+1): h-3]
// and no cell of hyperplane h reads a value that another cell of h writes, so all sweeps can be
// in flight at once: ONE thread-block cluster (up to 8 CTAs x 1024 threads) walks the hyperplanes, in
// place, with a hardware cluster barrier (barrier.cluster, release/acquire) between them -- a grid-wide
// software barrier costs ~2 us per hyperplane, the cluster barrier a few hundred cycles.  A thread owns
// the row tasks (t,i) = p, p + #threads, ... and advances each along j.  All accesses go to L2
// (ld.global.cg / st.global.cg): values travel between SMs every step.
// The critical path is 4(TSTEPS-2) + 3(N-2) barriers; the cluster is only as large as the tasks need.
//
// Arithmetic order as in oracle/stencil_oracle.c: npb_oracle_seidel2d; IEEE division; -fmad=false.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int S2_THREADS = 512;
constexpr int S2_SLOTS = 12;         // row tasks per thread kept decoded in registers (DSMEM kernel)

__global__ void __launch_bounds__(S2_THREADS, 1)
seidel2d_wavefront_kernel(int tsteps, int n, double *A) {
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int ni = n - 2;                                       // interior rows / columns
    const long long ntask = (long long)(tsteps - 1) * ni;       // row tasks (t, i)
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int h_lo = 3, h_hi = 4 * (tsteps - 2) + 3 * ni;       // first / last hyperplane
    for (int h = h_lo; h <= h_hi; ++h) {
        for (long long p = gtid; p < ntask; p += nthr) {
            const int t = (int)(p / ni), i = (int)(p - (long long)t * ni) + 1;
            const int j = h - 4 * t - 2 * i;
            if (j < 1 || j > ni) continue;
            const double *r = A + (long long)i * n + j;
            const double *u = r - n, *d = r + n;
            const double s = ((((((__ldcg(u - 1) + __ldcg(u)) + __ldcg(u + 1)) + __ldcg(r + 1)) + __ldcg(d - 1)) +
                               __ldcg(d)) + __ldcg(d + 1));
            const double v = (__ldcg(r) + s) + __ldcg(r - 1);
            __stcg(A + (long long)i * n + j, v / 9.0);
        }
        cluster.sync();
    }
}


// ---------------------------------------------------------------------------
// Grids that fit into the shared memory of one cluster (N <= ~440 with 8 CTAs; all NPBench presets):
// the grid lives in DISTRIBUTED SHARED MEMORY for the whole call.  CTA c of the cluster owns a block of
// R consecutive rows in its shared memory and runs the row tasks of those rows, so every write is local
// and only the rows next to a block edge are read from the neighbour CTA through DSMEM
// (cluster.map_shared_rank).  One hardware cluster barrier per hyperplane; no global-memory traffic
// between the initial load and the final store.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(S2_THREADS, 1)
seidel2d_dsmem_kernel(int tsteps, int n, int R, double *A) {
    extern __shared__ double rows[];                            // [R][n]: rows c*R .. c*R+R-1 of the grid
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int c = (int)cluster.block_rank(), nc = (int)cluster.num_blocks();
    const int row0 = c * R, nrow = max(0, min(R, n - row0));    // rows held here
    for (long long w = threadIdx.x; w < (long long)nrow * n; w += S2_THREADS) rows[w] = __ldg(A + (long long)row0 * n + w);
    const double *up_blk = c > 0 ? cluster.map_shared_rank(rows, c - 1) : rows;          // rows of the CTA above / below
    const double *dn_blk = c < nc - 1 ? cluster.map_shared_rank(rows, c + 1) : rows;
    // interior rows owned: global rows max(row0,1) .. min(row0+nrow-1, n-2)
    const int i_lo = max(row0, 1), i_hi = min(row0 + nrow - 1, n - 2);
    const int nown = max(0, i_hi - i_lo + 1);
    const int ni = n - 2;
    const long long ntask = (long long)(tsteps - 1) * nown;
    const int h_lo = 3, h_hi = 4 * (tsteps - 2) + 3 * ni;
    // the first S2_SLOTS tasks of a thread are decoded once: hyperplane offset 4t + 2i and local row
    int t_start[S2_SLOTS], t_li[S2_SLOTS];
#pragma unroll
    for (int k = 0; k < S2_SLOTS; ++k) {
        const long long p = threadIdx.x + (long long)k * S2_THREADS;
        t_start[k] = 1 << 30; t_li[k] = 0;                      // never active
        if (p < ntask) {
            const int t = (int)(p / nown), i = i_lo + (int)(p - (long long)t * nown);
            t_start[k] = 4 * t + 2 * i; t_li[k] = i - row0;
        }
    }
    auto cell = [&](int li, int j) {
        double *r = rows + (long long)li * n + j;
        const double *u = li > 0 ? r - n : up_blk + (long long)(R - 1) * n + j;       // row i-1
        const double *d = li < nrow - 1 ? r + n : dn_blk + j;                          // row i+1
        const double s = ((((((u[-1] + u[0]) + u[1]) + r[1]) + d[-1]) + d[0]) + d[1]);
        const double v = (r[0] + s) + r[-1];
        r[0] = v / 9.0;
    };
    cluster.sync();                                             // every block is loaded
    for (int h = h_lo; h <= h_hi; ++h) {
#pragma unroll
        for (int k = 0; k < S2_SLOTS; ++k) {
            const int j = h - t_start[k];
            if (j >= 1 && j <= ni) cell(t_li[k], j);
        }
        for (long long p = threadIdx.x + (long long)S2_SLOTS * S2_THREADS; p < ntask; p += S2_THREADS) {
            const int t = (int)(p / nown), i = i_lo + (int)(p - (long long)t * nown);
            const int j = h - 4 * t - 2 * i;
            if (j >= 1 && j <= ni) cell(i - row0, j);
        }
        cluster.sync();
    }
    for (long long w = threadIdx.x; w < (long long)nrow * n; w += S2_THREADS) {
        const long long g = (long long)row0 * n + w;
        const int gi = (int)(g / n), gj = (int)(g - (long long)gi * n);
        if (gi >= 1 && gi <= n - 2 && gj >= 1 && gj <= n - 2) A[g] = rows[w];
    }
}

// 1 launched, 0 not applicable
int launch_dsmem(int csize, int64_t tsteps, int64_t n, double *A) {
    const int R = (int)((n + csize - 1) / csize);
    const size_t smem = (size_t)R * n * sizeof(double);
    if (smem + 1024 > npb::st().smem_optin || n < 3 * csize) return 0;
    static size_t configured = 0;
    static bool nonportable = false;
    if (csize > 8 && !nonportable) {
        if (cudaFuncSetAttribute(seidel2d_dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        nonportable = true;
    }
    if (smem > configured) {
        if (cudaFuncSetAttribute(seidel2d_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        configured = smem;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(csize); cfg.blockDim = dim3(S2_THREADS); cfg.dynamicSmemBytes = smem;
    cfg.stream = npb::st().stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, seidel2d_dsmem_kernel, (int)tsteps, (int)n, R, A) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return 1;
}

// 1 launched, 0 not applicable.  16 CTAs per cluster (the B200 maximum, non-portable) when the work is large
// enough to be compute bound on 8 SMs; 8 otherwise.
int try_dsmem(int64_t tsteps, int64_t n, double *A) {
    static int big = -1;
    if (big < 0) { const char *e = getenv("NPB_SEIDEL_CLUSTER"); big = e ? atoi(e) : 16; }
    if (big == 16 && (tsteps - 1) * (n - 2) >= 8LL * S2_THREADS && launch_dsmem(16, tsteps, n, A) == 1) return 1;
    return launch_dsmem(8, tsteps, n, A);
}

int g_seidel_mode = 0;     // 0 dispatch (DSMEM kernel when the grid fits in one cluster), 1 L2 wavefront kernel
int g_seidel_last = 0;     // 1 DSMEM kernel, 2 L2 wavefront kernel

}  // namespace

extern "C" int npb_seidel2d_set_mode(int mode) { g_seidel_mode = mode; return 0; }
extern "C" int npb_seidel2d_last_path(void) { return g_seidel_last; }

extern "C" int npb_seidel2d_f64(int64_t tsteps, int64_t n, double *A) {
    NPB_REQUIRE_INIT();
    NPB_ARG(n >= 0, "npb_seidel2d_f64", "negative extent");
    NPB_ARG(n <= 46000 && tsteps < (1 << 24), "npb_seidel2d_f64", "problem too large for 32-bit hyperplane indices");
    if (tsteps <= 1 || n < 3) return 0;          // range(0, TSTEPS-1) empty / no interior
    if (g_seidel_mode == 0 && try_dsmem(tsteps, n, A) == 1) { g_seidel_last = 1; npb::count_launch(); return 0; }
    g_seidel_last = 2;
    const long long ntask = (tsteps - 1) * (n - 2);
    int csize = 1;                                // CTAs in the cluster: 1, 2, 4 or 8 (portable maximum)
    while (csize < 8 && (long long)csize * S2_THREADS < ntask) csize *= 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)csize); cfg.blockDim = dim3(S2_THREADS); cfg.dynamicSmemBytes = 0;
    cfg.stream = npb::st().stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    NPB_CUDA(cudaLaunchKernelEx(&cfg, seidel2d_wavefront_kernel, (int)tsteps, (int)n, A));
    npb::count_launch();
    return 0;
}
