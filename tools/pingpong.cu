// pingpong.cu -- latency floor of SM-to-SM signalling on a B200 (development probe, not part of the library).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pingpong tools/pingpong.cu && /tmp/pingpong
//
// Every resident-stencil design in csrc/ pays one neighbour exchange per sweep; this measures what one exchange
// costs by each mechanism, so that the per-sweep floor is a number and not a guess:
//   (1) L2 ping-pong between two CTAs: st.relaxed.gpu of a value, ld.relaxed.gpu spin on it (the inbox protocol)
//   (2) the same between all CTAs of a 12 x 12 tile grid, 4 neighbours each (the real pattern), 8-byte and 16-byte polls
//   (3) DSMEM ping-pong inside a 2-CTA cluster: st.shared::cluster into the peer + local volatile spin
//   (4) hardware cluster barrier (barrier.cluster.arrive.release + wait.acquire) per iteration, clusters of 2..16
//   (5) cooperative grid.sync() per iteration
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// (1) two CTAs, one thread each: CTA 0 writes iteration number to flag[1], CTA 1 answers on flag[0]
__global__ void l2_pingpong(unsigned long long *flag, int iters, long long *cycles, int mode) {
    const int me = blockIdx.x;
    if (threadIdx.x != 0) return;
    unsigned long long *mine = flag + 32 * me, *peer = flag + 32 * (1 - me);
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        if (me == 0) {
            st_relaxed(peer, (unsigned long long)it);
            while ((mode ? ld_volatile(mine) : ld_relaxed(mine)) != (unsigned long long)it) {}
        } else {
            while ((mode ? ld_volatile(mine) : ld_relaxed(mine)) != (unsigned long long)it) {}
            st_relaxed(peer, (unsigned long long)it);
        }
    }
    if (me == 0) *cycles = clock64() - t0;
}

// (2) PI x PJ CTAs; per iteration every CTA sends `vals` values per side to its 4 neighbours and waits for theirs.
//     box layout: [cta][parity 2][side 4][vals]; a value is the iteration number (so no re-arm is needed)
__global__ void l2_halo(unsigned long long *box, int PI, int PJ, int vals, int iters, long long *cycles, int vec, int work) {
    const int cta = blockIdx.x, ti = cta / PJ, tj = cta % PJ;
    const bool has[4] = {ti > 0, ti < PI - 1, tj > 0, tj < PJ - 1};
    const int nb[4] = {cta - PJ, cta + PJ, cta - 1, cta + 1};
    const int opp[4] = {1, 0, 3, 2};
    const size_t side = vals, slot = 4 * side, per = 2 * slot;
    const long long t0 = clock64();
    double acc = threadIdx.x;
    for (int it = 1; it <= iters; ++it) {
        const size_t so = (size_t)(it & 1) * slot;
        for (int s = 0; s < 4; ++s)
            if (has[s])
                for (int v = threadIdx.x * vec; v < vals; v += blockDim.x * vec) {
                    if (vec == 2)
                        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(box + nb[s] * per + so + opp[s] * side + v),
                                     "l"((unsigned long long)it) : "memory");
                    else
                        st_relaxed(box + nb[s] * per + so + opp[s] * side + v, (unsigned long long)it);
                }
        for (int w = 0; w < work; ++w) acc = acc * 1.0000001 + 0.5;     // stand-in for the halo-independent update
        for (int s = 0; s < 4; ++s)
            if (has[s])
                for (int v = threadIdx.x * vec; v < vals; v += blockDim.x * vec) {
                    const unsigned long long *p = box + cta * per + so + s * side + v;
                    if (vec == 2) {
                        unsigned long long a, b;
                        do {
                            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
                        } while (a != (unsigned long long)it || b != (unsigned long long)it);
                    } else {
                        while (ld_relaxed(p) != (unsigned long long)it) {}
                    }
                }
        __syncthreads();
    }
    if (cta == 0 && threadIdx.x == 0) { cycles[0] = clock64() - t0; cycles[1] = (long long)acc; }
}

// (3) DSMEM ping-pong in a 2-CTA cluster
__global__ void __cluster_dims__(2, 1, 1) dsmem_pingpong(int iters, long long *cycles) {
    __shared__ unsigned long long flag;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned me = cl.block_rank();
    if (threadIdx.x == 0) flag = 0;
    cl.sync();
    unsigned long long *peer = cl.map_shared_rank(&flag, 1 - me);
    volatile unsigned long long *mine = &flag;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int it = 1; it <= iters; ++it) {
            if (me == 0) {
                *peer = (unsigned long long)it;
                while (*mine != (unsigned long long)it) {}
            } else {
                while (*mine != (unsigned long long)it) {}
                *peer = (unsigned long long)it;
            }
        }
        if (me == 0) *cycles = clock64() - t0;
    }
    cl.sync();
}

// (3b) DSMEM halo in clusters of CS CTAs arranged in a ring: every CTA writes `vals` doubles into both ring neighbours'
//      shared memory, then a cluster barrier; per-iteration time
template <int CS>
__global__ void dsmem_halo(int vals, int iters, long long *cycles, int work) {
    extern __shared__ double buf[];      // [2 parities][2 sides][vals]
    cg::cluster_group cl = cg::this_cluster();
    const unsigned me = cl.block_rank();
    double *left = cl.map_shared_rank(buf, (me + CS - 1) % CS), *right = cl.map_shared_rank(buf, (me + 1) % CS);
    cl.sync();
    const long long t0 = clock64();
    double acc = threadIdx.x;
    for (int it = 1; it <= iters; ++it) {
        const int so = (it & 1) * 2 * vals;
        for (int v = threadIdx.x; v < vals; v += blockDim.x) {
            left[so + vals + v] = acc + v;
            right[so + v] = acc - v;
        }
        for (int w = 0; w < work; ++w) acc = acc * 1.0000001 + 0.5;
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        acc += buf[so + (threadIdx.x % vals)] + buf[so + vals + (threadIdx.x % vals)];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { cycles[0] = clock64() - t0; cycles[1] = (long long)acc; }
}

// (6) cost of a gpu-scope fence after a global store: fence.sc (what __threadfence() is) vs fence.acq_rel
__global__ void fence_cost(unsigned long long *buf, int iters, long long *cycles, int kind) {
    unsigned long long *p = buf + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        st_relaxed(p, (unsigned long long)it);
        if (kind == 0) asm volatile("fence.sc.gpu;" ::: "memory");
        else if (kind == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

// (5) cooperative grid sync
__global__ void grid_sync_loop(int iters, long long *cycles) {
    cg::grid_group g = cg::this_grid();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

template <int CS>
void run_dsmem_halo(int threads, int vals, int iters, long long *d_cycles, int work, int nclusters) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * nclusters); cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = (size_t)4 * vals * sizeof(double);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (CS > 8) CK(cudaFuncSetAttribute(dsmem_halo<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(dsmem_halo<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, dsmem_halo<CS>, vals, 10, d_cycles, work));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, dsmem_halo<CS>, vals, iters, d_cycles, work));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dsmem_halo cluster=%2d x%3d clusters threads=%4d vals/side=%5d work=%4d : %.3f us / iteration\n", CS, nclusters, threads, vals,
           work, ms * 1e3 / iters);
}

int main() {
    long long *d_cycles; CK(cudaMalloc(&d_cycles, 16));
    unsigned long long *flag; CK(cudaMalloc(&flag, 1 << 20));
    int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const int iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms; long long cyc[2];
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(flag, 0, 1 << 20));
        l2_pingpong<<<2, 32>>>(flag, 100, d_cycles, mode);
        CK(cudaMemset(flag, 0, 1 << 20));
        CK(cudaEventRecord(e0));
        l2_pingpong<<<2, 32>>>(flag, iters, d_cycles, mode);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(cyc, d_cycles, 8, cudaMemcpyDeviceToHost));
        printf("l2_pingpong (%s): round trip %.3f us = %lld cycles (one way %.3f us)\n", mode ? "ld.volatile" : "ld.relaxed.gpu",
               ms * 1e3 / iters, cyc[0] / iters, ms * 1e3 / iters / 2);
    }
    unsigned long long *box; CK(cudaMalloc(&box, (size_t)64 << 20));
    for (int vec = 1; vec <= 2; ++vec)
        for (int threads : {32, 128, 320, 512})
            for (int vals : {64, 408})
                for (int work : {0, 300}) {
                    const int PI = 12, PJ = 12;
                    CK(cudaMemset(box, 0, (size_t)64 << 20));
                    void *args[] = {&box, (void *)&PI, (void *)&PJ, &vals, (void *)&iters, &d_cycles, &vec, &work};
                    int warm = 10;
                    void *wargs[] = {&box, (void *)&PI, (void *)&PJ, &vals, &warm, &d_cycles, &vec, &work};
                    CK(cudaLaunchCooperativeKernel((void *)l2_halo, dim3(PI * PJ), dim3(threads), wargs, 0, 0));
                    CK(cudaMemset(box, 0, (size_t)64 << 20));
                    CK(cudaEventRecord(e0));
                    CK(cudaLaunchCooperativeKernel((void *)l2_halo, dim3(PI * PJ), dim3(threads), args, 0, 0));
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("l2_halo 12x12 CTAs threads=%3d vals/side=%3d vec=%d work=%3d : %.3f us / iteration\n", threads, vals, vec, work,
                           ms * 1e3 / iters);
                }
    dsmem_pingpong<<<2, 32>>>(100, d_cycles);
    CK(cudaEventRecord(e0));
    dsmem_pingpong<<<2, 32>>>(iters, d_cycles);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dsmem_pingpong: round trip %.3f us (one way %.3f us)\n", ms * 1e3 / iters, ms * 1e3 / iters / 2);
    for (int work : {0, 300}) {
        run_dsmem_halo<2>(320, 408, iters, d_cycles, work, 64);
        run_dsmem_halo<4>(320, 408, iters, d_cycles, work, 32);
        run_dsmem_halo<8>(320, 408, iters, d_cycles, work, 16);
        run_dsmem_halo<16>(320, 408, iters, d_cycles, work, 8);
        run_dsmem_halo<16>(1024, 408, iters, d_cycles, work, 8);
        run_dsmem_halo<16>(320, 64, iters, d_cycles, work, 1);
    }
    {
        int it2 = 500;
        void *args[] = {&it2, &d_cycles};
        for (int threads : {32, 320}) {
            CK(cudaLaunchCooperativeKernel((void *)grid_sync_loop, dim3(144), dim3(threads), args, 0, 0));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void *)grid_sync_loop, dim3(144), dim3(threads), args, 0, 0));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid.sync 144 CTAs x %d threads: %.3f us / sync\n", threads, ms * 1e3 / it2);
        }
    }
    for (int kind = 0; kind < 3; ++kind) {
        fence_cost<<<144, 320>>>(box, 200, d_cycles, kind);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(cyc, d_cycles, 8, cudaMemcpyDeviceToHost));
        printf("store + %s, 144 CTAs x 320 threads: %lld cycles per iteration\n", kind == 0 ? "fence.sc.gpu" : kind == 1 ? "fence.acq_rel.gpu" : "no fence", cyc[0] / 200);
    }
    printf("SM clock %.0f MHz\n", khz / 1e3);
    return 0;
}
