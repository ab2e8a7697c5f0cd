// ldg_lds.cu -- does a global (L2) load in flight delay a later shared-memory load?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldg_lds.bin tools/ldg_lds.cu && tools/ldg_lds.bin
// One CTA, W warps.  Every warp times (clock64) [a] an LDS whose result is consumed at once, [b] the same LDS issued
// right behind a relaxed gpu-scope global load whose result is NOT consumed until later, [c] the global load alone.
// If [b] ~ [a] the two complete independently; if [b] ~ [c] shared loads return in order behind global loads.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(const unsigned long long *g, long long *out, int mode) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i;
    __syncthreads();
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sh) + threadIdx.x * 8u;
    const unsigned long long *ga = g + (size_t)threadIdx.x * 32 + (size_t)blockIdx.x * 65536;
    long long acc[3] = {0, 0, 0};
    double sink = 0.0;
    unsigned long long gsink = 0;
    for (int it = 0; it < 64; ++it) {
        __syncthreads();
        long long t0 = clock64();
        double v; unsigned long long w = 0;
        if (mode == 0) {
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa) : "memory");
            sink += v;
            asm volatile("" : "+d"(sink));
            acc[0] += clock64() - t0;
        } else if (mode == 1) {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ga + (it & 15)) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa) : "memory");
            sink += v;
            asm volatile("" : "+d"(sink));
            long long t1 = clock64();
            gsink += w;
            asm volatile("" : "+l"(gsink));
            long long t2 = clock64();
            acc[1] += t1 - t0; acc[2] += t2 - t0;
        } else {
            // mode 2: only warp 0 issues the global load, the others time their shared load
            if (threadIdx.x < 32) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ga + (it & 15)) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa) : "memory");
            sink += v;
            asm volatile("" : "+d"(sink));
            long long t1 = clock64();
            gsink += w;
            asm volatile("" : "+l"(gsink));
            long long t2 = clock64();
            acc[1] += t1 - t0; acc[2] += t2 - t0;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        out[(threadIdx.x >> 5) * 4 + 0] = acc[0] / 64; out[(threadIdx.x >> 5) * 4 + 1] = acc[1] / 64;
        out[(threadIdx.x >> 5) * 4 + 2] = acc[2] / 64; out[(threadIdx.x >> 5) * 4 + 3] = (long long)sink + (long long)gsink;
    }
}

int main() {
    unsigned long long *g; long long *out, h[64];
    cudaMalloc(&g, 148ull * 65536 * 8); cudaMemset(g, 0, 148ull * 65536 * 8);
    cudaMalloc(&out, sizeof(h));
    for (int warps : {1, 4, 10}) for (int mode = 0; mode < 3; ++mode) {
        probe<<<1, warps * 32>>>(g, out, mode);
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        const int w = warps - 1;
        printf("warps=%2d mode=%d: LDS alone %lld | LDS behind LDG %lld, LDG done %lld (last warp)\n", warps, mode, h[w * 4], h[w * 4 + 1], h[w * 4 + 2]);
    }
    return cudaGetLastError() != cudaSuccess;
}
