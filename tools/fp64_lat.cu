// fp64_lat.cu -- FP64 dependent-issue latency and per-scheduler throughput on a B200 (development probe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fp64_lat tools/fp64_lat.cu && /tmp/fp64_lat
// K independent DADD chains per thread, W warps in one CTA (warp w sits on scheduler w % 4), 4096 iterations:
// cycles per DADD of one warp and warp-DADDs per cycle per scheduler.  Also SHFL.64 and LDS round trips.
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void chains(double *out, double y, int iters, long long *cyc) {
    double x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = threadIdx.x * 1e-3 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = x[k] + y;
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void shfl_chain(double *out, int iters, long long *cyc) {
    double x = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __shfl_up_sync(0xffffffffu, x, 1);
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_add_chain(double *out, double y, int iters, long long *cyc) {
    double x = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __shfl_up_sync(0xffffffffu, x, 1) + y;
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_bar_chain(double *out, int iters, long long *cyc) {
    __shared__ double sm[2][1024];
    double x = threadIdx.x;
    sm[0][threadIdx.x] = x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = sm[i & 1][(threadIdx.x + 32) % blockDim.x] + 1.0;
        sm[(i + 1) & 1][threadIdx.x] = x;
        __syncthreads();
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// MIO throughput: K independent 64-bit shuffles (2 SHFL each) / K independent LDS.128 per thread and iteration
template <int K>
__global__ void shfl_tput(double *out, int iters, long long *cyc) {
    double x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = threadIdx.x + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = __shfl_up_sync(0xffffffffu, x[k], 1);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += x[k];
    out[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K>
__global__ void lds_tput(double *out, int iters, long long *cyc) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double2 acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = make_double2(0, 0);
    const int base = (threadIdx.x * 2) & 2047;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double2 v = *reinterpret_cast<const double2 *>(sm + ((base + 2048 * (k & 1) + (i & 1) * 64) & 4095));
            acc[k].x += v.x; acc[k].y += v.y;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += acc[k].x + acc[k].y;
    out[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K>
void run(int warps, double *out, long long *cyc) {
    const int iters = 4096;
    chains<K><<<1, warps * 32>>>(out, 1e-9, iters, cyc);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = (double)c / ((double)iters * K);
    const double sched = warps < 4 ? warps : 4;
    printf("K=%2d warps=%2d: %7.2f cycles per DADD of one warp, %6.3f warp-DADD per cycle per scheduler\n", K, warps, per_warp,
           (double)iters * K * warps / sched / (double)c);
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 12);
    for (int w : {1, 4, 8, 16, 32}) {
        run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); run<32>(w, out, cyc);
    }
    long long c;
    for (int w : {1, 4, 8, 16, 32}) {
        shfl_tput<8><<<1, w * 32>>>(out, 2048, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("SHFL throughput, %2d warps x 8 independent 64-bit shuffles: %.2f SM cycles per 64-bit warp shuffle\n", w, c / (2048.0 * 8 * w));
    }
    for (int w : {1, 4, 8, 16, 32}) {
        lds_tput<8><<<1, w * 32>>>(out, 2048, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 (+2 DADD) throughput, %2d warps x 8 independent loads: %.2f SM cycles per warp LDS.128\n", w, c / (2048.0 * 8 * w));
    }
    shfl_chain<<<1, 32>>>(out, 4096, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent SHFL.64 chain: %.1f cycles per shuffle\n", c / 4096.0);
    shfl_add_chain<<<1, 32>>>(out, 1e-9, 4096, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent SHFL.64 + DADD chain: %.1f cycles per pair\n", c / 4096.0);
    for (int t : {32, 128, 512, 1024}) {
        lds_bar_chain<<<1, t>>>(out, 4096, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS + DADD + STS + __syncthreads chain, %4d threads: %.1f cycles per round\n", t, c / 4096.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
