// common.cuh -- shared state and helpers of libnpb_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/npb_b200.h"

namespace npb {

struct State {
    bool inited = false;
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    size_t l2_bytes = 0;
    cudaStream_t own_stream = nullptr;   // created by npb_init
    cudaStream_t stream = nullptr;       // the stream kernels are enqueued on
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    char err[512] = {0};
};

State &st();
int fail(const char *where, const char *msg);
int fail_cuda(const char *where, cudaError_t e);

// scratch buffers owned by the library, grown on demand, keyed by slot
void *workspace(int slot, size_t bytes);

inline void count_launch(int n = 1) { st().launches += (uint64_t)n; }

}  // namespace npb

#define NPB_REQUIRE_INIT()                                                        \
    do {                                                                          \
        if (!npb::st().inited) {                                                  \
            int rc_ = npb_init(-1);                                               \
            if (rc_) return rc_;                                                  \
        }                                                                         \
    } while (0)

#define NPB_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return npb::fail_cuda(#call, e_);                  \
    } while (0)

#define NPB_CHECK_LAUNCH(name)                                                    \
    do {                                                                          \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) return npb::fail_cuda(name, e_);                   \
    } while (0)

#define NPB_ARG(cond, where, msg)                                                 \
    do {                                                                          \
        if (!(cond)) return npb::fail(where, msg);                                \
    } while (0)

// Streaming (read-once) global load / store-once helpers.
__device__ __forceinline__ double ldg_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(double *p, double v) { __stcs(p, v); }
