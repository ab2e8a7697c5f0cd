// common.cuh -- shared state and helpers of libnpb_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/npb_b200.h"

#define NPB_MAX_DEVICES 16

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
};

State &st();          // state of the CURRENT device slot (slot 0 unless npb_mg_select changed it)
int cur_slot();
int cur_device();     // CUDA ordinal of the current slot's device, < NPB_MAX_DEVICES (function attributes are per device)
int fail(const char *where, const char *msg);
int fail_cuda(const char *where, cudaError_t e);

// scratch buffers owned by the library, grown on demand, keyed by slot
void *workspace(int slot, size_t bytes);

inline void count_launch(int n = 1) { st().launches += (uint64_t)n; }

// jacobi2d.cu: host-buffer call of a grid in the marching regime, pipelined over row chunks (H2D of chunk c + 1, the
// passes skewed by one chunk each, D2H of finished chunks all overlap).  1 = done, 0 = not eligible, < 0 error.
int jacobi2d_host_pipelined(int64_t tsteps, int64_t ni, int64_t nj, double *A, double *B);

// ---- CUDA-graph cache for launch-bound time loops (runtime.cu) --------------------
// A time loop of hundreds of tiny dependent launches is bound by launch cost.  The entry
// points capture their launch sequence once per (kernel family, extents, step count,
// device pointers) and replay it with one cudaGraphLaunch afterwards.
struct GraphKey {
    int kind;
    long long dims[4];
    const void *ptrs[7];
};
// true: a cached graph for `key` was launched (nothing else to do)
bool graph_replay(const GraphKey &key);
// begin capturing the current stream; false if capture is not possible (then launch directly)
bool graph_begin();
// end capture; rc == 0: instantiate, cache under `key` and launch once (returns 0 or a CUDA error code);
// rc != 0 (the caller's launch loop failed): discard the partial graph and return rc
int graph_end_and_launch(const GraphKey &key, int rc);

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

// Programmatic dependent launch for chains of short dependent kernels (one launch per sweep / time step, replayed as
// a CUDA graph): the next kernel of the chain is launched while this one still runs -- its CTAs take their SM slots
// as soon as these free up and wait in pdl_wait() until this grid has completed and its stores are visible -- so the
// ~1.5 us of launch latency between two dependent graph nodes overlaps with the predecessor instead of following it.
// A kernel launched with pdl_launch() MUST call pdl_wait() before its first access to memory a predecessor wrote or
// read, on every path.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// (Releasing the successor early with griddepcontrol.launch_dependents was measured slower both for grids that fill
// the machine -- it steals SM slots from this grid's own CTAs: heat_3d `paper` 5.6 -> 7.1 ms -- and for the tiny
// grids of cavity_flow: 52.4 -> 54.7 ms.)
template <class... KArgs, class... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Streaming (read-once) global load / store-once helpers.
__device__ __forceinline__ double ldg_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(double *p, double v) { __stcs(p, v); }
