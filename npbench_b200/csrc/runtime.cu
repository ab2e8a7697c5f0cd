// runtime.cu -- device selection, stream, caching allocator, copies, timing.
// Implements the "runtime" block of include/npb_b200.h: what NPBench's
// Framework.copy_func / copy_back_func / exec_str-sync need from a GPU plugin
// (npbench/infrastructure/framework.py:42-50, cupy_framework.py:32-58).
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace npb {

static State g_state;
State &st() { return g_state; }

int fail(const char *where, const char *msg) {
    snprintf(g_state.err, sizeof(g_state.err), "%s: %s", where, msg);
    return 1;
}

int fail_cuda(const char *where, cudaError_t e) {
    snprintf(g_state.err, sizeof(g_state.err), "%s: CUDA error %d (%s)", where, (int)e,
             cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky "last error" slot for non-fatal failures
    return (int)e ? (int)e : 1;
}

// ---- caching allocator --------------------------------------------------
// copy_func runs for every array_arg `repeat + 1` times per _execute
// (test.py:16-51), so allocations are recycled by size class; cudaFree is
// only called from npb_pool_trim / npb_shutdown.
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;       // rounded size -> block
    std::unordered_map<void *, size_t> live;         // block -> rounded size
    size_t cached_bytes = 0;
};
static Pool g_pool;

static size_t round_size(size_t b) {
    const size_t g = b < (1u << 20) ? 512 : (size_t)(2u << 20);
    return ((b + g - 1) / g) * g;
}

struct Workspace { void *p = nullptr; size_t bytes = 0; };
static Workspace g_ws[8];

void *workspace(int slot, size_t bytes) {
    Workspace &w = g_ws[slot];
    if (w.bytes >= bytes) return w.p;
    if (w.p) { cudaStreamSynchronize(st().stream); cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
    if (cudaMalloc(&w.p, bytes) != cudaSuccess) { w.p = nullptr; cudaGetLastError(); return nullptr; }
    w.bytes = bytes;
    return w.p;
}

// ---- CUDA-graph cache ------------------------------------------------------
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;     // kernels inside the graph (for npb_launch_count)
    uint64_t stamp = 0;        // LRU
};
static GraphEntry g_graphs[8];
static uint64_t g_graph_clock = 0;
static uint64_t g_capture_launch_base = 0;

static bool same_key(const GraphKey &a, const GraphKey &b) { return memcmp(&a, &b, sizeof(GraphKey)) == 0; }

bool graph_replay(const GraphKey &key) {
    for (auto &e : g_graphs)
        if (e.exec && same_key(e.key, key)) {
            if (cudaGraphLaunch(e.exec, st().stream) != cudaSuccess) { cudaGetLastError(); return false; }
            e.stamp = ++g_graph_clock;
            st().launches += e.launches;
            return true;
        }
    return false;
}

bool graph_begin() {
    if (st().stream == nullptr) return false;                       // legacy default stream cannot be captured
    if (cudaStreamBeginCapture(st().stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    g_capture_launch_base = st().launches;
    return true;
}

int graph_end_and_launch(const GraphKey &key, int rc) {
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(st().stream, &graph);
    if (rc != 0) {
        // the caller's launch loop stopped early (a host-side reject issues no CUDA error, so the capture
        // ends cleanly with a PARTIAL graph): drop it -- neither cache nor launch -- and keep the caller's
        // error message; the next identical call captures again and fails the same way
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        st().launches = g_capture_launch_base;
        return rc;
    }
    if (e != cudaSuccess || !graph) return fail_cuda("cudaStreamEndCapture", e == cudaSuccess ? cudaErrorUnknown : e);
    GraphEntry *slot = &g_graphs[0];
    for (auto &g : g_graphs) {
        if (!g.exec) { slot = &g; break; }
        if (g.stamp < slot->stamp) slot = &g;
    }
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
    e = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { slot->exec = nullptr; return fail_cuda("cudaGraphInstantiate", e); }
    slot->key = key;
    slot->launches = st().launches - g_capture_launch_base;          // counted while capturing
    slot->stamp = ++g_graph_clock;
    e = cudaGraphLaunch(slot->exec, st().stream);
    if (e != cudaSuccess) return fail_cuda("cudaGraphLaunch", e);
    return 0;
}

}  // namespace npb

using namespace npb;

extern "C" {

const char *npb_version(void) { return "0.1.0"; }
const char *npb_last_error(void) { return st().err; }

int npb_init(int device) {
    State &s = st();
    if (s.inited && (device < 0 || device == s.device)) return 0;
    if (s.inited) return fail("npb_init", "already initialised on another device (one GPU per process)");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("npb_init", "no CUDA device visible: libnpb_b200 has no CPU fallback");
    if (device < 0) device = 0;
    if (device >= n) return fail("npb_init", "device index out of range");
    NPB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    NPB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        return fail("npb_init", "this library is built for sm_100a (Blackwell B200) only");
    s.device = device;
    s.sm_count = p.multiProcessorCount;
    s.smem_optin = p.sharedMemPerBlockOptin;
    s.l2_bytes = (size_t)p.l2CacheSize;
    NPB_CUDA(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    NPB_CUDA(cudaEventCreate(&s.ev0));
    NPB_CUDA(cudaEventCreate(&s.ev1));
    s.inited = true;
    return 0;
}

int npb_shutdown(void) {
    State &s = st();
    if (!s.inited) return 0;
    cudaDeviceSynchronize();
    npb_pool_trim();
    for (auto &w : g_ws) { if (w.p) cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
    for (auto &g : g_graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    cudaEventDestroy(s.ev0); cudaEventDestroy(s.ev1);
    cudaStreamDestroy(s.own_stream);
    s = State();
    return 0;
}

int npb_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes,
                    size_t *smem_per_block_optin, size_t *total_mem) {
    NPB_REQUIRE_INIT();
    cudaDeviceProp p;
    NPB_CUDA(cudaGetDeviceProperties(&p, st().device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    if (smem_per_block_optin) *smem_per_block_optin = p.sharedMemPerBlockOptin;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return 0;
}

int npb_set_stream(void *cuda_stream) {
    NPB_REQUIRE_INIT();
    st().stream = cuda_stream ? (cudaStream_t)cuda_stream : st().own_stream;
    return 0;
}
void *npb_get_stream(void) { return (void *)st().stream; }

int npb_sync(void) {
    NPB_REQUIRE_INIT();
    NPB_CUDA(cudaStreamSynchronize(st().stream));
    return 0;
}

int npb_malloc(size_t bytes, void **dptr) {
    NPB_REQUIRE_INIT();
    NPB_ARG(dptr != nullptr, "npb_malloc", "null output pointer");
    const size_t r = round_size(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        auto it = g_pool.free_blocks.find(r);
        if (it != g_pool.free_blocks.end()) {
            *dptr = it->second;
            g_pool.free_blocks.erase(it);
            g_pool.cached_bytes -= r;
            g_pool.live[*dptr] = r;
            return 0;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, r);
    if (e != cudaSuccess) {  // release the cache and retry once
        cudaGetLastError();
        npb_pool_trim();
        e = cudaMalloc(&p, r);
        if (e != cudaSuccess) return fail_cuda("npb_malloc", e);
    }
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.live[p] = r;
    *dptr = p;
    return 0;
}

int npb_free(void *dptr) {
    if (!dptr) return 0;
    std::lock_guard<std::mutex> lk(g_pool.mu);
    auto it = g_pool.live.find(dptr);
    if (it == g_pool.live.end()) return fail("npb_free", "pointer was not allocated by npb_malloc");
    // Stream-ordered reuse: every consumer of this library enqueues on one
    // stream, so a recycled block is only touched after earlier work on it.
    g_pool.free_blocks.emplace(it->second, dptr);
    g_pool.cached_bytes += it->second;
    g_pool.live.erase(it);
    return 0;
}

int npb_pool_trim(void) {
    if (!st().inited) return 0;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (auto &kv : g_pool.free_blocks) cudaFree(kv.second);
    g_pool.free_blocks.clear();
    g_pool.cached_bytes = 0;
    return 0;
}

int npb_host_alloc(size_t bytes, void **hptr) {
    NPB_REQUIRE_INIT();
    NPB_ARG(hptr != nullptr, "npb_host_alloc", "null output pointer");
    NPB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}
int npb_host_free(void *hptr) {
    if (hptr) NPB_CUDA(cudaFreeHost(hptr));
    return 0;
}

int npb_h2d(void *dst, const void *src, size_t bytes) {
    NPB_REQUIRE_INIT();
    if (bytes) NPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st().stream));
    return 0;
}
int npb_d2h(void *dst, const void *src, size_t bytes) {
    NPB_REQUIRE_INIT();
    if (bytes) NPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st().stream));
    return 0;
}
int npb_d2d(void *dst, const void *src, size_t bytes) {
    NPB_REQUIRE_INIT();
    if (bytes) NPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st().stream));
    return 0;
}
int npb_memset(void *dst, int value, size_t bytes) {
    NPB_REQUIRE_INIT();
    if (bytes) NPB_CUDA(cudaMemsetAsync(dst, value, bytes, st().stream));
    return 0;
}

int npb_timer_start(void) {
    NPB_REQUIRE_INIT();
    NPB_CUDA(cudaEventRecord(st().ev0, st().stream));
    return 0;
}
int npb_timer_stop(float *ms) {
    NPB_REQUIRE_INIT();
    NPB_CUDA(cudaEventRecord(st().ev1, st().stream));
    NPB_CUDA(cudaEventSynchronize(st().ev1));
    float t = 0.f;
    NPB_CUDA(cudaEventElapsedTime(&t, st().ev0, st().ev1));
    if (ms) *ms = t;
    return 0;
}

uint64_t npb_launch_count(void) { return st().launches; }

int npb_l2_flush(void) {
    NPB_REQUIRE_INIT();
    const size_t bytes = st().l2_bytes ? 2 * st().l2_bytes : ((size_t)256 << 20);
    void *p = workspace(7, bytes);
    if (!p) return fail("npb_l2_flush", "cannot allocate the flush buffer");
    NPB_CUDA(cudaMemsetAsync(p, 0x5a, bytes, st().stream));
    return 0;
}

}  // extern "C"
