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

// One State per device slot.  Slot 0 is the device npb_init selected (the one GPU of a normal, one-process-per-GPU
// run); npb_mg_init adds slots for the single-process multi-device entry points (npb_hdiff_f64_mg / npb_vadv_f64_mg:
// column shards need no exchange, so one process can drive all GPUs).  st() is the state of the CURRENT slot; every
// kernel launcher goes through it, so selecting a slot redirects stream, workspaces, allocator and graph cache.
static State g_states[NPB_MAX_DEVICES];
static int g_cur = 0, g_nslots = 1;
static char g_err[512] = {0};
State &st() { return g_states[g_cur]; }
int cur_slot() { return g_cur; }
int cur_device() { return g_states[g_cur].device & (NPB_MAX_DEVICES - 1); }

int fail(const char *where, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, msg);
    return 1;
}

int fail_cuda(const char *where, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, (int)e,
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
static Pool g_pools[NPB_MAX_DEVICES];
#define g_pool (g_pools[g_cur])

static size_t round_size(size_t b) {
    const size_t g = b < (1u << 20) ? 512 : (size_t)(2u << 20);
    return ((b + g - 1) / g) * g;
}

struct Workspace { void *p = nullptr; size_t bytes = 0; };
static Workspace g_ws_all[NPB_MAX_DEVICES][12];
#define g_ws (g_ws_all[g_cur])

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
    int slot = 0;              // device slot the graph was captured on
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
        if (e.exec && e.slot == g_cur && same_key(e.key, key)) {
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
    slot->slot = g_cur;
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
const char *npb_last_error(void) { return g_err; }

static int init_slot(State &s, int device) {
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

int npb_init(int device) {
    State &s = g_states[0];
    if (s.inited && (device < 0 || device == s.device)) return 0;
    if (s.inited) return fail("npb_init", "already initialised on another device (one GPU per process)");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("npb_init", "no CUDA device visible: libnpb_b200 has no CPU fallback");
    if (device < 0) device = 0;
    if (device >= n) return fail("npb_init", "device index out of range");
    g_cur = 0;
    return init_slot(s, device);
}

int npb_shutdown(void) {
    for (int k = g_nslots - 1; k >= 0; --k) {
        State &s = g_states[k];
        if (!s.inited) continue;
        g_cur = k;
        cudaSetDevice(s.device);
        cudaDeviceSynchronize();
        npb_pool_trim();
        for (auto &w : g_ws) { if (w.p) cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
        for (auto &g : g_graphs) { if (g.exec && g.slot == k) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; } }
        cudaEventDestroy(s.ev0); cudaEventDestroy(s.ev1);
        cudaStreamDestroy(s.own_stream);
        s = State();
    }
    g_cur = 0; g_nslots = 1;
    return 0;
}

// ---- single-process multi-device plumbing (column-sharded hdiff / vadv: no exchange, so no NCCL) ----------------
// Slot 0 is npb_init's device.  npb_mg_init(n, devices) makes sure slot k drives devices[k] for k < n (devices[0]
// must be slot 0's device or slot 0 must be uninitialised); a device may appear more than once (shards that share
// a GPU -- how the one-GPU test box exercises this path).  Returns 0 or an error code.
int npb_mg_init(int ndev, const int *devices) {
    NPB_ARG(ndev >= 1 && ndev <= NPB_MAX_DEVICES && devices != nullptr, "npb_mg_init", "1..16 devices");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return fail("npb_mg_init", "no CUDA device visible: libnpb_b200 has no CPU fallback");
    for (int k = 0; k < ndev; ++k)
        NPB_ARG(devices[k] >= 0 && devices[k] < n, "npb_mg_init", "device index out of range");
    if (!g_states[0].inited) { const int rc = npb_init(devices[0]); if (rc) return rc; }
    NPB_ARG(g_states[0].device == devices[0], "npb_mg_init", "devices[0] must be the device npb_init selected");
    for (int k = 1; k < ndev; ++k) {
        State &s = g_states[k];
        if (s.inited && s.device == devices[k]) continue;
        NPB_ARG(!s.inited, "npb_mg_init", "slot already drives another device");
        const int rc = init_slot(s, devices[k]);
        if (rc) { cudaSetDevice(g_states[0].device); return rc; }
    }
    if (ndev > g_nslots) g_nslots = ndev;
    g_cur = 0;
    NPB_CUDA(cudaSetDevice(g_states[0].device));
    return 0;
}
int npb_mg_count(void) { return g_nslots; }
// make slot k current (stream, allocator, workspaces, graph cache follow); k = 0 returns to the primary device
int npb_mg_select(int slot) {
    NPB_ARG(slot >= 0 && slot < g_nslots && g_states[slot].inited, "npb_mg_select", "slot not initialised");
    g_cur = slot;
    NPB_CUDA(cudaSetDevice(g_states[slot].device));
    return 0;
}
int npb_mg_current(void) { return g_cur; }

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
    if (g_nslots == 1) {
        NPB_CUDA(cudaStreamSynchronize(st().stream));
        return 0;
    }
    const int keep = g_cur;                     // every device this process drives
    for (int k = 0; k < g_nslots; ++k) {
        if (!g_states[k].inited) continue;
        NPB_CUDA(cudaSetDevice(g_states[k].device));
        NPB_CUDA(cudaStreamSynchronize(g_states[k].stream));
    }
    NPB_CUDA(cudaSetDevice(g_states[keep].device));
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
    for (int k = 0; k < g_nslots; ++k) {            // the block goes back to the pool of the slot that allocated it
        Pool &pool = g_pools[k];
        std::lock_guard<std::mutex> lk(pool.mu);
        auto it = pool.live.find(dptr);
        if (it == pool.live.end()) continue;
        // Stream-ordered reuse: every consumer of this library enqueues on one
        // stream per device, so a recycled block is only touched after earlier work on it.
        pool.free_blocks.emplace(it->second, dptr);
        pool.cached_bytes += it->second;
        pool.live.erase(it);
        return 0;
    }
    return fail("npb_free", "pointer was not allocated by npb_malloc");
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
