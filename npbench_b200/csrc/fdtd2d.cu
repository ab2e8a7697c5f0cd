// fdtd2d.cu -- 2-D FDTD (ex, ey, hz), one fused kernel per time step (sm_100a).
//
// Replaces kernel(TMAX, ex, ey, hz, _fict_),
// npbench/benchmarks/polybench/fdtd_2d/fdtd_2d_numpy.py:4-11: four dependent
// whole-array statements per step (11 ufunc passes).  Here one kernel per step
// reads the old fields and writes the new ones out of place: the hz update
// needs the NEW ex(i,j+1) and ey(i+1,j), which are recomputed in registers
// from old values, so every field is read and written once per step (48 B per
// cell) and no intermediate array is materialised.  Fields ping-pong between
// the caller's arrays and a library workspace.
//
// Row-slab form: local rows [0,nrows) are global rows [row0,row0+nrows) of an
// nx_global-row grid, so the same kernel serves the halo-sharded multi-GPU
// driver (rows whose stencil leaves the slab are ghost rows: copied).
//
// Arithmetic in NumPy order, -fmad=false (oracle: npb_oracle_fdtd2d).
#include <cooperative_groups.h>

#include "common.cuh"
#include "inbox.cuh"

namespace {

constexpr int FD_THREADS = 256;

struct FdtdParams {
    long long nx_global, row0, nrows, ny;
    long long r_lo, r_hi;     // local rows [r_lo, r_hi) are updated by this launch
    const double *ex, *ey, *hz;
    double *exo, *eyo, *hzo;
    const double *fict_ptr;   // if non-null, _fict_[t] is read from here
    double fict_val;
};

__global__ void __launch_bounds__(FD_THREADS)
fdtd2d_step_kernel(FdtdParams p) {
    pdl_wait();                                // launched with pdl_launch (common.cuh)
    const long long j = (long long)blockIdx.y * FD_THREADS + threadIdx.x;
    const long long i = p.r_lo + blockIdx.x;   // local row
    if (j >= p.ny) return;
    const long long ny = p.ny;
    const long long gi = p.row0 + i;
    const long long o = i * ny + j;
    const double hz_c = __ldg(p.hz + o);
    const double ex_c = __ldg(p.ex + o);
    const double ey_c = __ldg(p.ey + o);

    // fdtd_2d_numpy.py:7-8
    double ey_n;
    if (gi == 0) {
        ey_n = p.fict_ptr ? __ldg(p.fict_ptr) : p.fict_val;
    } else if (i == 0) {
        ey_n = ey_c;                                   // ghost row of a slab: no row above
    } else {
        ey_n = ey_c - 0.5 * (hz_c - __ldg(p.hz + o - ny));
    }
    // :9
    double ex_n = ex_c;
    if (j >= 1) ex_n = ex_c - 0.5 * (hz_c - __ldg(p.hz + o - 1));
    // :10-11 (needs new ex(i,j+1) and new ey(i+1,j); both always take the update formula)
    double hz_n = hz_c;
    if (gi < p.nx_global - 1 && i < p.nrows - 1 && j < ny - 1) {
        const double ex_r = __ldg(p.ex + o + 1) - 0.5 * (__ldg(p.hz + o + 1) - hz_c);
        const double ey_d = __ldg(p.ey + o + ny) - 0.5 * (__ldg(p.hz + o + ny) - hz_c);
        hz_n = hz_c - 0.7 * (((ex_r - ex_n) + ey_d) - ey_n);
    }
    p.exo[o] = ex_n;
    p.eyo[o] = ey_n;
    p.hzo[o] = hz_n;
}

int launch_step(const FdtdParams &p) {
    const long long rows = p.r_hi - p.r_lo;
    if (rows <= 0 || p.ny <= 0) return 0;
    const long long jb = (p.ny + FD_THREADS - 1) / FD_THREADS;
    if (jb > 65535 || rows >= (1LL << 31)) return npb::fail("fdtd2d", "grid too large");
    dim3 grid((unsigned)rows, (unsigned)jb);
    static const bool pdl = !(getenv("NPB_PDL") && atoi(getenv("NPB_PDL")) == 0);
    if (pdl) {
        if (pdl_launch(fdtd2d_step_kernel, grid, dim3(FD_THREADS), 0, npb::st().stream, p) != cudaSuccess)
            return npb::fail_cuda("fdtd2d_step_kernel", cudaGetLastError());
    } else {
        fdtd2d_step_kernel<<<grid, FD_THREADS, 0, npb::st().stream>>>(p);
    }
    NPB_CHECK_LAUNCH("fdtd2d_step_kernel");
    npb::count_launch();
    return 0;
}

#include "fdtd2d_march.cuh"

// Passes of the time loop.  Marching: up to FM_MAX_STEPS steps per pass, and an even number of passes whenever
// tmax allows it, so that the result ends in the caller's arrays without a copy; the steps are spread evenly
// (the first `rem` passes take base + 1).  Otherwise one step per pass.
void fdtd_plan(int64_t tmax, bool march, int64_t *passes, int64_t *base, int64_t *rem) {
    int64_t n = tmax;
    if (march) {
        n = (tmax + FM_MAX_STEPS - 1) / FM_MAX_STEPS;
        if ((n & 1) && tmax > n) ++n;
    }
    *passes = n;
    *base = n ? tmax / n : 0;
    *rem = n ? tmax % n : 0;
}

int g_fd_mode = 0;      // 0 dispatch by size, 1 one launch per step, 2 marching passes whenever TMAX >= 2, 3 = 0
int g_fd_rc = 0;        // rows per chunk override for the marching kernel (0 = automatic)
int g_fd_last = 0;      // 1 one launch per step, 2 marching passes, 3 register-tile resident kernel

#include "fdtd2d_regtile.cuh"

// ---- register-tile resident kernel (fdtd2d_regtile.cuh): configuration, inbox arming, cooperative launch ----
struct F2Config { int rb, cb, nw, T, PI, PJ; };
struct F2Armed { unsigned long long *box = nullptr; size_t words = 0; int geo[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };
F2Armed g_f2_armed;
F2Config g_f2_last = {0, 0, 0, 0, 0, 0};

// Modelled microseconds per step (see j2_cost in jacobi2d.cu for the measurements behind the constants): per warp
// and step the MIO pipe moves 2 RB 64-bit shuffles (old hz from the left lane, new ex from the right one) and six
// rows of CB doubles through shared memory; the FP64 pipe 11 operations per cell + 3 per recomputed edge cell.
double f2_cost(const F2Config &c) {
    const double cells = (double)c.rb * c.cb;
    const double mio = 4.0 * c.rb + 12.0 * c.cb, fp64 = (11.0 * cells + 3.0 * c.cb) / 2.0;
    const double per_warp = 1.15 * (mio > fp64 ? mio : fp64) / 1965.0;
    const double step = 0.20 + 0.02 * cells + c.nw * per_warp;
    const double exchange = (c.PI * c.PJ > 1) ? 4.5 : 0.0;         // three fields: measured 4 - 6.6 us per exchange
    return step + exchange / c.T;
}

bool f2_tiles(int64_t nx, int64_t ny, int sms, F2Config &c) {
    const int cap_i = c.nw * c.rb - 2 * c.T, cap_j = 32 * c.cb - 2 * c.T;
    if (cap_i < 1 || cap_j < 1) return false;
    int64_t PI = (nx + cap_i - 1) / cap_i, PJ = (ny + cap_j - 1) / cap_j;
    if (PI * PJ > sms) return false;
    // halos come from the adjacent tiles only, and a thread's block never touches two opposite tile edges
    if (PI > 1 && nx / PI < 2 * c.T + c.rb - 1) return false;
    if (PJ > 1 && ny / PJ < 2 * c.T + c.cb - 1) return false;
    c.PI = (int)PI; c.PJ = (int)PJ;
    return true;
}

int f2_max_warps(int rb, int cb) { return rb * cb > 8 ? 16 : (rb * cb > 4 ? 16 : 20); }   // register budgets

bool pick_f2_config(int64_t tmax, int64_t nx, int64_t ny, int sms, F2Config &best) {
    if (const char *e = getenv("NPB_F2R_CFG")) {          // "rb,cb,nw,T": experiments
        F2Config c{0, 0, 0, 0, 0, 0};
        if (sscanf(e, "%d,%d,%d,%d", &c.rb, &c.cb, &c.nw, &c.T) == 4 && (c.rb == 2 || c.rb == 4 || c.rb == 8) && c.cb == 2 &&
            c.nw >= 1 && c.nw <= f2_max_warps(c.rb, c.cb) && c.T >= 1 && f2_tiles(nx, ny, sms, c)) { best = c; return true; }
        return false;
    }
    // 2 x 2 cells per thread only when forced: the model likes its many small warps at preset S, the measurement
    // (profiles/r02_f2rt_full_sweep.log) does not (0.031 vs 0.027 ms)
    static const int shapes[][2] = {{4, 2}, {8, 2}};
    double best_cost = -1.0;
    for (const auto &sh : shapes)
        for (int nw = 1; nw <= f2_max_warps(sh[0], sh[1]); ++nw)
            for (int T = 1; T <= 16 && T <= tmax; ++T) {
                F2Config c{sh[0], sh[1], nw, T, 0, 0};
                if (!f2_tiles(nx, ny, sms, c)) continue;
                const double cost = f2_cost(c);
                if (best_cost < 0.0 || cost < best_cost) { best_cost = cost; best = c; }
            }
    return best_cost >= 0.0 && best_cost <= 2.5;          // beyond that one launch per step is as fast
}

template <int RB, int CB, int MAXT>
int f2_launch(const f2rt::Params &rp, size_t smem) {
    static size_t cfg[NPB_MAX_DEVICES] = {0};                         // per device
    size_t &configured = cfg[npb::cur_device()];
    auto kern = f2rt::fdtd2d_regtile_kernel<RB, CB, MAXT>;
    if (smem > configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError(); return 0;
        }
        configured = smem;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, rp.NW * 32, smem) != cudaSuccess) {
        cudaGetLastError(); return 0;
    }
    if ((long)per_sm * npb::st().sm_count < (long)rp.PI * rp.PJ) return 0;       // all CTAs must be co-resident
    f2rt::Params q = rp;
    void *args[] = {&q};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)kern, dim3(rp.PI * rp.PJ), dim3(rp.NW * 32), args, smem,
                                                npb::st().stream);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    return 1;
}

// Returns 1 if the resident kernel ran, 0 if the grid is not eligible, < 0 on a launch error.
int try_f2_regtile(int64_t tmax, int64_t nx, int64_t ny, double *ex, double *ey, double *hz, const double *fict) {
    if (tmax < 1 || nx < 1 || ny < 1 || nx * ny > (1LL << 22)) return 0;
    F2Config c;
    if (!pick_f2_config(tmax, nx, ny, npb::st().sm_count, c)) return 0;
    const size_t cells = (size_t)c.nw * c.rb * 32 * c.cb;
    const size_t smem = (size_t)6 * (c.nw + 2) * 32 * c.cb * sizeof(double) + (size_t)c.nw * 32 * sizeof(f2rt::Desc);
    if (smem + 2048 > npb::st().smem_optin) return 0;
    const size_t words = (size_t)c.PI * c.PJ * f2rt::SLOTS * 3 * cells;
    if (words >= (1ULL << 31)) return 0;
    unsigned long long *inbox = nullptr;
    if (c.PI * c.PJ > 1) {
        inbox = (unsigned long long *)npb::workspace(4, words * sizeof(unsigned long long));
        if (!inbox) return 0;
        const int geo[8] = {c.rb * 16 + c.cb, c.nw, c.T, c.PI, c.PJ, (int)nx, (int)ny, 0};
        if (g_f2_armed.box != inbox || g_f2_armed.words != words || memcmp(g_f2_armed.geo, geo, sizeof(geo)) != 0) {
            // first call on this geometry (or the workspace moved): arm every inbox cell; a completed run leaves them armed
            f2rt::fdtd2d_inbox_arm_kernel<<<4 * npb::st().sm_count, 256, 0, npb::st().stream>>>(inbox, words);
            if (cudaGetLastError() != cudaSuccess) return -1;
            npb::count_launch();
            g_f2_armed.box = inbox; g_f2_armed.words = words; memcpy(g_f2_armed.geo, geo, sizeof(geo));
        }
    }
    f2rt::Params rp{(int)nx, (int)ny, c.PI, c.PJ, c.T, (int)tmax, c.nw, ex, ey, hz, fict, inbox};
    int r = 0;
    if (c.rb == 2) r = f2_launch<2, 2, 640>(rp, smem);
    else if (c.rb == 4) r = f2_launch<4, 2, 512>(rp, smem);
    else if (c.nw <= 12) r = f2_launch<8, 2, 384>(rp, smem);
    else r = f2_launch<8, 2, 512>(rp, smem);          // 128 registers: ~100 bytes of spills outside the step loop
    if (r == 1) { npb::count_launch(); g_f2_last = c; }
    else g_f2_armed.box = nullptr;
    return r;
}

}  // namespace

// mode & 3: 0 = dispatch by size (grids that fit on chip -- NPBench S / M / L -- run in ONE cooperative launch,
// fdtd2d_regtile_kernel; marching passes of up to five steps for grids of >= 4M cells, else one launch per step),
// 1 = always one launch per step, 2 = marching passes at any size, 3 = same as 0;
// mode >> 8: rows per chunk of the marching kernel (0 = automatic)
extern "C" int npb_fdtd2d_set_mode(int mode) {
    g_fd_mode = mode & 3;
    g_fd_rc = mode >> 8;
    return 0;
}
extern "C" int npb_fdtd2d_last_path(void) { return g_fd_last; }
// host logic only: the configuration the register-tile kernel would run a grid with on `sms` SMs (1, out6 filled) or 0
extern "C" int npb_fdtd2d_regtile_plan(int64_t tmax, int64_t nx, int64_t ny, int sms, int *out6) {
    F2Config c{0, 0, 0, 0, 0, 0};
    if (!out6 || tmax < 1 || nx < 1 || ny < 1 || nx * ny > (1LL << 22) || sms < 1) return 0;
    if (!pick_f2_config(tmax, nx, ny, sms, c)) return 0;
    out6[0] = c.rb; out6[1] = c.cb; out6[2] = c.nw; out6[3] = c.T; out6[4] = c.PI; out6[5] = c.PJ;
    return 1;
}
// configuration of the last register-tile launch: {rows, columns of cells per thread, warps per CTA, steps per
// halo exchange, tiles along i, tiles along j}
extern "C" int npb_fdtd2d_regtile_config(int *out6) {
    if (!out6) return npb::fail("npb_fdtd2d_regtile_config", "null output");
    out6[0] = g_f2_last.rb; out6[1] = g_f2_last.cb; out6[2] = g_f2_last.nw; out6[3] = g_f2_last.T;
    out6[4] = g_f2_last.PI; out6[5] = g_f2_last.PJ;
    return 0;
}

// host logic only (no device work): the pass plan npb_fdtd2d_f64 uses; writes min(passes, cap) entries
extern "C" int npb_fdtd2d_pass_plan(int64_t tmax, int march, int32_t *steps, int cap) {
    if (tmax <= 0) return 0;
    int64_t passes, base, rem;
    fdtd_plan(tmax, march != 0 && tmax >= 2, &passes, &base, &rem);
    for (int64_t q = 0; q < passes && q < cap; ++q) steps[q] = (int32_t)(base + (q < rem ? 1 : 0));
    return (int)passes;
}

extern "C" int npb_fdtd2d_step_f64(int64_t nx_global, int64_t row0, int64_t nrows, int64_t ny,
                                   const double *ex, const double *ey, const double *hz,
                                   double *ex_out, double *ey_out, double *hz_out, double fict_t,
                                   int64_t row_lo, int64_t row_hi) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nrows >= 0 && ny >= 0 && row0 >= 0 && row0 + nrows <= nx_global, "npb_fdtd2d_step_f64",
            "slab outside the grid");
    if (row_lo < 0) row_lo = 0;
    if (row_hi < 0 || row_hi > nrows) row_hi = nrows;
    FdtdParams p{nx_global, row0, nrows, ny, row_lo, row_hi, ex, ey, hz, ex_out, ey_out, hz_out, nullptr, fict_t};
    return launch_step(p);
}

// ns (2..5) steps src -> dst over the output rows [row_lo, row_hi) of a row slab (fdtd2d_march_kernel): the building
// block of the sharded driver's passes.  `fict_dev` points at _fict_[t] of the pass's first step IN DEVICE MEMORY.
// Rows within ns of a slab edge that is not a grid edge come out as garbage (ghost rows, >= ns deep by contract).
extern "C" int npb_fdtd2d_march_f64(int ns, int64_t nx_global, int64_t row0, int64_t nrows, int64_t ny,
                                    const double *ex, const double *ey, const double *hz, double *ex_out,
                                    double *ey_out, double *hz_out, const double *fict_dev, int64_t row_lo,
                                    int64_t row_hi) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nrows >= 2 && ny >= 1 && row0 >= 0 && row0 + nrows <= nx_global, "npb_fdtd2d_march_f64", "slab outside the grid");
    NPB_ARG(ns >= 2 && ns <= FM_MAX_STEPS, "npb_fdtd2d_march_f64", "steps per pass out of range (2..5)");
    return launch_march(ns, nrows, ny, ex, ey, hz, ex_out, ey_out, hz_out, fict_dev, g_fd_rc, row0, nx_global, row_lo, row_hi);
}

extern "C" int npb_fdtd2d_f64(int64_t tmax, int64_t nx, int64_t ny, double *ex, double *ey,
                              double *hz, const double *fict) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nx >= 0 && ny >= 0, "npb_fdtd2d_f64", "negative extent");
    if (tmax <= 0 || nx == 0 || ny == 0) return 0;
    if (g_fd_mode == 0 || g_fd_mode == 3) {   // grids that fit on chip: all three fields in registers for the whole time loop
        const int r = try_f2_regtile(tmax, nx, ny, ex, ey, hz, fict);
        if (r < 0) return npb::fail("npb_fdtd2d_f64", "cooperative launch of fdtd2d_regtile_kernel failed (is the GPU "
                                                      "shared with other work?); no silent fallback to the slow path");
        if (r == 1) { g_fd_last = 3; return 0; }
    }
    const size_t cells = (size_t)nx * (size_t)ny;
    double *ws = (double *)npb::workspace(0, 3 * cells * sizeof(double));
    NPB_ARG(ws != nullptr, "npb_fdtd2d_f64", "cannot allocate the ping-pong workspace");
    double *u[3] = {ex, ey, hz};
    double *w[3] = {ws, ws + cells, ws + 2 * cells};
    // passes of `ns` steps each; an even number of passes ends in the caller's arrays
    const bool march = tmax >= 2 && nx >= 2 &&
                       (g_fd_mode == 2 || (g_fd_mode == 0 && (int64_t)cells >= FM_AUTO_MIN_CELLS && ny >= 4 * FM_STRIP));
    int64_t passes, base, rem;
    fdtd_plan(tmax, march, &passes, &base, &rem);
    g_fd_last = march ? 2 : 1;
    // TMAX short dependent launches: capture once per (extents, pointers), replay as one graph
    npb::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.kind = 2; key.dims[0] = tmax; key.dims[1] = nx; key.dims[2] = ny; key.dims[3] = march ? 1 + g_fd_rc : 0;
    key.ptrs[0] = ex; key.ptrs[1] = ey; key.ptrs[2] = hz; key.ptrs[3] = fict; key.ptrs[4] = ws;
    const bool use_graph = passes > 8;
    if (use_graph && npb::graph_replay(key)) return 0;
    const bool capturing = use_graph && npb::graph_begin();
    int rc = 0;
    int64_t t = 0;
    for (int64_t q = 0; q < passes && !rc; ++q) {
        double **s = (q & 1) ? w : u;
        double **d = (q & 1) ? u : w;
        const int ns = (int)(base + (q < rem ? 1 : 0));
        if (ns == 1) {
            FdtdParams p{nx, 0, nx, ny, 0, nx, s[0], s[1], s[2], d[0], d[1], d[2], fict + t, 0.0};
            rc = launch_step(p);
        } else {
            rc = launch_march(ns, nx, ny, s[0], s[1], s[2], d[0], d[1], d[2], fict + t, g_fd_rc);
        }
        t += ns;
    }
    if (!rc && (passes & 1)) {   // result lives in the workspace: bring it home
        for (int f = 0; f < 3 && !rc; ++f)
            if (cudaMemcpyAsync(u[f], w[f], cells * sizeof(double), cudaMemcpyDeviceToDevice, npb::st().stream) !=
                cudaSuccess)
                rc = npb::fail_cuda("fdtd2d copy-back", cudaGetLastError());
    }
    if (capturing) {
        const int rc2 = npb::graph_end_and_launch(key, rc);
        if (!rc) rc = rc2;
    }
    return rc;
}
