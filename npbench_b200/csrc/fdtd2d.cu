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
#include "common.cuh"

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
    fdtd2d_step_kernel<<<grid, FD_THREADS, 0, npb::st().stream>>>(p);
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

int g_fd_mode = 0;      // 0 dispatch by size, 1 one launch per step, 2 marching passes whenever TMAX >= 2
int g_fd_rc = 0;        // rows per chunk override for the marching kernel (0 = automatic)
int g_fd_last = 0;      // 1 one launch per step, 2 marching passes

}  // namespace

// mode & 3: 0 = dispatch by size (marching passes of up to four steps for grids of >= 4M cells,
// else one launch per step), 1 = always one launch per step, 2 = marching passes at any size;
// mode >> 8: rows per chunk of the marching kernel (0 = automatic)
extern "C" int npb_fdtd2d_set_mode(int mode) {
    g_fd_mode = mode & 3;
    g_fd_rc = mode >> 8;
    return 0;
}
extern "C" int npb_fdtd2d_last_path(void) { return g_fd_last; }

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
