// fdtd_2d, up to five time steps per pass over memory (included by fdtd2d.cu).
//
// The one-step kernel moves 48 B per cell and step and is HBM-bound
// (fdtd_2d_numpy.py:6-11).  Here a WARP owns a strip of 128 columns (4 adjacent
// columns per lane) and marches down the rows; the NS steps are pipelined one row
// apart: when row r of the source state arrives, step s consumes row r-s of state s
// and completes row r-s-1 of state s+1, which is handed to step s+1 in registers;
// the last step's row goes to global memory.  Everything a step needs from other
// rows lives in the lane's registers (previous row's hz, new ex, new ey); the two
// lateral neighbours (old hz to the left, new ex to the right) of the lane's edge
// columns come from the neighbouring lanes by shuffle.  No shared memory, no
// barriers; warps are independent.
//
// A step's result at column j depends on columns j-1..j+1 of the state before, so
// a strip keeps NS garbage columns per side, rounded up to whole lanes: one halo lane
// per side up to four steps (120 of 128 columns stored), two for five (112 of 128);
// 16-byte aligned vector accesses when ny is even.  Rows are cut into chunks; a chunk re-runs NS rows above it (state s is
// valid from the chunk's first loaded row + s; row 0 needs no ramp because its ey is
// _fict_[t]) and NS rows below it.
//
// Traffic per cell and pass: 24 B * 128/112 read + 24 B written, i.e. 10.3 B per
// cell and step at NS = 5 against 48 B.
#pragma once

constexpr int FM_WARPS = 4;            // warps (strips) per CTA
constexpr int FM_COLS = 4;             // columns per lane
constexpr int FM_STRIP = 32 * FM_COLS; // 128 columns loaded per strip
constexpr int FM_MAX_STEPS = 5;             // six steps spill at 255 registers
// halo lanes per side for NS steps per pass (FM_COLS garbage columns per lane): 1 up to 4 steps, 2 up to 8
__host__ __device__ constexpr int fm_halo_lanes(int ns) { return (ns + FM_COLS - 1) / FM_COLS; }
__host__ __device__ constexpr int fm_out_cols(int ns) { return FM_STRIP - 2 * FM_COLS * fm_halo_lanes(ns); }   // 120 or 112 stored per strip
constexpr long long FM_AUTO_MIN_CELLS = 4000000;   // default dispatch: grids at least this large march

struct FmParams {
    long long nx, ny;          // rows of the (local) array, columns
    long long row0, nx_global; // local row 0 is global row row0 of an nx_global-row grid (row slabs: distributed.py)
    long long row_lo, row_hi;  // output rows [row_lo, row_hi) of this launch
    long long nstrips;
    int rc;                    // output rows per chunk
    int pfd;                   // L2 prefetch distance in rows (0 = off)
    const double *ex, *ey, *hz;
    double *exo, *eyo, *hzo;
    const double *fict;        // _fict_[t0 ...]
};

__device__ __forceinline__ double fm_shfl_up(double v) {
    return __shfl_up_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ double fm_shfl_down(double v) {
    return __shfl_down_sync(0xffffffffu, v, 1);
}

template <bool VEC>
__device__ __forceinline__ void fm_load_row(const double *__restrict__ g, long long ny, long long col0, bool full,
                                            double (&v)[FM_COLS]) {
    if (VEC && full) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(g));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(g) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int m = 0; m < FM_COLS; ++m) {
            const long long c = col0 + m;
            v[m] = (c >= 0 && c < ny) ? __ldg(g + m) : 0.0;
        }
    }
}

template <bool VEC>
__device__ __forceinline__ void fm_store_row(double *__restrict__ g, long long ny, long long col0, bool full,
                                             const double (&v)[FM_COLS]) {
    if (VEC && full) {
        reinterpret_cast<double2 *>(g)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2 *>(g)[1] = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int m = 0; m < FM_COLS; ++m) {
            const long long c = col0 + m;
            if (c >= 0 && c < ny) g[m] = v[m];
        }
    }
}

template <int NS, bool VEC>
__global__ void __launch_bounds__(FM_WARPS * 32)
fdtd2d_march_kernel(FmParams p) {
    const int lane = threadIdx.x & 31;
    const long long strip = (long long)blockIdx.x * FM_WARPS + (threadIdx.x >> 5);
    if (strip >= p.nstrips) return;
    const long long nx = p.nx, ny = p.ny;
    constexpr int HL = fm_halo_lanes(NS);
    const long long col0 = strip * fm_out_cols(NS) - FM_COLS * HL + FM_COLS * lane;
    const bool full = col0 >= 0 && col0 + FM_COLS <= ny;
    const bool storing = lane >= HL && lane < 32 - HL;
    const long long r0 = p.row_lo + (long long)blockIdx.y * p.rc;
    const long long r1 = (r0 + p.rc < p.row_hi) ? r0 + p.rc : p.row_hi;   // output rows [r0, r1)
    const bool top_slab = (p.row0 == 0);
    const long long r_first = (r0 - NS > 0) ? r0 - NS : 0;
    const long long r_last = r1 - 1 + NS;                            // rows >= nx are virtual (flush the pipeline)
    const long long r_load_last = (r_last < nx - 1) ? r_last : nx - 1;

    double hzp[NS][FM_COLS], exn[NS][FM_COLS], eyn[NS][FM_COLS], fict[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        fict[s] = __ldg(p.fict + s);
#pragma unroll
        for (int m = 0; m < FM_COLS; ++m) { hzp[s][m] = 0.0; exn[s][m] = 0.0; eyn[s][m] = 0.0; }
    }
    bool first_col[FM_COLS], hz_col[FM_COLS];       // ex keeps column 0, hz keeps column ny-1 (fdtd_2d_numpy.py:9-11)
#pragma unroll
    for (int m = 0; m < FM_COLS; ++m) { first_col[m] = (col0 + m == 0); hz_col[m] = (col0 + m < ny - 1); }

    double nex[FM_COLS], ney[FM_COLS], nhz[FM_COLS];
    long long src = r_first * ny + col0;              // element offset of the row being prefetched
    long long dst = (r_first - NS) * ny + col0;       // ... of the row being stored
    fm_load_row<VEC>(p.ex + src, ny, col0, full, nex);
    fm_load_row<VEC>(p.ey + src, ny, col0, full, ney);
    fm_load_row<VEC>(p.hz + src, ny, col0, full, nhz);
    for (long long r = r_first; r <= r_last; ++r, dst += ny) {
        double ex_o[FM_COLS], ey_o[FM_COLS], hz_o[FM_COLS];
#pragma unroll
        for (int m = 0; m < FM_COLS; ++m) { ex_o[m] = nex[m]; ey_o[m] = ney[m]; hz_o[m] = nhz[m]; }
        if (p.pfd > 0 && full && r + p.pfd <= r_load_last) {   // pull a row further ahead into L2 (no registers held)
            const long long a = src + (long long)p.pfd * ny;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ex + a));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ey + a));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.hz + a));
        }
        if (r + 1 <= r_load_last) {                      // prefetch the next source row
            src += ny;
            fm_load_row<VEC>(p.ex + src, ny, col0, full, nex);
            fm_load_row<VEC>(p.ey + src, ny, col0, full, ney);
            fm_load_row<VEC>(p.hz + src, ny, col0, full, nhz);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const long long q = r - s;                   // row of state s being consumed; completes row q-1 of state s+1
            const double hz_lane_left = fm_shfl_up(hz_o[FM_COLS - 1]);
            const double ex_lane_right = fm_shfl_down(exn[s][0]);
            // the first / last LOCAL row of an inner slab is a ghost row: whatever is computed there is garbage that the
            // caller's ghost depth absorbs; only the grid's own first row takes _fict_ (the last local row always keeps hz)
            const bool top = (q == 0) && top_slab, upd_hz = (q - 1 < nx - 1);
            double out_hz[FM_COLS], ex_new[FM_COLS], ey_new[FM_COLS];
#pragma unroll
            for (int m = 0; m < FM_COLS; ++m) {
                const double hz_left = m ? hz_o[m - 1] : hz_lane_left;
                const double ex_right = (m < FM_COLS - 1) ? exn[s][m + 1] : ex_lane_right;
                // :7-8  ey[0,:] = _fict_[t]; ey[1:,:] -= 0.5 * (hz[1:,:] - hz[:-1,:])
                ey_new[m] = top ? fict[s] : ey_o[m] - 0.5 * (hz_o[m] - hzp[s][m]);
                // :9    ex[:,1:] -= 0.5 * (hz[:,1:] - hz[:,:-1])
                ex_new[m] = first_col[m] ? ex_o[m] : ex_o[m] - 0.5 * (hz_o[m] - hz_left);
                // :10-11 hz[:-1,:-1] -= 0.7 * (ex[:-1,1:] - ex[:-1,:-1] + ey[1:,:-1] - ey[:-1,:-1])   (row q-1)
                const double h = hzp[s][m] - 0.7 * (((ex_right - exn[s][m]) + ey_new[m]) - eyn[s][m]);
                out_hz[m] = (upd_hz && hz_col[m]) ? h : hzp[s][m];
            }
            // keep row q for the next iteration; hand row q-1 of state s+1 to the next step
#pragma unroll
            for (int m = 0; m < FM_COLS; ++m) {
                hzp[s][m] = hz_o[m];
                hz_o[m] = out_hz[m];
                ex_o[m] = exn[s][m]; ey_o[m] = eyn[s][m];
                exn[s][m] = ex_new[m]; eyn[s][m] = ey_new[m];
            }
        }
        const long long q_out = r - NS;
        if (storing && q_out >= r0 && q_out < r1) {
            fm_store_row<VEC>(p.exo + dst, ny, col0, full, ex_o);
            fm_store_row<VEC>(p.eyo + dst, ny, col0, full, ey_o);
            fm_store_row<VEC>(p.hzo + dst, ny, col0, full, hz_o);
        }
    }
}

template <int NS>
int launch_march_ns(const FmParams &p, dim3 grid, bool vec) {
    if (vec) fdtd2d_march_kernel<NS, true><<<grid, FM_WARPS * 32, 0, npb::st().stream>>>(p);
    else fdtd2d_march_kernel<NS, false><<<grid, FM_WARPS * 32, 0, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("fdtd2d_march_kernel");
    npb::count_launch();
    return 0;
}

// one pass: ns (2..FM_MAX_STEPS) steps src -> dst over the output rows [row_lo, row_hi) of an nx-row array whose row 0
// is global row `row0` of an nx_global-row grid (whole grid: row0 = 0, nx_global = nx, all rows)
int launch_march(int ns, int64_t nx, int64_t ny, const double *ex, const double *ey, const double *hz, double *exo,
                 double *eyo, double *hzo, const double *fict_t, int rc_override, int64_t row0 = 0, int64_t nx_global = -1,
                 int64_t row_lo = 0, int64_t row_hi = -1) {
    if (nx_global < 0) nx_global = nx;
    if (row_lo < 0) row_lo = 0;
    if (row_hi < 0 || row_hi > nx) row_hi = nx;
    if (row_lo >= row_hi) return 0;
    const long long span = row_hi - row_lo;
    const long long out_cols = fm_out_cols(ns);
    const long long nstrips = (ny + out_cols - 1) / out_cols;
    const long long blocks_x = (nstrips + FM_WARPS - 1) / FM_WARPS;
    // enough row chunks for ~48 warps per SM over the whole launch; each chunk re-runs 2*ns rows
    static const int rule = getenv("NPB_FDTD_CHUNKS") ? atoi(getenv("NPB_FDTD_CHUNKS")) : 48;
    static const int env_rc = getenv("NPB_FDTD_RC") ? atoi(getenv("NPB_FDTD_RC")) : 0;
    long long chunks = ((long long)rule * npb::st().sm_count + nstrips - 1) / nstrips;
    long long rc = (span + chunks - 1) / chunks;
    if (rc < 64) rc = 64;
    if (env_rc > 0) rc = env_rc;
    if (rc_override > 0) rc = rc_override;
    if (rc > span) rc = span;
    chunks = (span + rc - 1) / rc;
    if (blocks_x >= (1LL << 31) || chunks > 65535) return npb::fail("fdtd2d", "grid too large");
    const uintptr_t bits = (uintptr_t)ex | (uintptr_t)ey | (uintptr_t)hz | (uintptr_t)exo | (uintptr_t)eyo | (uintptr_t)hzo;
    const bool vec = (ny % 2 == 0) && (bits % 16 == 0);
    // measured at 8192 x 16384: 2.94 ms without, 2.71 ms at distance 2..4, 2.86 ms at 8
    static const int pfd = getenv("NPB_FDTD_PFD") ? atoi(getenv("NPB_FDTD_PFD")) : 3;
    FmParams p{nx, ny, row0, nx_global, row_lo, row_hi, nstrips, (int)rc, pfd, ex, ey, hz, exo, eyo, hzo, fict_t};
    dim3 grid((unsigned)blocks_x, (unsigned)chunks);
    switch (ns) {
        case 2: return launch_march_ns<2>(p, grid, vec);
        case 3: return launch_march_ns<3>(p, grid, vec);
        case 4: return launch_march_ns<4>(p, grid, vec);
        case 5: return launch_march_ns<5>(p, grid, vec);
        default: return npb::fail("fdtd2d", "steps per pass out of range");
    }
}
