// jacobi_2d, up to seven sweeps per pass over memory in registers (included by jacobi2d.cu).
//
// Same scheme as fdtd2d_march.cuh: a WARP owns a strip of 128 columns (4 adjacent
// columns per lane, 16-byte accesses) and marches down the rows; sweep s of the pass
// consumes row r-s of state s and completes row r-s-1 of state s+1 from the two rows
// of state s it keeps in registers, handing the result to sweep s+1 in registers.
// The lateral neighbours of a lane's edge columns come from the neighbouring lanes
// by shuffle.  No barriers, no shared tiles (jacobi_2d_numpy.py:6-10: two sweeps
// per t, 0.2 * (c + left + right + down + up) in that order).
//
// NS is odd, so a pass goes src -> dst and state q's constant border equals dst's
// border for odd q and src's for even q (the blocked kernel's parity argument).
// Border cells are never written.  Border rows are re-read from the right array
// when a sweep completes row 0 or N-1 (a handful of row steps per chunk); the lane
// that owns column 0 or N-1 keeps the src / dst values of the last 16 rows in a
// private shared ring, loaded two rows ahead.  Row steps that can touch neither run
// a check-free instantiation, two rows at a time and skewed by one sweep, so that
// eight independent FP64 chains interleave per lane.
// NS garbage columns per side, rounded up to whole lanes: 112 of 128 columns stored
// for NS = 5 or 7, 120 for NS = 3.  Rows are cut into chunks with an NS-row ramp
// above and below; a launch may be restricted to a row range (the sharded driver's
// boundary / interior split).
//
// Traffic per cell and pass at NS = 7: 8 B * 128/112 read + 8 B written = 17.1 B,
// i.e. 2.4 B per cell update against 16 B.
#pragma once
#include <type_traits>

constexpr int JM_WARPS = 4;
constexpr int JM_COLS = 4;
constexpr int JM_STRIP = 32 * JM_COLS;
constexpr int JM_PF = 2;               // source rows in flight per lane (a row step is shorter than a DRAM round trip)
constexpr int JM_RING = 16;            // rows of border-column values kept per warp (>= NS + JM_PF + 2)
__host__ __device__ constexpr int jm_halo_lanes(int ns) { return (ns + JM_COLS - 1) / JM_COLS; }
__host__ __device__ constexpr int jm_out_cols(int ns) { return JM_STRIP - 2 * JM_COLS * jm_halo_lanes(ns); }

struct JmParams {
    long long ni, nj;
    long long nstrips;
    long long row_lo, row_hi;  // rows [row_lo, row_hi) are written by this launch (the sharded driver's ranges)
    int rc;                    // output rows per chunk
    int pfd;                   // L2 prefetch distance in rows (0 = off)
    const double *src;
    double *dst;
    double *dst2;              // DUAL instantiations: the state before the last sweep of the pass goes here (interior only)
};

template <bool VEC>
__device__ __forceinline__ void jm_load_row(const double *__restrict__ g, long long nj, long long col0, bool full,
                                            double (&v)[JM_COLS]) {
    if (VEC && full) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(g));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(g) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) {
            const long long c = col0 + m;
            v[m] = (c >= 0 && c < nj) ? __ldg(g + m) : 0.0;
        }
    }
}

// border context of one lane (edge strips and border rows only)
struct JmEdge {
    const double *src, *dst;
    long long ni, nj, col0;
    const double (*ring)[2][JM_RING];
    int bw, bm;
    bool edge_strip;
    bool inside[JM_COLS];
};

// The NS sweeps of one row step.  P[s][U] holds row q-2 of state s (q = r - s) and is overwritten with
// the incoming row q, P[s][U^1] holds row q-1; U alternates with the row, so nothing is ever moved.
// CHECKS = this row step may complete a border row, or the strip holds a border column.
template <int NS, int U, bool CHECKS>
__device__ __forceinline__ void jm_sweeps(double (&P)[NS][2][JM_COLS], double (&c)[JM_COLS], long long r,
                                          const JmEdge &e) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double left_lane = __shfl_up_sync(0xffffffffu, P[s][U ^ 1][JM_COLS - 1], 1);
        const double right_lane = __shfl_down_sync(0xffffffffu, P[s][U ^ 1][0], 1);
        double out[JM_COLS];
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) {
            const double left = m ? P[s][U ^ 1][m - 1] : left_lane;
            const double right = (m < JM_COLS - 1) ? P[s][U ^ 1][m + 1] : right_lane;
            out[m] = 0.2 * ((((P[s][U ^ 1][m] + left) + right) + c[m]) + P[s][U][m]);   // jacobi_2d_numpy.py:7-10
        }
        if (CHECKS) {
            // constant border of state s+1: dst's for odd states, src's for even ones
            const long long qm1 = r - s - 1;             // row of state s+1 completed by this sweep
            const double *border = ((s + 1) & 1) ? e.dst : e.src;
            if (qm1 == 0 || qm1 == e.ni - 1) {
#pragma unroll
                for (int m = 0; m < JM_COLS; ++m)
                    if (e.inside[m]) out[m] = __ldg(border + qm1 * e.nj + e.col0 + m);
            } else if (e.edge_strip && qm1 > 0 && qm1 < e.ni - 1) {
                if (e.bm >= 0) {
                    const double v = e.ring[e.bw][(s + 1) & 1][qm1 & (JM_RING - 1)];
#pragma unroll
                    for (int m = 0; m < JM_COLS; ++m)
                        if (m == e.bm) out[m] = v;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) { P[s][U][m] = c[m]; c[m] = out[m]; }
    }
}

// one check-free sweep s of one row: P[s][U] (row q-2) is overwritten with the incoming row c, c becomes row q-1 of
// state s+1
template <int NS, int U>
__device__ __forceinline__ void jm_sweep_fast(double (&P)[NS][2][JM_COLS], double (&c)[JM_COLS], const int s) {
    const double left_lane = __shfl_up_sync(0xffffffffu, P[s][U ^ 1][JM_COLS - 1], 1);
    const double right_lane = __shfl_down_sync(0xffffffffu, P[s][U ^ 1][0], 1);
    double out[JM_COLS];
#pragma unroll
    for (int m = 0; m < JM_COLS; ++m) {
        const double left = m ? P[s][U ^ 1][m - 1] : left_lane;
        const double right = (m < JM_COLS - 1) ? P[s][U ^ 1][m + 1] : right_lane;
        out[m] = 0.2 * ((((P[s][U ^ 1][m] + left) + right) + c[m]) + P[s][U][m]);   // jacobi_2d_numpy.py:7-10
    }
#pragma unroll
    for (int m = 0; m < JM_COLS; ++m) { P[s][U][m] = c[m]; c[m] = out[m]; }
}

// Two consecutive rows, skewed by one sweep: sweep t of row r and sweep t-1 of row r+1 touch different P[s] and
// different rows, so their FP64 chains are independent and interleave (eight chains per lane instead of four).
template <int NS>
__device__ __forceinline__ void jm_sweeps_pair(double (&P)[NS][2][JM_COLS], double (&c0)[JM_COLS],
                                               double (&c1)[JM_COLS]) {
    jm_sweep_fast<NS, 0>(P, c0, 0);
#pragma unroll
    for (int t = 1; t < NS; ++t) {
        jm_sweep_fast<NS, 0>(P, c0, t);
        jm_sweep_fast<NS, 1>(P, c1, t - 1);
    }
    jm_sweep_fast<NS, 1>(P, c1, NS - 1);
}

// DUAL: the pass also stores state NS - 1 (into p.dst2): the closing pass of kernel(TSTEPS, A, B) leaves state S in A and
// state S - 1 in B (jacobi_2d_numpy.py:8-10) without a separate single-sweep pass over memory.  After the last sweep of
// a row step P[NS - 1][u] holds the row that entered it -- row r - NS + 1 of state NS - 1 -- so nothing extra is kept.
template <int NS, bool VEC, bool DUAL>
__global__ void __launch_bounds__(JM_WARPS * 32, (NS <= 5) ? 3 : 2)
jacobi2d_march_kernel(JmParams p) {
    static_assert(NS & 1, "a pass must go src -> dst");
    constexpr int HL = jm_halo_lanes(NS);
    const int lane = threadIdx.x & 31;
    const long long strip = (long long)blockIdx.x * JM_WARPS + (threadIdx.x >> 5);
    if (strip >= p.nstrips) return;
    const long long ni = p.ni, nj = p.nj;
    const long long strip_c0 = strip * jm_out_cols(NS) - JM_COLS * HL;
    const long long col0 = strip_c0 + JM_COLS * lane;
    const bool full = col0 >= 0 && col0 + JM_COLS <= nj;
    const bool edge_strip = strip_c0 <= 0 || strip_c0 + JM_STRIP >= nj;       // holds column 0 or nj-1 (warp-uniform)
    const long long r0 = p.row_lo + (long long)blockIdx.y * p.rc;
    const long long r1 = (r0 + p.rc < p.row_hi) ? r0 + p.rc : p.row_hi;       // rows [r0, r1) of the last state
    const long long r_first = (r0 - NS > 0) ? r0 - NS : 0;
    const long long r_last = r1 - 1 + NS;                                     // rows >= ni are virtual (flush)
    const long long r_load_last = (r_last < ni - 1) ? r_last : ni - 1;

    bool inside[JM_COLS], bcol[JM_COLS], wr[JM_COLS];
    bool any_wr = false;
#pragma unroll
    for (int m = 0; m < JM_COLS; ++m) {
        const long long c = col0 + m;
        inside[m] = c >= 0 && c < nj;
        bcol[m] = (c == 0 || c == nj - 1);
        wr[m] = (lane >= HL && lane < 32 - HL) && c >= 1 && c <= nj - 2;
        any_wr = any_wr || wr[m];
    }
    const bool vec_wr = VEC && full && wr[0] && wr[JM_COLS - 1];

    double P[NS][2][JM_COLS];                          // per sweep: the two newest rows of its input state
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) { P[s][0][m] = 0.0; P[s][1][m] = 0.0; }

    // border columns (edge strips only): the lane that owns column 0 or nj-1 keeps the src / dst values of the
    // last rows in a small shared ring of its own (written and read by the same lane, no synchronisation), loaded
    // JM_PF rows ahead, so that no sweep waits for a global load
    __shared__ double ring_all[JM_WARPS][2][2][JM_RING];
    double (*ring)[2][JM_RING] = ring_all[threadIdx.x >> 5];
    int bm = -1;
#pragma unroll
    for (int m = 0; m < JM_COLS; ++m)
        if (bcol[m] && bm < 0) bm = m;
    const int bw = (bm >= 0 && col0 + bm != 0) ? 1 : 0;          // ring slot: 0 = column 0, 1 = column nj-1
    double bs_reg = 0.0, bd_reg = 0.0;

    double nxt[JM_PF][JM_COLS];
    long long src_off = r_first * nj + col0;            // element offset of the newest row in flight
    long long dst_off = (r_first - NS) * nj + col0;     // ... of the row being stored
#pragma unroll
    for (int d = 0; d < JM_PF; ++d) {
        if (r_first + d <= r_load_last) {
            if (d) src_off += nj;
            jm_load_row<VEC>(p.src + src_off, nj, col0, full, nxt[d]);
            if (edge_strip && bm >= 0) {
                ring[bw][0][(r_first + d) & (JM_RING - 1)] = __ldg(p.src + src_off + bm);
                ring[bw][1][(r_first + d) & (JM_RING - 1)] = __ldg(p.dst + src_off + bm);
            }
        } else {
#pragma unroll
            for (int m = 0; m < JM_COLS; ++m) nxt[d][m] = 0.0;
        }
    }
    long long pend_row = -1;                            // row whose border-column values sit in bs_reg / bd_reg
    // the row loop is unrolled by JM_PF so that every row in flight has its own registers (no moves that would
    // wait for the newest load): row r lives in slot (r - r_first) % JM_PF
    static_assert(JM_PF == 2, "the row loop below is written out for two rows in flight");
    JmEdge edge{p.src, p.dst, ni, nj, col0, ring, bw, bm, edge_strip, {inside[0], inside[1], inside[2], inside[3]}};
    // DUAL: row q2 of state NS - 1 -> dst2 (dst_off points at row q2 - 1)
    auto store_row2 = [&](const double (&c)[JM_COLS], const long long q2) {
        if (any_wr && q2 >= r0 && q2 < r1 && q2 >= 1 && q2 <= ni - 2) {
            double *g = p.dst2 + dst_off + nj;
            if (vec_wr) {
                reinterpret_cast<double2 *>(g)[0] = make_double2(c[0], c[1]);
                reinterpret_cast<double2 *>(g)[1] = make_double2(c[2], c[3]);
            } else {
#pragma unroll
                for (int m = 0; m < JM_COLS; ++m)
                    if (wr[m]) g[m] = c[m];
            }
        }
    };
    auto row_step = [&](const long long r, auto uc) {
        constexpr int u = decltype(uc)::value;
        double c[JM_COLS];
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) c[m] = nxt[u][m];
        if (edge_strip && bm >= 0 && pend_row >= 0) {   // values requested one row step ago
            ring[bw][0][pend_row & (JM_RING - 1)] = bs_reg;
            ring[bw][1][pend_row & (JM_RING - 1)] = bd_reg;
            pend_row = -1;
        }
        if (p.pfd > 0 && full && r + JM_PF + p.pfd <= r_load_last)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src + src_off + (long long)(1 + p.pfd) * nj));
        if (r + JM_PF <= r_load_last) {                  // refill the slot with row r + JM_PF
            src_off += nj;
            jm_load_row<VEC>(p.src + src_off, nj, col0, full, nxt[u]);
            if (edge_strip && bm >= 0) {
                bs_reg = __ldg(p.src + src_off + bm);
                bd_reg = __ldg(p.dst + src_off + bm);
                pend_row = r + JM_PF;
            }
        }
        // border rows can be completed only near the ends: rows 0 (r <= NS) and ni-1 (r >= ni-1)
        if (edge_strip || r <= NS || r >= ni - 1) jm_sweeps<NS, u, true>(P, c, r, edge);
        else jm_sweeps<NS, u, false>(P, c, r, edge);
        const long long q_out = r - NS;
        if (any_wr && q_out >= r0 && q_out < r1 && q_out >= 1 && q_out <= ni - 2) {
            double *g = p.dst + dst_off;
            if (vec_wr) {
                reinterpret_cast<double2 *>(g)[0] = make_double2(c[0], c[1]);
                reinterpret_cast<double2 *>(g)[1] = make_double2(c[2], c[3]);
            } else {
#pragma unroll
                for (int m = 0; m < JM_COLS; ++m)
                    if (wr[m]) g[m] = c[m];
            }
        }
        if (DUAL) store_row2(P[NS - 1][u], q_out + 1);
        dst_off += nj;
    };
    auto store_row = [&](const double (&c)[JM_COLS], const long long q_out) {
        if (any_wr && q_out >= r0 && q_out < r1 && q_out >= 1 && q_out <= ni - 2) {
            double *g = p.dst + dst_off;
            if (vec_wr) {
                reinterpret_cast<double2 *>(g)[0] = make_double2(c[0], c[1]);
                reinterpret_cast<double2 *>(g)[1] = make_double2(c[2], c[3]);
            } else {
#pragma unroll
                for (int m = 0; m < JM_COLS; ++m)
                    if (wr[m]) g[m] = c[m];
            }
        }
        dst_off += nj;
    };
    // two rows at once when neither can complete a border row and the strip holds no border column
    auto row_pair = [&](const long long r) {
        double c0[JM_COLS], c1[JM_COLS];
#pragma unroll
        for (int m = 0; m < JM_COLS; ++m) { c0[m] = nxt[0][m]; c1[m] = nxt[1][m]; }
        if (p.pfd > 0 && full && r + 1 + JM_PF + p.pfd <= r_load_last) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src + src_off + (long long)(1 + p.pfd) * nj));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src + src_off + (long long)(2 + p.pfd) * nj));
        }
        if (r + JM_PF <= r_load_last) {
            src_off += nj;
            jm_load_row<VEC>(p.src + src_off, nj, col0, full, nxt[0]);
        }
        if (r + 1 + JM_PF <= r_load_last) {
            src_off += nj;
            jm_load_row<VEC>(p.src + src_off, nj, col0, full, nxt[1]);
        }
        jm_sweeps_pair<NS>(P, c0, c1);
        if (DUAL) store_row2(P[NS - 1][0], r - NS + 1);
        store_row(c0, r - NS);
        if (DUAL) store_row2(P[NS - 1][1], r - NS + 2);
        store_row(c1, r + 1 - NS);
    };
    for (long long rb = r_first; rb <= r_last; rb += JM_PF) {
        if (!edge_strip && rb > NS && rb + 1 < ni - 1 && rb + 1 <= r_last) {
            row_pair(rb);
        } else {
            row_step(rb, std::integral_constant<int, 0>{});
            if (rb + 1 <= r_last) row_step(rb + 1, std::integral_constant<int, 1>{});
        }
    }
}

template <int NS>
int launch_jm_ns(const JmParams &p, dim3 grid, bool vec) {
    if (p.dst2) {
        if (vec) jacobi2d_march_kernel<NS, true, true><<<grid, JM_WARPS * 32, 0, npb::st().stream>>>(p);
        else jacobi2d_march_kernel<NS, false, true><<<grid, JM_WARPS * 32, 0, npb::st().stream>>>(p);
    } else {
        if (vec) jacobi2d_march_kernel<NS, true, false><<<grid, JM_WARPS * 32, 0, npb::st().stream>>>(p);
        else jacobi2d_march_kernel<NS, false, false><<<grid, JM_WARPS * 32, 0, npb::st().stream>>>(p);
    }
    NPB_CHECK_LAUNCH("jacobi2d_march_kernel");
    npb::count_launch();
    return 0;
}

// Rows per chunk of one marching launch over `rows` rows of an nj-column grid on `sms` SMs (pure host logic; exported as
// npb_jacobi2d_march_rows_per_chunk for the CPU tests).
inline long long jm_rows_per_chunk(int ns, long long rows, long long nj, int sms, int rc_override) {
    const long long out_cols = jm_out_cols(ns);
    const long long nstrips = (nj + out_cols - 1) / out_cols;
    const long long blocks_x = (nstrips + JM_WARPS - 1) / JM_WARPS;
    // Rows per chunk.  Short chunks win by a wide margin although every chunk re-runs 2 * ns ramp rows: the warps of
    // neighbouring strips start a chunk on the same rows and drift apart as they march, and with them the DRAM pages
    // and the shared halo columns they touch.  Measured (ns = 5 / 7, B200): the 10240 x 81920 slab 33.1 ms per 40 sweeps
    // at 1024 rows per chunk (the round-1 rule: ~48 warps per SM over the launch), 26.2 at 256, 25.6 at 192, 25.9 at
    // 128, 27.2 at 64; 16384^2: 11.1 ms at 335 rows, 9.6 at 128, 9.5 at 96, 9.6 at 64, 10.6 at 32.  Hence ~256 warps
    // per SM over the launch, at least 96 (small grids: 64) rows.  NPB_J2_CHUNKS / NPB_J2_RC override the rule / the rows per chunk.
    static const int rule = getenv("NPB_J2_CHUNKS") ? atoi(getenv("NPB_J2_CHUNKS")) : 256;
    static const int env_rc = getenv("NPB_J2_RC") ? atoi(getenv("NPB_J2_RC")) : 0;
    const long long chunks0 = ((long long)rule * sms + nstrips - 1) / nstrips;
    long long rc = (rows + chunks0 - 1) / chunks0;
    // (64 rows where 96 would leave fewer than ~12 CTAs per SM over the launch: 4096^2 2.67 ms at 64, 3.08 ms at 96)
    const long long rc_min = (blocks_x * (rows / 96) >= 12LL * sms) ? 96 : 64;
    if (rc < rc_min) rc = rc_min;
    if (env_rc > 0) rc = env_rc;
    if (rc_override > 0) rc = rc_override;
    if (rc > rows) rc = rows;
    if ((rows + rc - 1) / rc > 65535) rc = (rows + 65534) / 65535;      // gridDim.y limit: longer chunks on very tall grids
    return rc;
}

// one pass: ns (1, 3, 5 or 7) sweeps src -> dst; dst2 != nullptr: the state before the last sweep goes there as well
int launch_jm(int ns, int64_t ni, int64_t nj, const double *src, double *dst, int rc_override, int64_t row_lo = 0,
              int64_t row_hi = -1, double *dst2 = nullptr) {
    if (row_hi < 0 || row_hi > ni) row_hi = ni;
    if (row_lo < 0) row_lo = 0;
    if (row_lo >= row_hi) return 0;
    const long long rows = row_hi - row_lo;
    const long long out_cols = jm_out_cols(ns);
    const long long nstrips = (nj + out_cols - 1) / out_cols;
    const long long blocks_x = (nstrips + JM_WARPS - 1) / JM_WARPS;
    const bool vec = (nj % 2 == 0) && (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)dst2) % 16 == 0);
    const long long rc = jm_rows_per_chunk(ns, rows, nj, npb::st().sm_count, rc_override);
    const long long chunks = (rows + rc - 1) / rc;
    if (blocks_x >= (1LL << 31) || chunks > 65535) return npb::fail("jacobi2d", "grid too large");
    static const int pfd = getenv("NPB_J2_PFD") ? atoi(getenv("NPB_J2_PFD")) : 3;
    JmParams p{ni, nj, nstrips, row_lo, row_hi, (int)rc, pfd, src, dst, dst2};
    dim3 grid((unsigned)blocks_x, (unsigned)chunks);
    switch (ns) {
        case 1: return launch_jm_ns<1>(p, grid, vec);     // the closing single sweep of a big grid: a plain HBM stream
        case 3: return launch_jm_ns<3>(p, grid, vec);
        case 5: return launch_jm_ns<5>(p, grid, vec);
        case 7: return launch_jm_ns<7>(p, grid, vec);
        default: return npb::fail("jacobi2d", "sweeps per pass out of range");
    }
}
