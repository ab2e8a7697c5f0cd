// jacobi2d.cu -- 5-point 2-D Jacobi, temporally blocked in shared memory (sm_100a).
//
// Replaces kernel(TSTEPS, A, B),
// npbench/benchmarks/polybench/jacobi_2d/jacobi_2d_numpy.py:4-10
// (2*(TSTEPS-1) sweeps ping-ponging A -> B -> A; borders of A and B are never
// written, so every odd state carries B's border and every even state A's).
//
// One launch advances the grid by `nsteps` sweeps (odd, <= 7): a CTA loads an
// output tile plus an nsteps-deep halo into shared memory, runs the sweeps on
// a shrinking region between two shared buffers, and writes the tile centre.
// Because nsteps is odd the pass always goes from one user array to the other
// (no scratch copy of a possibly 50 GB grid), and HBM sees one read + one
// write of the grid per nsteps sweeps.  Inside a sweep each thread owns a pair
// of adjacent columns and marches down its row segment with a 3-row register
// window (16-byte shared loads/stores for the pair, 8-byte for the two side
// neighbours): 24 B of shared traffic per cell update.
//
// Arithmetic: 0.2*((((c + left) + right) + down) + up), NumPy's order
// (oracle/stencil_oracle.c: jacobi2d_sweep); compiled with -fmad=false.
#include <cooperative_groups.h>

#include "common.cuh"
#include "inbox.cuh"

namespace {

constexpr int JB_HP = 8;                       // halo padding (even, >= max nsteps)
constexpr int JB_THREADS = 256;

// Tile geometry.  Big grids use the 32 x 112 tile (least halo redundancy); small grids switch
// to smaller tiles so that every SM gets work (NPBench S/M/L are 150^2 .. 700^2).
template <int C_, int TI_>
struct JacobiTile {
    static constexpr int C = C_;                          // shared tile columns
    static constexpr int TJ = C_ - 2 * JB_HP;             // output columns per tile
    static constexpr int TI = TI_;                        // output rows per tile
    static constexpr int R = TI_ + 2 * NPB_JACOBI2D_MAX_BLOCK;   // shared tile rows
    static constexpr int SEGS = JB_THREADS / (C_ / 2);    // row segments
    static constexpr int LD = (R * C_ + JB_THREADS - 1) / JB_THREADS;   // tile elements per thread
    static constexpr size_t SMEM = (size_t)2 * R * C_ * sizeof(double);
};
using TileBig = JacobiTile<128, 32>;     // 46 x 128 region, 94 KB, 2 CTAs/SM
using TileMid = JacobiTile<128, 16>;     // 30 x 128 region, 61 KB
using TileSmall = JacobiTile<64, 8>;     // 22 x 64 region, 22 KB
using Tile24 = JacobiTile<128, 24>;      // in-between heights: chosen when they avoid a nearly empty second wave
using Tile20 = JacobiTile<128, 20>;      // (700^2: 308 tiles of 16 rows on 296 CTA slots vs 245 tiles of 20 rows)
using Tile12 = JacobiTile<128, 12>;

template <class T>
__global__ void __launch_bounds__(JB_THREADS, 2)
jacobi2d_block_kernel(int nsteps, long long ni, long long nj, const double *__restrict__ src,
                      double *__restrict__ dst, long long tile_row0) {
    extern __shared__ __align__(16) double sm[];
    double *buf0 = sm;                 // states of src parity
    double *buf1 = sm + T::R * T::C;   // states of dst parity
    const int h = nsteps;
    const long long i0 = 1 + (tile_row0 + blockIdx.y) * T::TI;   // first output row
    const long long j0 = 1 + (long long)blockIdx.x * T::TJ;      // first output column
    // shared (r, c)  <->  global (i0 - MAXB + r, j0 - HP + c)
    const long long gi_base = i0 - NPB_JACOBI2D_MAX_BLOCK;
    const long long gj_base = j0 - JB_HP;

    // ---- load: rows [i0-h, i0+TI+h), cols [j0-h, j0+TJ+h), clipped to the grid
    const long long r_lo = max(0LL, i0 - h), r_hi = min(ni - 1, i0 + T::TI - 1 + h);
    const long long c_lo = max(0LL, j0 - h), c_hi = min(nj - 1, j0 + T::TJ - 1 + h);
    // All T::LD loads of a thread are issued before the first shared store, so ~23 x 8 B per
    // thread are in flight (the load phase was global-latency bound when unrolled only x4).
    {
        double v[T::LD];
#pragma unroll
        for (int u = 0; u < T::LD; ++u) {
            const int idx = threadIdx.x + u * JB_THREADS;
            const int r = idx / T::C, c = idx % T::C;
            const long long gi = gi_base + r, gj = gj_base + c;
            const bool in = idx < T::R * T::C && gi >= r_lo && gi <= r_hi && gj >= c_lo && gj <= c_hi;
            v[u] = in ? __ldg(src + gi * nj + gj) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < T::LD; ++u) {
            const int idx = threadIdx.x + u * JB_THREADS;
            const int r = idx / T::C, c = idx % T::C;
            const long long gi = gi_base + r, gj = gj_base + c;
            if (idx < T::R * T::C && gi >= r_lo && gi <= r_hi && gj >= c_lo && gj <= c_hi) {
                buf0[idx] = v[u];
                if (gi == 0 || gi == ni - 1 || gj == 0 || gj == nj - 1)
                    buf1[idx] = __ldg((const double *)dst + gi * nj + gj);   // dst's own constant border
            }
        }
    }
    __syncthreads();

    // ---- nsteps sweeps on a shrinking region
    const int pair = threadIdx.x % (T::C / 2);   // column pair: shared cols 2*pair, 2*pair+1
    const int seg = threadIdx.x / (T::C / 2);    // row segment
    const int cs0 = 2 * pair;
    for (int s = 1; s <= nsteps; ++s) {
        const double *in = (s & 1) ? buf0 : buf1;
        double *out = (s & 1) ? buf1 : buf0;
        // update region in global coordinates, clipped to the interior
        const long long ui_lo = max(1LL, i0 - h + s), ui_hi = min(ni - 2, i0 + T::TI - 1 + h - s);
        const long long uj_lo = max(1LL, j0 - h + s), uj_hi = min(nj - 2, j0 + T::TJ - 1 + h - s);
        const int rr_lo = (int)(ui_lo - gi_base), rr_hi = (int)(ui_hi - gi_base);   // shared rows
        const int cc_lo = (int)(uj_lo - gj_base), cc_hi = (int)(uj_hi - gj_base);   // shared cols
        const int nrows = rr_hi - rr_lo + 1;
        if (nrows > 0 && cs0 + 1 >= cc_lo && cs0 <= cc_hi) {
            const int per = (nrows + T::SEGS - 1) / T::SEGS;
            const int ra = rr_lo + seg * per;
            const int rb = min(rr_hi, ra + per - 1);
            if (ra <= rb) {
                const bool w0 = (cs0 >= cc_lo), w1 = (cs0 + 1 <= cc_hi);
                const double *pin = in + ra * T::C + cs0;
                double *pout = out + ra * T::C + cs0;
                double2 up = *reinterpret_cast<const double2 *>(pin - T::C);
                double2 ce = *reinterpret_cast<const double2 *>(pin);
                for (int r = ra; r <= rb; ++r) {
                    const double2 dn = *reinterpret_cast<const double2 *>(pin + T::C);
                    const double left = pin[-1];
                    const double right = pin[2];
                    double2 res;
                    res.x = 0.2 * ((((ce.x + left) + ce.y) + dn.x) + up.x);
                    res.y = 0.2 * ((((ce.y + ce.x) + right) + dn.y) + up.y);
                    if (w0 && w1) {
                        *reinterpret_cast<double2 *>(pout) = res;
                    } else if (w0) {
                        pout[0] = res.x;
                    } else {
                        pout[1] = res.y;
                    }
                    up = ce; ce = dn;
                    pin += T::C; pout += T::C;
                }
            }
        }
        __syncthreads();
    }

    // ---- store the tile centre (interior cells only) from the last buffer
    const double *fin = (nsteps & 1) ? buf1 : buf0;
    const long long o_ihi = min(ni - 2, i0 + T::TI - 1), o_jhi = min(nj - 2, j0 + T::TJ - 1);
    for (int idx = threadIdx.x; idx < T::TI * T::TJ; idx += JB_THREADS) {
        const int r = idx / T::TJ, c = idx % T::TJ;
        const long long gi = i0 + r, gj = j0 + c;
        if (gi <= o_ihi && gj <= o_jhi)
            dst[gi * nj + gj] = fin[(r + NPB_JACOBI2D_MAX_BLOCK) * T::C + (c + JB_HP)];
    }
}

// ---------------------------------------------------------------------------
// Resident variant for grids that fit on chip (NPBench presets S / M / L).
//
// One cooperative launch runs the whole time loop.  The interior is cut into PI x PJ tiles, one
// CTA (= one SM) each; a CTA keeps its tile plus a T-deep halo ring in shared memory, double
// buffered (even / odd states), and exchanges halos with its 8 neighbours only every T sweeps
// (T even, <= 8): between exchanges it updates a shrinking region (tile + T-1, ..., tile + 0
// rings), recomputing the neighbours' rim redundantly.  Halos travel through sentinel-armed L2
// inboxes (inbox.cuh).  DRAM sees the grid twice (initial load, final two states); everything else
// is shared-memory traffic: 5 loads + 1 store per cell update.
// ---------------------------------------------------------------------------
constexpr int JR_THREADS = 512;
constexpr int JR_SLOTS = 8;          // inbox ring depth, in exchanges
constexpr int JR_FENCE_EVERY = 2;    // gpu-scope fence cadence, in exchanges
constexpr int JR_RECV = 8;           // inbox cells requested per thread before the first test
constexpr int JR_TMAX = 8;

struct JacobiResidentParams {
    int ni, nj, PI, PJ, ti_max, tj_max;
    int T;                       // sweeps per exchange (even)
    int nsweeps;                 // total sweeps (even)
    double *A, *B;
    unsigned long long *inbox;   // [PI*PJ][JR_SLOTS][(ti_max+2T)*(tj_max+2T)]
    int *halo_list;              // global scratch: [PI*PJ][max_halo] ring-linear indices of the halo cells
    int max_halo;
};

__global__ void __launch_bounds__(JR_THREADS, 1)
jacobi2d_resident_kernel(JacobiResidentParams p) {
    extern __shared__ double sm[];
    __shared__ int s_nhalo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.T;
    const int ti = blockIdx.x / p.PJ, tj = blockIdx.x % p.PJ;
    int ilo, ihi, jlo, jhi;
    tile_bounds(p.ni - 2, p.PI, ti, ilo, ihi);
    tile_bounds(p.nj - 2, p.PJ, tj, jlo, jhi);
    const int nit = ihi - ilo, njt = jhi - jlo;
    const int W = p.tj_max + 2 * T;                      // region row pitch (shared and inbox)
    const size_t bufsz = (size_t)(p.ti_max + 2 * T) * W;
    double *const buf0 = sm, *const buf1 = sm + bufsz;   // even (A-parity) / odd (B-parity) states
    const size_t slot_sz = bufsz, box_sz = (size_t)JR_SLOTS * slot_sz;
    unsigned long long *my_box = p.inbox + (size_t)blockIdx.x * box_sz;
    int *halo = p.halo_list + (size_t)blockIdx.x * p.max_halo;
    // region (ri, rj) <-> global (ilo - T + ri, jlo - T + rj); own tile: ri in [T, T+nit), rj in [T, T+njt)

    for (size_t w = tid; w < box_sz; w += JR_THREADS) my_box[w] = HR_SENTINEL;
    // initial state: the T-deep region of A -> buf0 and of B -> buf1 (each parity's constant border)
    for (int w = tid; w < (nit + 2 * T) * (njt + 2 * T); w += JR_THREADS) {
        const int ri = w / (njt + 2 * T), rj = w - ri * (njt + 2 * T);
        const int gi = ilo - T + ri, gj = jlo - T + rj;
        if (gi < 0 || gi >= p.ni || gj < 0 || gj >= p.nj) continue;
        buf0[ri * W + rj] = __ldg(p.A + (long long)gi * p.nj + gj);
        buf1[ri * W + rj] = __ldg(p.B + (long long)gi * p.nj + gj);
    }
    if (tid == 0) {
        int n = 0;
        for (int ri = 0; ri < nit + 2 * T; ++ri)
            for (int rj = 0; rj < njt + 2 * T; ++rj) {
                const int gi = ilo - T + ri, gj = jlo - T + rj;
                if (gi < 1 || gi > p.ni - 2 || gj < 1 || gj > p.nj - 2) continue;      // interior cells only
                if (ri >= T && ri < T + nit && rj >= T && rj < T + njt) continue;         // own cell
                halo[n++] = ri * W + rj;
            }
        s_nhalo = n;
    }
    // the 8 neighbours: region origin and inbox base of each (for the sends)
    int nb_i0[8], nb_i1[8], nb_j0[8], nb_j1[8];
    long long nb_base[8];
    {
        int q = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                if (di == 0 && dj == 0) continue;
                const int ti2 = ti + di, tj2 = tj + dj;
                nb_base[q] = -1; nb_i0[q] = nb_i1[q] = nb_j0[q] = nb_j1[q] = 0;
                if (ti2 >= 0 && ti2 < p.PI && tj2 >= 0 && tj2 < p.PJ) {
                    int a0, a1, b0, b1;
                    tile_bounds(p.ni - 2, p.PI, ti2, a0, a1);
                    tile_bounds(p.nj - 2, p.PJ, tj2, b0, b1);
                    nb_i0[q] = a0 - T; nb_i1[q] = a1 + T; nb_j0[q] = b0 - T; nb_j1[q] = b1 + T;
                    nb_base[q] = (long long)(ti2 * p.PJ + tj2) * (long long)box_sz;
                }
                ++q;
            }
    }
    __threadfence();
    cooperative_groups::this_grid().sync();              // inboxes armed, halo lists written

    const int nhalo = s_nhalo;
    int rlin[JR_RECV];
    unsigned rmask = 0;
#pragma unroll
    for (int u = 0; u < JR_RECV; ++u) {
        const int w = u * JR_THREADS + tid;
        rlin[u] = 0;
        if (w < nhalo) { rlin[u] = halo[w]; rmask |= 1u << u; }
    }

    const int nper = (p.nsweeps + T - 1) / T;
    for (int pr = 0; pr < nper; ++pr) {
        const int Tp = min(T, p.nsweeps - pr * T);       // sweeps in this period (even)
        const bool last = (pr == nper - 1);
        if (pr > 0) {
            unsigned long long *slot = my_box + (size_t)(pr % JR_SLOTS) * slot_sz;
            unsigned pending = rmask;
            while (pending) {
                unsigned long long v[JR_RECV];
#pragma unroll
                for (int u = 0; u < JR_RECV; ++u)
                    if (pending & (1u << u)) v[u] = ld_relaxed_u64(slot + rlin[u]);
#pragma unroll
                for (int u = 0; u < JR_RECV; ++u)
                    if ((pending & (1u << u)) && v[u] != HR_SENTINEL) {
                        buf0[rlin[u]] = __longlong_as_double((long long)v[u]);
                        st_relaxed_u64(slot + rlin[u], HR_SENTINEL);       // re-arm
                        pending &= ~(1u << u);
                    }
                if (pending) __nanosleep(100);
            }
        }
        __syncthreads();
        if ((pr % JR_FENCE_EVERY) == 0) __threadfence();  // see inbox.cuh
        for (int q = 1; q <= Tp; ++q) {
            const double *src = (q & 1) ? buf0 : buf1;
            double *dst = (q & 1) ? buf1 : buf0;
            const int e = Tp - q;                        // rings around the tile still updated
            const int r_lo = max(T - e, 1 - (ilo - T)), r_hi = min(T + nit + e, (p.ni - 1) - (ilo - T));   // [r_lo, r_hi)
            const int c_lo = max(T - e, 1 - (jlo - T)), c_hi = min(T + njt + e, (p.nj - 1) - (jlo - T));
            const bool to_B = last && q == Tp - 1, to_A = last && q == Tp, send = !last && q == Tp;
            unsigned long long *out_base = p.inbox + (size_t)((pr + 1) % JR_SLOTS) * slot_sz;
            for (int r = r_lo + warp; r < r_hi; r += JR_THREADS / 32) {
                const bool own_r = (r >= T && r < T + nit);
                const int gi = ilo - T + r;
                for (int c = c_lo + lane; c < c_hi; c += 32) {
                    const double *x = src + r * W + c;
                    const double v = 0.2 * ((((x[0] + x[-1]) + x[1]) + x[W]) + x[-W]);
                    dst[r * W + c] = v;
                    const bool own = own_r && c >= T && c < T + njt;
                    if (!own) continue;
                    const int gj = jlo - T + c;
                    if (to_B) p.B[(long long)gi * p.nj + gj] = v;
                    if (to_A) p.A[(long long)gi * p.nj + gj] = v;
                    if (send && (r < 2 * T || r >= nit || c < 2 * T || c >= njt)) {     // within T of the tile rim
#pragma unroll
                        for (int n = 0; n < 8; ++n)
                            if (nb_base[n] >= 0 && gi >= nb_i0[n] && gi < nb_i1[n] && gj >= nb_j0[n] && gj < nb_j1[n])
                                st_relaxed_f64((double *)(out_base + nb_base[n] + (long long)(gi - nb_i0[n]) * W +
                                                          (gj - nb_j0[n])), v);
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Returns 1 if the resident kernel ran, 0 if the grid is not eligible.
int try_resident(int64_t nsweeps, int64_t ni, int64_t nj, double *A, double *B) {
    if (nsweeps < 8 || (nsweeps & 1) || ni * nj > (1LL << 23)) return 0;
    const int sms = npb::st().sm_count;
    const int in0 = (int)ni - 2, in1 = (int)nj - 2;
    if (in0 < 4 || in1 < 4) return 0;
    // tiles: as square as possible, PI*PJ <= #SMs, every tile at least 2 wide
    int PI = 1, PJ = 1;
    {
        long best = -1;
        for (int a = 1; a <= in0 / 2 && a <= sms; ++a) {
            int b = sms / a;
            if (b > in1 / 2) b = in1 / 2;
            if (b < 1) continue;
            const int ta = (in0 + a - 1) / a, tb = (in1 + b - 1) / b;
            const long cost = (long)(ta + 4) * (tb + 4);
            if (best < 0 || cost < best) { best = cost; PI = a; PJ = b; }
        }
    }
    const int ti_max = (in0 + PI - 1) / PI, tj_max = (in1 + PJ - 1) / PJ;
    const int ti_min = in0 / PI, tj_min = in1 / PJ;
    int T = JR_TMAX;
    for (;; T -= 2) {
        if (T < 2) return 0;
        if (T > ti_min || T > tj_min) continue;                      // halos must come from adjacent tiles
        const long region = (long)(ti_max + 2 * T) * (tj_max + 2 * T);
        if ((region - (long)ti_max * tj_max) > (long)JR_THREADS * JR_RECV) continue;
        if ((size_t)2 * region * sizeof(double) + 1024 > npb::st().smem_optin) continue;
        break;
    }
    const long region = (long)(ti_max + 2 * T) * (tj_max + 2 * T);
    const size_t smem = (size_t)2 * region * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(jacobi2d_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        configured = smem;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi2d_resident_kernel, JR_THREADS, smem) !=
        cudaSuccess) { cudaGetLastError(); return 0; }
    if ((long)per_sm * sms < (long)PI * PJ) return 0;
    const size_t box = (size_t)JR_SLOTS * region;
    const int max_halo = (int)(region - (long)ti_max * tj_max);
    const size_t inbox_bytes = box * PI * PJ * sizeof(unsigned long long);
    const size_t list_bytes = (size_t)max_halo * PI * PJ * sizeof(int);
    char *ws = (char *)npb::workspace(3, inbox_bytes + list_bytes + 256);
    if (!ws) return 0;
    JacobiResidentParams rp{(int)ni, (int)nj, PI, PJ, ti_max, tj_max, T, (int)nsweeps, A, B,
                            (unsigned long long *)ws, (int *)(ws + inbox_bytes), max_halo};
    void *args[] = {&rp};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)jacobi2d_resident_kernel, dim3(PI * PJ), dim3(JR_THREADS),
                                                args, smem, npb::st().stream);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    npb::count_launch();
    return 1;
}

int g_jacobi_mode = 0;      // 0/1 blocked passes (default: faster measured), 2 resident kernel when the grid fits on chip
int g_jacobi_last = 0;      // 1 resident, 2 blocked passes, 3 marching passes
int g_jacobi_rc = 0;        // rows per chunk override for the marching kernel

template <class T>
int launch_block_t(int nsteps, int64_t ni, int64_t nj, const double *src, double *dst, int64_t tr_lo,
                   int64_t tr_hi) {
    static bool configured = false;
    if (!configured) {
        NPB_CUDA(cudaFuncSetAttribute(jacobi2d_block_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)T::SMEM));
        configured = true;
    }
    const int64_t tiles_i = (ni - 2 + T::TI - 1) / T::TI;
    const int64_t tiles_j = (nj - 2 + T::TJ - 1) / T::TJ;
    if (tr_hi < 0 || tr_hi > tiles_i) tr_hi = tiles_i;
    if (tr_lo < 0) tr_lo = 0;
    // grid.y is limited to 65535 tile rows per launch
    for (int64_t t0 = tr_lo; t0 < tr_hi; t0 += 65535) {
        const int64_t cnt = (tr_hi - t0 < 65535) ? (tr_hi - t0) : 65535;
        dim3 grid((unsigned)tiles_j, (unsigned)cnt);
        jacobi2d_block_kernel<T><<<grid, JB_THREADS, T::SMEM, npb::st().stream>>>(
            nsteps, (long long)ni, (long long)nj, src, dst, (long long)t0);
        NPB_CHECK_LAUNCH("jacobi2d_block_kernel");
        npb::count_launch();
    }
    return 0;
}

// tile choice for whole-grid passes.  A pass is one wave of CTAs when it can be: the small tile if all its CTAs are
// co-resident (3 per SM), else the LOWEST 128-column tile whose CTAs fit the 2-per-SM slots (most CTAs in flight, no
// nearly empty second wave: 700^2 is 308 tiles of 16 rows on 296 slots but 245 tiles of 20 rows -- measured 1.12 ms vs
// 0.87 ms per call), else the big tile (least halo redundancy once there are many waves anyway).
// Measured at S / M / L / paper: small, small, 20 rows, big are the fastest of the six.
int pick_tile(int64_t ni, int64_t nj) {
    const int64_t sms = npb::st().sm_count;
    static const int forced = getenv("NPB_J2_TILE") ? atoi(getenv("NPB_J2_TILE")) : -1;
    if (forced >= 0 && forced <= 5) return forced;
    auto ctas = [&](int ti, int tj) { return ((ni - 2 + ti - 1) / ti) * ((nj - 2 + tj - 1) / tj); };
    if (ctas(TileSmall::TI, TileSmall::TJ) <= 3 * sms) return 2;
    if (ctas(Tile12::TI, Tile12::TJ) <= 2 * sms) return 5;
    if (ctas(TileMid::TI, TileMid::TJ) <= 2 * sms) return 1;
    if (ctas(Tile20::TI, Tile20::TJ) <= 2 * sms) return 4;
    if (ctas(Tile24::TI, Tile24::TJ) <= 2 * sms) return 3;
    return 0;
}

int launch_block(int tile, int nsteps, int64_t ni, int64_t nj, const double *src, double *dst, int64_t tr_lo,
                 int64_t tr_hi) {
    switch (tile) {
        case 0: return launch_block_t<TileBig>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
        case 1: return launch_block_t<TileMid>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
        case 3: return launch_block_t<Tile24>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
        case 4: return launch_block_t<Tile20>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
        case 5: return launch_block_t<Tile12>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
        default: return launch_block_t<TileSmall>(nsteps, ni, nj, src, dst, tr_lo, tr_hi);
    }
}

}  // namespace

// 0/1: temporally blocked passes (default; measured faster at S/M/L);
// 2: on-chip resident kernel when the grid is eligible (opt-in), blocked passes otherwise
#include "jacobi2d_march.cuh"

constexpr long long JM_AUTO_MIN_CELLS = 14000000;  // default dispatch: measured 2800^2 0.77x, 4096^2 1.2x, 8192^2 1.8x, 16384^2 2.0x

// mode & 7: 0 dispatch by size, 1 blocked shared-memory passes, 2 resident kernel when the grid fits,
// 3 marching passes at any size; mode >> 8: rows per chunk of the marching kernel (0 = automatic)
extern "C" int npb_jacobi2d_set_mode(int mode) { g_jacobi_mode = mode & 7; g_jacobi_rc = mode >> 8; return 0; }
extern "C" int npb_jacobi2d_last_path(void) { return g_jacobi_last; }

extern "C" int npb_jacobi2d_tile_rows(void) { return TileBig::TI; }

extern "C" int npb_jacobi2d_block_f64(int nsteps, int64_t ni, int64_t nj, const double *src,
                                      double *dst, int64_t tile_row_lo, int64_t tile_row_hi) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nsteps >= 1 && nsteps <= NPB_JACOBI2D_MAX_BLOCK && (nsteps & 1), "npb_jacobi2d_block_f64",
            "nsteps must be odd and in 1..7");
    NPB_ARG(ni >= 0 && nj >= 0, "npb_jacobi2d_block_f64", "negative extent");
    NPB_ARG(nj - 2 < (int64_t)TileBig::TJ * 2147483647LL, "npb_jacobi2d_block_f64", "row too long");
    if (ni < 3 || nj < 3) return 0;   // no interior
    // big slabs: the same rows by marching passes (tile rows -> interior rows [1 + lo*TI, 1 + hi*TI))
    if (nj >= 8 &&
        (g_jacobi_mode == 3 || (g_jacobi_mode == 0 && ni * nj >= JM_AUTO_MIN_CELLS && nj >= 4 * JM_STRIP))) {
        const int64_t tiles_i = (ni - 2 + TileBig::TI - 1) / TileBig::TI;
        if (tile_row_hi < 0 || tile_row_hi > tiles_i) tile_row_hi = tiles_i;
        if (tile_row_lo < 0) tile_row_lo = 0;
        const int64_t row_lo = 1 + tile_row_lo * TileBig::TI;
        const int64_t row_hi = (tile_row_hi >= tiles_i) ? ni - 1 : 1 + tile_row_hi * TileBig::TI;
        return launch_jm(nsteps, ni, nj, src, dst, g_jacobi_rc, row_lo, row_hi);
    }
    return launch_block(0, nsteps, ni, nj, src, dst, tile_row_lo, tile_row_hi);   // the sharded driver's big tile
}

extern "C" int npb_jacobi2d_f64(int64_t tsteps, int64_t ni, int64_t nj, double *A, double *B) {
    NPB_REQUIRE_INIT();
    NPB_ARG(ni >= 0 && nj >= 0, "npb_jacobi2d_f64", "negative extent");
    if (tsteps <= 1 || ni < 3 || nj < 3) return 0;   // range(1, TSTEPS) empty / no interior
    // 2*(TSTEPS-1) sweeps.  The last one must be a single sweep B -> A so that
    // B keeps state S-1 and A gets state S; the S-1 sweeps before it are split
    // into an ODD number of ODD-sized blocked passes (A->B, B->A, ..., A->B).
    if (g_jacobi_mode == 2 && try_resident(2 * (tsteps - 1), ni, nj, A, B) == 1) { g_jacobi_last = 1; return 0; }
    const bool march = (g_jacobi_mode == 3 || (g_jacobi_mode == 0 && ni * nj >= JM_AUTO_MIN_CELLS && nj >= 4 * JM_STRIP)) &&
                       tsteps >= 3 && nj >= 8;
    g_jacobi_last = march ? 3 : 2;
    const int64_t M = 2 * (tsteps - 1) - 1;
    static const int march_max = getenv("NPB_J2_MAXNS") ? atoi(getenv("NPB_J2_MAXNS")) : 7;
    const int64_t max_block = march ? (march_max >= 7 ? 7 : march_max >= 5 ? 5 : 3) : NPB_JACOBI2D_MAX_BLOCK;
    int64_t n = (M + max_block - 1) / max_block;
    if ((n & 1) == 0) ++n;
    int64_t extra_pairs = (M - n) / 2;            // distribute in units of 2 sweeps
    const int64_t cap = (max_block - 1) / 2;
    const int tile = pick_tile(ni, nj);
    // many short dependent passes on small grids: capture once, replay as one graph launch
    npb::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.kind = 3; key.dims[0] = tsteps; key.dims[1] = ni; key.dims[2] = nj; key.dims[3] = march ? 1 + g_jacobi_rc : 0;
    key.ptrs[0] = A; key.ptrs[1] = B;
    const bool use_graph = (n >= 8) && ni * nj <= (1LL << 24);
    if (use_graph && npb::graph_replay(key)) return 0;
    const bool capturing = use_graph && npb::graph_begin();
    double *src = A, *dst = B;
    int rc = 0;
    for (int64_t p = 0; p < n && !rc; ++p) {
        const int64_t left = n - p;
        int64_t take = (extra_pairs + left - 1) / left;   // spread evenly
        if (take > cap) take = cap;
        extra_pairs -= take;
        const int ns = (int)(1 + 2 * take);
        if (march) rc = launch_jm(ns, ni, nj, src, dst, g_jacobi_rc);
        else rc = launch_block(tile, ns, ni, nj, src, dst, 0, -1);
        double *t = src; src = dst; dst = t;
    }
    // now src == B (state S-1), dst == A
    if (!rc) rc = march ? launch_jm(1, ni, nj, src, dst, g_jacobi_rc) : launch_block(tile, 1, ni, nj, src, dst, 0, -1);
    if (capturing) {
        const int rc2 = npb::graph_end_and_launch(key, rc);
        if (!rc) rc = rc2;
    }
    return rc;
}
