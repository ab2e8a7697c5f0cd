// hdiff.cu -- COSMO horizontal diffusion, fused single pass (sm_100a).
//
// Replaces hdiff(in_field, out_field, coeff),
// npbench/benchmarks/weather_stencils/hdiff/hdiff_numpy.py:5-29: the reference
// materialises lap_field, flx_field and fly_field as full-size temporaries and
// makes 21 ufunc passes; here Laplacian, both flux limiters and the output
// stage are fused in registers -- `in` and `coeff` are read once, `out` is
// written once, nothing else touches HBM.
//
// Layout: in (I+4, J+4, K), out/coeff (I, J, K), K contiguous.  (j,k) is
// flattened to one contiguous axis: output column c = j*K + k reads the input
// row at c + 2K + dq*K for dq in -2..2, so every access is unit-stride across
// lanes for any K.  Each thread owns one flattened column and marches along i
// with a rolling register window (13-point footprint; 5 new loads per output);
// the Laplacians of column q and the x-flux are carried from row to row.
//
// Arithmetic follows NumPy's evaluation order exactly (see oracle/
// stencil_oracle.c: npb_oracle_hdiff) and the TU is compiled with -fmad=false.
#include "common.cuh"

namespace {

constexpr int HD_THREADS = 256;

__device__ __forceinline__ double lap5(double c, double ip, double im, double jp, double jm) {
    // hdiff_numpy.py:7-9   4*in[c] - (in[i+1] + in[i-1] + in[j+1] + in[j-1])
    return 4.0 * c - (((ip + im) + jp) + jm);
}

__device__ __forceinline__ double limit(double r, double din) {
    // hdiff_numpy.py:12-17 / 20-25   np.where(res * d_in > 0, 0, res)
    return (r * din > 0.0) ? 0.0 : r;
}

// rows_per_chunk output rows per thread; grid.y = number of chunks.
__global__ void __launch_bounds__(HD_THREADS)
hdiff_march_kernel(int I, int JK, int K, long long in_pitch,   // in_pitch = (J+4)*K
                   const double *__restrict__ in, double *__restrict__ out,
                   const double *__restrict__ coeff, int rows_per_chunk) {
    const int c = blockIdx.x * HD_THREADS + threadIdx.x;
    if (c >= JK) return;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int i1 = min(I, i0 + rows_per_chunk);
    if (i0 >= i1) return;

    // pointer to in[(i0+2), q, k] : centre of the first output row
    const double *pc = in + (long long)(i0 + 2) * in_pitch + c + 2 * (long long)K;
    const long long P = in_pitch;

    // window: column q rows p-2..p+2; columns q-1/q+1 rows p-1..p+1; q-2/q+2 row p
    double c_m2 = __ldg(pc - 2 * P), c_m1 = __ldg(pc - P), c_0 = __ldg(pc), c_p1 = __ldg(pc + P),
           c_p2 = __ldg(pc + 2 * P);
    double l_m1 = __ldg(pc - P - K), l_0 = __ldg(pc - K), l_p1 = __ldg(pc + P - K);
    double r_m1 = __ldg(pc - P + K), r_0 = __ldg(pc + K), r_p1 = __ldg(pc + P + K);
    double ll_0 = __ldg(pc - 2 * K), rr_0 = __ldg(pc + 2 * K);

    double lap_m = lap5(c_m1, c_0, c_m2, r_m1, l_m1);   // lap(p-1, q)
    double lap_c = lap5(c_0, c_p1, c_m1, r_0, l_0);     // lap(p,   q)
    double flx_m = limit(lap_c - lap_m, c_0 - c_m1);    // flux between (p-1,q) and (p,q)

    const double *pco = coeff + (long long)i0 * JK + c;
    double *po = out + (long long)i0 * JK + c;

    for (int i = i0; i < i1; ++i) {
        // prefetch the 5 new values of the next row's window (rows exist up to I+3)
        double n_c_p2 = 0.0, n_l_p1 = 0.0, n_r_p1 = 0.0, n_ll = 0.0, n_rr = 0.0;
        const bool more = (i + 1 < i1);
        if (more) {
            n_c_p2 = __ldg(pc + 3 * P);
            n_l_p1 = __ldg(pc + 2 * P - K);
            n_r_p1 = __ldg(pc + 2 * P + K);
            n_ll = __ldg(pc + P - 2 * K);
            n_rr = __ldg(pc + P + 2 * K);
        }
        const double cf = ldg_stream(pco);

        const double lap_p = lap5(c_p1, c_p2, c_0, r_p1, l_p1);   // lap(p+1, q)
        const double lap_r = lap5(r_0, r_p1, r_m1, rr_0, c_0);    // lap(p, q+1)
        const double lap_l = lap5(l_0, l_p1, l_m1, c_0, ll_0);    // lap(p, q-1)

        const double flx_c = limit(lap_p - lap_c, c_p1 - c_0);
        const double fly_c = limit(lap_r - lap_c, r_0 - c_0);
        const double fly_m = limit(lap_c - lap_l, c_0 - l_0);

        // hdiff_numpy.py:27-29
        const double res = c_0 - cf * (((flx_c - flx_m) + fly_c) - fly_m);
        stg_stream(po, res);

        // roll the window one row down
        c_m2 = c_m1; c_m1 = c_0; c_0 = c_p1; c_p1 = c_p2; c_p2 = n_c_p2;
        l_m1 = l_0; l_0 = l_p1; l_p1 = n_l_p1;
        r_m1 = r_0; r_0 = r_p1; r_p1 = n_r_p1;
        ll_0 = n_ll; rr_0 = n_rr;
        lap_m = lap_c; lap_c = lap_p; flx_m = flx_c;
        pc += P; pco += JK; po += JK;
    }
}


// ---------------------------------------------------------------------------
// v2: persistent, TMA-bulk fed ring (used when the problem is large enough).
//
// The marching kernel above is bound by global-load latency (ncu: 75% of the
// stall samples are long-scoreboard waits at ~45% occupancy).  Here memory
// level parallelism no longer depends on occupancy: one producer thread per
// CTA streams whole input rows (and the matching coeff rows) into a ring of
// shared-memory slots with cp.async.bulk (the TMA engine, 1-D bulk form),
// completion tracked by mbarriers; 15 consumer warps take each row out of the
// ring exactly once -- 6 x 8-byte shared loads per thread (columns q-2..q+3 of its
// two j-adjacent cells) into a register window that carries the rows still needed -- so a slot is
// released the moment it has been read and all other slots are look-ahead
// (up to 7 rows, ~140 KB in flight per SM).  Work units (i-block, j-tile) are dealt
// round-robin so that concurrently running CTAs share their halos through L2.
// ---------------------------------------------------------------------------
namespace ring {

constexpr int CONSUMERS = 480;               // 15 warps, two j-adjacent cells per thread (512 threads => 128 regs each)
constexpr int THREADS = CONSUMERS + 32;      // + producer warp

struct Params {
    int I, J, K, TJ, n_jtiles, RB, n_iblocks;
    int slot_in_elems;        // (TJ+4)*K
    int slot_elems;           // slot stride in doubles (in row + coeff row, 128-byte multiple)
    const double *in;
    double *out;
    const double *coeff;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One consumer step: row `r` of the current unit has landed in `row`; W is the static position of that row in
// the 4-deep circular register windows (W == (r - i0) & 3 after 4x unrolling), so no register is ever moved.
// A thread owns TWO j-ADJACENT output cells A = (q, k) and B = (q+1, k) and keeps the six input columns
// q-2 .. q+3 (X[0..5]; X[2] = A, X[3] = B): the Laplacian of each of its cells is carried from row to row, so
// lap(p, q+1) for A and lap(p, q-1) for B come for free, and the y-flux between A and B is computed once --
// 4 Laplacians and 5 limiters per two cells instead of 6 and 6 (25 instead of 32 flops per cell; the consumer
// warps were issue bound).  Same per-cell expressions in the same order, so the results are bit-identical.
template <int W>
__device__ __forceinline__ void consume_row(const Params &p, const double *row, int r, int i0, int j0, int ebase,
                                            bool actA, bool actB, double (&X)[6][4], double &lapA, double &lapB,
                                            double &flxmA, double &flxmB, unsigned long long *empty_bar, int lane) {
    const int K = p.K;
    double cfA = 0.0, cfB = 0.0;
    if (actA) {
        const double *c = row + ebase + 2 * K;        // in[r, q, k]
#pragma unroll
        for (int d = 0; d < 6; ++d) X[d][W] = c[(d - 2) * K];
        if (r - 4 >= i0) {
            cfA = row[p.slot_in_elems + ebase];
            if (actB) cfB = row[p.slot_in_elems + ebase + K];
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar);            // this warp is done with the slot
    // rows: r -> W, r-1 -> W+3, r-2 (centre row p) -> W+2, r-3 -> W+1   (indices mod 4)
    constexpr int P2 = W, P1 = (W + 3) & 3, P0 = (W + 2) & 3, M1 = (W + 1) & 3;
    const double lap_pA = lap5(X[2][P1], X[2][P2], X[2][P0], X[3][P1], X[1][P1]);     // lap(p+1, q)
    const double lap_pB = lap5(X[3][P1], X[3][P2], X[3][P0], X[4][P1], X[2][P1]);     // lap(p+1, q+1)
    const double lap_l = lap5(X[1][P0], X[1][P1], X[1][M1], X[2][P0], X[0][P0]);      // lap(p, q-1)
    const double lap_r = lap5(X[4][P0], X[4][P1], X[4][M1], X[5][P0], X[3][P0]);      // lap(p, q+2)
    const double flx_cA = limit(lap_pA - lapA, X[2][P1] - X[2][P0]);
    const double flx_cB = limit(lap_pB - lapB, X[3][P1] - X[3][P0]);
    const double fly_l = limit(lapA - lap_l, X[2][P0] - X[1][P0]);                    // between (p, q-1) and (p, q)
    const double fly_ab = limit(lapB - lapA, X[3][P0] - X[2][P0]);                    // between A and B
    const double fly_r = limit(lap_r - lapB, X[4][P0] - X[3][P0]);                    // between (p, q+1) and (p, q+2)
    if (actA && r - 4 >= i0) {
        double *o = p.out + ((long long)(r - 4) * p.J + j0) * K + ebase;
        __stcs(o, X[2][P0] - cfA * (((flx_cA - flxmA) + fly_ab) - fly_l));
        if (actB) __stcs(o + K, X[3][P0] - cfB * (((flx_cB - flxmB) + fly_r) - fly_ab));
    }
    lapA = lap_pA; lapB = lap_pB; flxmA = flx_cA; flxmB = flx_cB;
}

template <int NS>
__global__ void __launch_bounds__(THREADS, 1)
hdiff_ring_kernel(Params p) {
    extern __shared__ __align__(128) double ring_mem[];
    __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS];

    const int tid = threadIdx.x;
    const int K = p.K, I = p.I, J = p.J;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Work units = (i-block of RB rows, j-tile), j-tile fastest, dealt round-robin to the CTAs:
    // at any moment the CTAs work on neighbouring j-tiles of the same i-block, so the j-halo
    // columns and the i-halo rows that two units share are fetched from DRAM once and hit L2.
    const int n_units = p.n_iblocks * p.n_jtiles;
    unsigned job = 0;     // ring position, identical sequence in producer and consumers

    if (tid >= CONSUMERS) {
        // ------------------------------ producer (one elected thread) -------------
        if (tid == CONSUMERS) {
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int ib = u / p.n_jtiles, jt = u - ib * p.n_jtiles;
                const int i0 = ib * p.RB, i1 = min(I, i0 + p.RB);
                const int j0 = jt * p.TJ, tjc = min(p.TJ, J - j0);
                const unsigned bytes_in = (unsigned)((tjc + 4) * K) * 8u, bytes_c = (unsigned)(tjc * K) * 8u;
                for (int r = i0; r < i1 + 4; ++r, ++job) {
                    const int slot = job & (NS - 1);
                    const unsigned ph = (job / NS) & 1u;
                    mbar_wait(&empty_bar[slot], ph ^ 1u);           // slot drained by all consumer warps
                    const bool has_c = (r - 4 >= i0);
                    double *dst = ring_mem + (size_t)slot * p.slot_elems;
                    mbar_arrive_expect_tx(&full_bar[slot], bytes_in + (has_c ? bytes_c : 0u));
                    bulk_g2s(dst, p.in + ((long long)r * (J + 4) + j0) * K, bytes_in, &full_bar[slot]);
                    if (has_c)
                        bulk_g2s(dst + p.slot_in_elems, p.coeff + ((long long)(r - 4) * J + j0) * K, bytes_c,
                                 &full_bar[slot]);
                }
            }
        }
        return;
    }

    // ---------------------------------- consumers ---------------------------------
    const int pj = tid / K, kk = tid - pj * K;    // pair of output columns (2*pj, 2*pj+1) of the tile, level kk
    const int ebase = 2 * pj * K + kk;            // flattened (column, k) of cell A inside the tile's output row
    const int lane = tid & 31;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int ib = u / p.n_jtiles, jt = u - ib * p.n_jtiles;
        const int i0 = ib * p.RB, i1 = min(I, i0 + p.RB);
        const int j0 = jt * p.TJ, tjc = min(p.TJ, J - j0);
        const bool actA = 2 * pj < tjc, actB = 2 * pj + 1 < tjc;
        double X[6][4];
#pragma unroll
        for (int d = 0; d < 6; ++d)
#pragma unroll
            for (int q = 0; q < 4; ++q) X[d][q] = 0.0;
        double lapA = 0.0, lapB = 0.0, flxmA = 0.0, flxmB = 0.0;
        const int r_end = i1 + 4;
#define NPB_HD_STEP(WPOS)                                                                              \
        if (r < r_end) {                                                                               \
            const int slot = job & (NS - 1);                                                           \
            mbar_wait(&full_bar[slot], (job / NS) & 1u);                                               \
            consume_row<WPOS>(p, ring_mem + (size_t)slot * p.slot_elems, r, i0, j0, ebase, actA, actB, \
                              X, lapA, lapB, flxmA, flxmB, &empty_bar[slot], lane);                    \
            ++r; ++job;                                                                                \
        }
        for (int r = i0; r < r_end;) {
            NPB_HD_STEP(0) NPB_HD_STEP(1) NPB_HD_STEP(2) NPB_HD_STEP(3)
        }
#undef NPB_HD_STEP
    }
}

}  // namespace ring

int g_hdiff_mode = 0;        // 0 dispatch, 1 force marching kernel, 2 force ring kernel (if legal)
int g_hdiff_last = 0;        // 1 marching, 2 ring

// Returns 1 if the ring kernel was launched, 0 if not applicable.
int try_ring(int64_t I, int64_t J, int64_t K, const double *in, double *out, const double *coeff, bool force) {
    if ((K & 1) || K > ring::CONSUMERS || K < 2) return 0;           // 16-byte aligned rows; one row tile per CTA
    // cp.async.bulk needs 16-byte aligned global addresses (a misaligned one is a sticky fault); K even
    // only guarantees the row stride, so a view offset by an odd number of doubles takes the marching kernel
    if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)coeff) & 15) return 0;
    int TJ = 2 * (int)(ring::CONSUMERS / K);                       // column pairs x K threads <= CONSUMERS
    if (TJ > J) TJ = (int)((J + 1) & ~1LL);
    if (TJ < 2) return 0;
    const int n_jtiles = (int)((J + TJ - 1) / TJ);
    const int sms = npb::st().sm_count;
    if (!force && (long long)n_jtiles * I < 16LL * sms) return 0;     // too little work per SM
    const int slot_in = (int)((TJ + 4) * K);
    int slot_elems = slot_in + (int)(TJ * K);
    slot_elems = (slot_elems + 15) & ~15;
    const size_t slot_bytes = (size_t)slot_elems * sizeof(double);
    const int fit = (int)((npb::st().smem_optin - 1024) / slot_bytes);
    const int n_slots = fit >= 8 ? 8 : (fit >= 4 ? 4 : 0);
    if (n_slots == 0) return 0;
    // rows per i-block: the busiest CTA gets `rounds` units of the round-robin deal, each RB rows plus a 4-row ramp --
    // minimise that; ties go to the fuller last round (more SMs busy), then to the longer block.  (Cutting the rows
    // into equal contiguous j-tile-major ranges instead balances perfectly but was measured 12 % SLOWER at `paper`:
    // CTAs on neighbouring j-tiles no longer walk the same rows at the same time, so the shared halo columns miss L2.)
    int RB = 8;
    {
        long best_rows = -1, best_units = 0;
        const int rb_env = getenv("NPB_HDIFF_RB") ? atoi(getenv("NPB_HDIFF_RB")) : 0;
        for (int rb = 8; rb <= 96; ++rb) {
            const long units = (long)((I + rb - 1) / rb) * n_jtiles;
            const long rounds = (units + sms - 1) / sms;
            const long rows = rounds * (rb + 4);
            if (best_rows < 0 || rows < best_rows || (rows == best_rows && units >= best_units)) {
                best_rows = rows; best_units = units; RB = rb;
            }
        }
        if (rb_env > 0) RB = rb_env;
        if (RB > I) RB = (int)I;
    }
    const int n_iblocks = (int)((I + RB - 1) / RB);
    const size_t smem = slot_bytes * n_slots;
    ring::Params rp{(int)I, (int)J, (int)K, TJ, n_jtiles, RB, n_iblocks, slot_in, slot_elems, in, out, coeff};
    const long long units = (long long)n_iblocks * n_jtiles;
    const int grid = (int)(units < sms ? units : sms);
    // function attributes are per device: one cache per device (npb_mg_select)
    static size_t cfg8[NPB_MAX_DEVICES] = {0}, cfg4[NPB_MAX_DEVICES] = {0};
    size_t &configured8 = cfg8[npb::cur_device()], &configured4 = cfg4[npb::cur_device()];
    if (n_slots == 8) {
        if (smem > configured8) {
            if (cudaFuncSetAttribute(ring::hdiff_ring_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess) { cudaGetLastError(); return 0; }
            configured8 = smem;
        }
        ring::hdiff_ring_kernel<8><<<grid, ring::THREADS, smem, npb::st().stream>>>(rp);
    } else {
        if (smem > configured4) {
            if (cudaFuncSetAttribute(ring::hdiff_ring_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess) { cudaGetLastError(); return 0; }
            configured4 = smem;
        }
        ring::hdiff_ring_kernel<4><<<grid, ring::THREADS, smem, npb::st().stream>>>(rp);
    }
    return 1;
}

}  // namespace

// 0: dispatch by size; 1: always the marching kernel; 2: ring kernel whenever it is legal
extern "C" int npb_hdiff_set_mode(int mode) { g_hdiff_mode = mode; return 0; }
extern "C" int npb_hdiff_last_path(void) { return g_hdiff_last; }

extern "C" int npb_hdiff_f64(int64_t I, int64_t J, int64_t K, const double *in_field,
                             double *out_field, const double *coeff) {
    NPB_REQUIRE_INIT();
    NPB_ARG(I >= 0 && J >= 0 && K >= 0, "npb_hdiff_f64", "negative extent");
    if (I == 0 || J == 0 || K == 0) return 0;
    NPB_ARG(J * K < (1LL << 31) && I < (1LL << 31), "npb_hdiff_f64", "plane too large for 32-bit column index");
    if (g_hdiff_mode != 1) {
        if (try_ring(I, J, K, in_field, out_field, coeff, g_hdiff_mode == 2)) {
            NPB_CHECK_LAUNCH("hdiff_ring_kernel");
            npb::count_launch();
            g_hdiff_last = 2;
            return 0;
        }
    }
    g_hdiff_last = 1;
    const int JK = (int)(J * K);
    // enough chunks along i to fill the machine a few times over, but long
    // enough marches to amortise the 13-load window prologue
    const int col_blocks = (JK + HD_THREADS - 1) / HD_THREADS;
    int rows = 16;
    while (rows > 4 && (long long)col_blocks * ((I + rows - 1) / rows) < 4LL * npb::st().sm_count) rows >>= 1;
    dim3 grid(col_blocks, (unsigned)((I + rows - 1) / rows));
    hdiff_march_kernel<<<grid, HD_THREADS, 0, npb::st().stream>>>(
        (int)I, JK, (int)K, (long long)(J + 4) * K, in_field, out_field, coeff, rows);
    NPB_CHECK_LAUNCH("hdiff_march_kernel");
    npb::count_launch();
    return 0;
}
