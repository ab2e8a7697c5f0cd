// heat3d_regtile.cuh -- on-chip resident heat_3d for grids that fit in the SMs' registers + shared memory
// (NPBench presets S / M / L): ONE cooperative launch runs all sweeps of kernel(TSTEPS, A, B),
// npbench/benchmarks/polybench/heat_3d/heat_3d_numpy.py:4-20.
//
// Why this shape.  Round 1's resident kernel kept the tiles in shared memory and was bound by shared-memory
// traffic and instruction issue (ncu: l1tex 58 %, 9 shared accesses + a table entry per cell, FP64 pipe 13 %),
// plus 0.5 us per sweep of polling.  Measured on this machine (tools/pingpong.cu): one SM-to-SM signal through the L2
// (relaxed store -> relaxed load sees it) takes 0.43 us, a 16-byte atomic exchange far longer, a cluster barrier
// 0.45 us, grid.sync 1.2 us -- so a sweep that exchanges halos through the L2 cannot beat ~0.5 us, and everything
// else has to hide under that round trip.  Hence:
//
//   * The interior (i, j) plane is cut into PI x PJ tiles (one CTA = one SM each, all k).  A tile holds only
//     ~2400 cells, so its STATE LIVES IN REGISTERS: every thread owns a 2 x 2 x 2 block of cells for the whole
//     time loop.  Of a cell's six neighbours three are the thread's own registers; the other three come from
//     shared memory, where every thread publishes its eight new values once per sweep and reads the six faces of
//     its block -- 3.5 shared accesses per cell instead of 9, no index tables.  Lanes run along k; a column is stored
//     de-interleaved (odd cells, then even cells: see lds2), so the pairs AND the single k - 1 / k + 2 neighbours are
//     conflict-free 8-byte accesses.
//   * Halos travel through per-CTA inboxes in global memory (L2) with the sentinel protocol of inbox.cuh, but
//     the thread that needs a halo value polls it STRAIGHT INTO ITS REGISTERS and the thread that computed a
//     face value sends it straight from its registers: no staging copy, no second barrier -- one
//     __syncthreads per sweep.  The polls are issued at the top of the sweep.  (The intent was to run the
//     halo-independent part of the update -- 2c, the k term, the first differences: 40 of the 104 FP64 operations --
//     under their round trip.  ptxas does not schedule it that way, and forcing it measured slower: see face2 and
//     DESIGN.md section 7.)
//   * The loop is issue bound, so its control flow is flat: face loads, re-arms and sends are PREDICATED on a
//     per-thread role mask (no divergent branches).
//   * Inboxes stay armed between calls (every cell a sweep sends is consumed and re-armed in the next one,
//     the last sweep sends nothing), so only the first call on a geometry pays for arming and no grid-wide
//     barrier is needed; the launch is cooperative only to guarantee that all CTAs are co-resident.
//
// Odd extents: blocks are aligned to pairs counted from the first interior cell, so a block can stick out
// only over the GLOBAL border (never into a neighbour tile).  Such "invalid" cells shadow the constant border
// ring that both shared buffers carry (buf0: A's borders = even states, buf1: B's = odd states): after each
// sweep their registers are reloaded from the ring, so neighbours see exactly the border value.
//
// Arithmetic: NumPy order, one rounding per operation (-fmad=false), as in heat3d_sweep_kernel.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "inbox.cuh"

namespace regtile {

constexpr int SLOTS = 18;          // inbox ring depth (sweeps)
// gpu-scope fence cadence (sweeps); needs 2 * cadence + 1 <= SLOTS.  The hazard: the consumer re-arms a cell of slot
// (s - 1) % SLOTS in sweep s, the producer writes that cell again in its sweep s + SLOTS - 1; the re-arm must be
// ordered before that write.  Every thread fences at the END of the sweeps f = 0 mod FENCE_EVERY (behind that sweep's
// sends and re-arms).  Chain: re-arm(s) -> the consumer's fence at the end of sweep f <= s + FENCE_EVERY - 1 -> its
// sends of sweep f + 1 -> the producer's poll observes them in sweep f + 2 -> the producer's fence at the end of
// sweep f' = f + FENCE_EVERY -> its sends of every sweep >= f' + 1 <= s + 2 FENCE_EVERY <= s + SLOTS - 1.
// (Across a side the same two threads are each other's producer and consumer.)  Round 2 ran 16 slots with a
// cadence of 8: one sweep short of this argument.
constexpr int FENCE_EVERY = 8;

struct Params {
    int n0, n1, n2;
    int PI, PJ;                  // tiles along i and j
    int pi, pj, pk;              // cell pairs along i, j, k: ceil(interior / 2)
    int BI, BJ;                  // largest number of pairs per tile along i / j
    int rows;                    // column blocks per inbox side = max(BI, BJ)
    int nsweeps;
    double *A, *B;
    unsigned long long *inbox;   // [PI * PJ][SLOTS][4 sides][rows][pk][2 face rows][2 k]
    int flags;                   // timing experiments only (results are wrong with 2 or 4): 1 no fences, 2 no polls, 4 no sends
    long long *trace;            // TRACE instantiation only: [2 threads][5 phases] accumulated clock64 deltas of the centre CTA
};

__device__ __forceinline__ void pair_range(int npairs, int parts, int t, int &lo, int &cnt) {
    const int base = npairs / parts, rem = npairs % parts;
    lo = t * base + min(t, rem);
    cnt = base + (t < rem ? 1 : 0);
}

// One face value pair: from the inbox (L2) if bit BIT of `mask` is set, else from shared memory.  Predicated,
// not branched: a branch costs a BSSY / BRA / BSYNC triple per face.  Both loads write the SAME registers, so in a warp
// with halo lanes the (predicated-off) shared loads wait for the scoreboard of the inbox load in front of them: the L2
// round trip is exposed once per sweep (ncu: 12 % of the warp samples).  Giving the polls registers of their own hides
// it but costs 32 selects and a second set of predicates per sweep -- measured slower.
template <unsigned BIT>
__device__ __forceinline__ void face2(unsigned mask, const unsigned long long *g, unsigned sa, unsigned h8, double2 &v) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %4, %6;\n\tsetp.ne.b32 q, t, 0;\n\t"
                 "@q ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];\n\t"
                 "@!q ld.shared.f64 %0, [%3];\n\t@!q ld.shared.f64 %1, [%5];\n\t}"
                 : "=d"(v.x), "=d"(v.y) : "l"(g), "r"(sa), "r"(mask), "r"(sa + h8), "n"(BIT));
}
template <unsigned BIT>
__device__ __forceinline__ void repoll2(unsigned mask, const unsigned long long *g, double2 &v) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.b32 q, t, 0;\n\t"
                 "@q ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(v.x), "+d"(v.y) : "l"(g), "r"(mask), "n"(BIT));
}
template <unsigned BIT>
__device__ __forceinline__ void send2(unsigned mask, unsigned long long *g, double a, double b) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.b32 q, t, 0;\n\t"
                 "@q st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};\n\t}"
                 ::"l"(g), "d"(a), "d"(b), "r"(mask), "n"(BIT));
}
template <unsigned BIT>
__device__ __forceinline__ void arm2(unsigned mask, unsigned long long *g) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.b32 q, t, 0;\n\t"
                 "@q st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};\n\t}"
                 ::"l"(g), "l"(HR_SENTINEL), "r"(mask), "n"(BIT));
}
// arithmetic never produces a NaN whose high word is the sentinel's (its results are quiet NaNs), so the high
// word alone tells an armed cell from a delivered value; values read from shared memory always pass
__device__ __forceinline__ bool delivered(const double2 &v) {
    return __double2hiint(v.x) != (int)(HR_SENTINEL >> 32) && __double2hiint(v.y) != (int)(HR_SENTINEL >> 32);
}
// shared memory through 32-bit addresses (the extents are run-time values: explicit addresses keep the
// per-sweep integer work at one add per access instead of re-derived index products)
// A column is stored DE-INTERLEAVED: its odd cells k = 2b + 1 (the low halves of the threads' k pairs) at b + 1, its even
// cells k = 2b + 2 at HALF + b + 1 (HALF = pk + 2; k = 0 at HALF).  A pair is two 8-byte accesses H8 = 8 HALF bytes apart,
// consecutive lanes touch consecutive doubles: every access is conflict free (2 wavefronts per warp), including the
// k - 1 / k + 2 neighbours of a pair.  With interleaved pairs those two were 8-byte accesses at a 16-byte stride: 4
// wavefronts each, 20 % of the kernel's shared-memory cycles (ncu round 2: 6.4 M bank-conflict cycles).
__device__ __forceinline__ double2 lds2(unsigned a, unsigned h8) {     // the pair whose low half sits at a
    double2 v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v.x) : "r"(a));
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v.y) : "r"(a + h8));
    return v;
}
__device__ __forceinline__ double lds1_below(unsigned a, unsigned h8) {   // the cell at k - 1 of the pair at a: high half of the pair below
    double v;
    asm volatile("ld.shared.f64 %0, [%1+-8];" : "=d"(v) : "r"(a + h8));
    return v;
}
__device__ __forceinline__ double lds1_above(unsigned a) {          // the cell at k + 2: low half of the pair above
    double v;
    asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(unsigned a, unsigned h8, double x, double y) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a + h8), "d"(y) : "memory");
}

__global__ void heat3d_inbox_arm_kernel(unsigned long long *box, size_t n) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x)
        box[w] = HR_SENTINEL;
}

template <int MAXT, bool TRACE>
__global__ void __launch_bounds__(MAXT, 1) heat3d_regtile_kernel(Params p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    // TRACE: %globaltimer stamps (ns) of the centre CTA's thread 0 -> trace[10..15]: entry, state loaded, after sweep 1,
    // after sweep 17, loop left, results stored
    long long stamp[6] = {0, 0, 0, 0, 0, 0};
#define RT_STAMP(i) do { if (TRACE) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamp[i])); } while (0)
    RT_STAMP(0);
    const int ti = blockIdx.x / p.PJ, tj = blockIdx.x % p.PJ;
    int ip0, nip, jp0, njp;
    pair_range(p.pi, p.PI, ti, ip0, nip);
    pair_range(p.pj, p.PJ, tj, jp0, njp);
    const int ilo = 1 + 2 * ip0, jlo = 1 + 2 * jp0;                 // first interior cell of the tile
    const int nit = min(2 * nip, p.n0 - 1 - ilo), njt = min(2 * njp, p.n1 - 1 - jlo);
    const int n2 = p.n2, nk = n2 - 2, pk = p.pk;
    const int HALF = pk + 2, KS = 2 * HALF;                          // shared stride of one (i, j) column, de-interleaved (see lds2)
    const int CJ = 2 * p.BJ + 2;                                     // columns per i row (tile + ring)
    const int RS = CJ * KS;                                          // shared stride of one i row
    const int bufsz = (2 * p.BI + 2) * RS;
    double *const buf0 = sm, *const buf1 = sm + bufsz;               // even (A borders) / odd (B borders) states
    const long long grs = n2, gps = (long long)p.n1 * n2;

    // never-written cells of the buffers (corners of the ring, padding beyond odd extents) are read as don't-care
    // face values: give them a defined, non-sentinel content
    for (int w = tid; w < 2 * bufsz; w += blockDim.x) sm[w] = 0.0;
    __syncthreads();
    // ---- initial state: tile + one-cell ring of A -> buf0, of B -> buf1; cell (ii, jj, k) of the ring-inclusive
    //      region sits at (ii * CJ + jj) * KS + (k odd ? (k + 1) / 2 : HALF + k / 2)
    //      Asynchronous 8-byte copies (LDGSTS): every thread has all of its ~2 x 14 loads in flight at once.  A loop of
    //      load -> store pairs paid one cold DRAM round trip per iteration: ~20 us of a 285 us call at L.
    {
        const int rows = (nit + 2) * (njt + 2);
        const unsigned s0 = (unsigned)__cvta_generic_to_shared(buf0), s1 = (unsigned)__cvta_generic_to_shared(buf1);
        for (int w = tid; w < rows * n2; w += blockDim.x) {
            const int r = w / n2, k = w - r * n2;
            const int ii = r / (njt + 2), jj = r - ii * (njt + 2);
            const long long g = (long long)(ilo - 1 + ii) * gps + (long long)(jlo - 1 + jj) * grs + k;
            const unsigned l = (unsigned)((ii * CJ + jj) * KS + ((k & 1) ? (k + 1) / 2 : HALF + k / 2)) * 8u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + l), "l"(p.A + g) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s1 + l), "l"(p.B + g) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    RT_STAMP(1);

    // ---- this thread's block: cells (ilo + a0 + di, jlo + b0 + dj, 1 + 2 * bk + dk)
    const int nblk = nip * njp * pk;
    const bool active = tid < nblk;
    const int cb = tid / pk, bk = tid - cb * pk;
    const int bi = cb / njp, bj = cb - bi * njp;
    const int a0 = 2 * bi, b0 = 2 * bj;
    const bool vi1 = a0 + 1 < nit, vj1 = b0 + 1 < njt, vk1 = 2 * bk + 1 < nk;
    const bool all_valid = vi1 && vj1 && vk1;
    // halo sides of this thread (its block touches a tile edge that has a neighbour tile)
    const bool e_im = active && bi == 0 && ti > 0, e_ip = active && bi == nip - 1 && ti < p.PI - 1;
    const bool e_jm = active && bj == 0 && tj > 0, e_jp = active && bj == njp - 1 && tj < p.PJ - 1;
    // shared byte addresses (buffer 0) of the block's column (0, 0) at the low half of its k pair
    const unsigned KS8 = (unsigned)KS * 8u, RS8 = (unsigned)RS * 8u, BUF8 = (unsigned)bufsz * 8u, H8 = (unsigned)HALF * 8u;
    const unsigned a00 = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(((a0 + 1) * CJ + (b0 + 1)) * KS + 1 + bk) * 8u;

    // inbox word offsets (32 bit).  A side holds [column block along the edge][k pair][2 columns][2 k] values, so the
    // two face rows of a block are 32 contiguous bytes and the lanes of a warp touch consecutive groups.
    // receive: side 0 = from i-1, 1 = from i+1, 2 = from j-1, 3 = from j+1
    const unsigned side_sz = (unsigned)p.rows * pk * 4u, slot_sz = 4u * side_sz, box_sz = (unsigned)SLOTS * slot_sz;
    const unsigned my = blockIdx.x * box_sz;
    const unsigned r_im = 0u * side_sz + (unsigned)(bj * pk + bk) * 4u, r_ip = r_im + side_sz;
    const unsigned r_jm = 2u * side_sz + (unsigned)(bi * pk + bk) * 4u, r_jp = r_jm + side_sz;
    // send: my i-1 face lands in the i-1 neighbour's side 1, and so on
    const unsigned t_im = (blockIdx.x - p.PJ) * box_sz + r_ip, t_ip = (blockIdx.x + p.PJ) * box_sz + r_im;
    const unsigned t_jm = (blockIdx.x - 1) * box_sz + r_jp, t_jp = (blockIdx.x + 1) * box_sz + r_jm;
    unsigned long long *const box = p.inbox;
    // role bits of this thread: bit 2 * side = first face row of that side crosses to a neighbour tile, bit 2 * side + 1 =
    // second row does (the second row of a block can lie over the global border)
    const unsigned emask = (e_im ? (1u | (vj1 ? 2u : 0u)) : 0u) | (e_ip ? (4u | (vj1 ? 8u : 0u)) : 0u) |
                           (e_jm ? (16u | (vi1 ? 32u : 0u)) : 0u) | (e_jp ? (64u | (vi1 ? 128u : 0u)) : 0u);

    double o[2][2][2];                                               // own cells, state s - 1
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj) {
            const double2 v = active ? lds2(a00 + di * RS8 + dj * KS8, H8) : make_double2(0.0, 0.0);
            o[di][dj][0] = v.x; o[di][dj][1] = v.y;
        }

    unsigned in_off = 0, out_off = slot_sz;                          // slot (s - 1) % SLOTS and s % SLOTS, in words
    const bool polling = !(p.flags & 2), sending = !(p.flags & 4), fencing = !(p.flags & 1);
    // TRACE: thread 0 (a corner block: two halo sides) and the middle thread of the centre CTA accumulate the
    // cycles spent in [faces + halo-independent part | wait | rest + sends | re-arm + publish | fence + barrier]
    const bool tracer = TRACE && blockIdx.x == (unsigned)((p.PI / 2) * p.PJ + p.PJ / 2) && (tid == 0 || tid == nblk / 2);
    long long tr[5] = {0, 0, 0, 0, 0}, tc = 0;
    if (TRACE) tc = clock64();
#define RT_MARK(i) do { if (TRACE && tracer) { const long long t_ = clock64(); tr[i] += t_ - tc; tc = t_; } } while (0)
    for (int s = 1; s <= p.nsweeps; ++s) {
        if (active) {
            const unsigned co = (s & 1) ? 0u : BUF8;                 // state s - 1 lives in buffer (s - 1) & 1
            const unsigned c00 = a00 + co, c01 = c00 + KS8, c10 = c00 + RS8, c11 = c10 + KS8;
            const unsigned pmask = (s > 1 && polling) ? emask : 0u;  // state 0 halos came with the initial load
            unsigned long long *const q_im = box + (my + in_off + r_im), *const q_ip = box + (my + in_off + r_ip);
            unsigned long long *const q_jm = box + (my + in_off + r_jm), *const q_jp = box + (my + in_off + r_jp);
            // ---- the six faces of the block.  i / j faces come from shared memory or, across a tile edge, straight from
            //      the inbox
            double2 im[2], ip[2], jm[2], jp[2];                      // i faces: [dj], j faces: [di]; .x / .y = dk
            face2<1u>(pmask, q_im, c00 - RS8, H8, im[0]); face2<2u>(pmask, q_im + 2, c01 - RS8, H8, im[1]);
            face2<4u>(pmask, q_ip, c10 + RS8, H8, ip[0]); face2<8u>(pmask, q_ip + 2, c11 + RS8, H8, ip[1]);
            face2<16u>(pmask, q_jm, c00 - KS8, H8, jm[0]); face2<32u>(pmask, q_jm + 2, c10 - KS8, H8, jm[1]);
            face2<64u>(pmask, q_jp, c01 + KS8, H8, jp[0]); face2<128u>(pmask, q_jp + 2, c11 + KS8, H8, jp[1]);
            double km[2][2], kp[2][2];
            km[0][0] = lds1_below(c00, H8); kp[0][0] = lds1_above(c00);
            km[0][1] = lds1_below(c01, H8); kp[0][1] = lds1_above(c01);
            km[1][0] = lds1_below(c10, H8); kp[1][0] = lds1_above(c10);
            km[1][1] = lds1_below(c11, H8); kp[1][1] = lds1_above(c11);
            // ---- the part of the update (heat_3d_numpy.py:7-13 order) that needs no i / j face: 2c, the k term, and
            //      the first difference of the i / j terms whose "+" neighbour is one of the thread's own cells
            double c2[2][2][2], t3[2][2][2], e1[2][2], e2[2][2];     // e1[dj][dk] = o[1][dj][dk] - 2 o[0][dj][dk], e2 alike
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                    for (int dk = 0; dk < 2; ++dk) {
                        c2[di][dj][dk] = 2.0 * o[di][dj][dk];
                        const double zp = dk ? kp[di][dj] : o[di][dj][1];
                        const double zm = dk ? o[di][dj][0] : km[di][dj];
                        t3[di][dj][dk] = 0.125 * ((zp - c2[di][dj][dk]) + zm);
                    }
#pragma unroll
            for (int d = 0; d < 2; ++d)
#pragma unroll
                for (int dk = 0; dk < 2; ++dk) {
                    e1[d][dk] = o[1][d][dk] - c2[0][d][dk];
                    e2[d][dk] = o[d][1][dk] - c2[d][0][dk];
                }
            // (An empty asm with "+d" operands keeps NVVM from sinking this work below the wait, but leaves no trace in
            // the PTX: ptxas moves it there all the same.  Pinning it in front -- with polled values in registers of their
            // own so that the shared loads do not wait for the polls -- was measured 7 % SLOWER: DESIGN.md section 7.)
            asm volatile("" : "+d"(t3[0][0][0]), "+d"(t3[0][0][1]), "+d"(t3[0][1][0]), "+d"(t3[0][1][1]),
                              "+d"(t3[1][0][0]), "+d"(t3[1][0][1]), "+d"(t3[1][1][0]), "+d"(t3[1][1][1]));
            asm volatile("" : "+d"(e1[0][0]), "+d"(e1[0][1]), "+d"(e1[1][0]), "+d"(e1[1][1]),
                              "+d"(e2[0][0]), "+d"(e2[0][1]), "+d"(e2[1][0]), "+d"(e2[1][1]));
            RT_MARK(0);
            // ---- wait for the faces that were not there yet (values that came from shared memory always pass)
            while (!(delivered(im[0]) && delivered(im[1]) && delivered(ip[0]) && delivered(ip[1]) &&
                     delivered(jm[0]) && delivered(jm[1]) && delivered(jp[0]) && delivered(jp[1]))) {
                repoll2<1u>(pmask, q_im, im[0]); repoll2<2u>(pmask, q_im + 2, im[1]);
                repoll2<4u>(pmask, q_ip, ip[0]); repoll2<8u>(pmask, q_ip + 2, ip[1]);
                repoll2<16u>(pmask, q_jm, jm[0]); repoll2<32u>(pmask, q_jm + 2, jm[1]);
                repoll2<64u>(pmask, q_jp, jp[0]); repoll2<128u>(pmask, q_jp + 2, jp[1]);
            }
            RT_MARK(1);
            // ---- the rest of the update
            double v[2][2][2];
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                    for (int dk = 0; dk < 2; ++dk) {
                        const double xin = di ? (dk ? ip[dj].y : ip[dj].x) : (dk ? im[dj].y : im[dj].x);
                        const double yin = dj ? (dk ? jp[di].y : jp[di].x) : (dk ? jm[di].y : jm[di].x);
                        const double t1 = di ? 0.125 * ((xin - c2[1][dj][dk]) + o[0][dj][dk]) : 0.125 * (e1[dj][dk] + xin);
                        const double t2 = dj ? 0.125 * ((yin - c2[di][1][dk]) + o[di][0][dk]) : 0.125 * (e2[di][dk] + yin);
                        v[di][dj][dk] = ((t1 + t2) + t3[di][dj][dk]) + o[di][dj][dk];
                    }
            // ---- cells beyond the interior shadow the constant border of state s (ring of the other buffer)
            const unsigned n00 = a00 + (BUF8 - co);
            if (!all_valid) {
#pragma unroll
                for (int di = 0; di < 2; ++di)
#pragma unroll
                    for (int dj = 0; dj < 2; ++dj) {
                        const double2 b = lds2(n00 + di * RS8 + dj * KS8, H8);
                        if ((di && !vi1) || (dj && !vj1)) { v[di][dj][0] = b.x; v[di][dj][1] = b.y; }
                        else if (!vk1) v[di][dj][1] = b.y;
                    }
            }
            // ---- faces to the neighbour tiles, straight from the registers (nobody consumes the last state)
            {
                const unsigned smask = (s < p.nsweeps && sending) ? emask : 0u;
                unsigned long long *const u_im = box + (t_im + out_off), *const u_ip = box + (t_ip + out_off);
                unsigned long long *const u_jm = box + (t_jm + out_off), *const u_jp = box + (t_jp + out_off);
                send2<1u>(smask, u_im, v[0][0][0], v[0][0][1]); send2<2u>(smask, u_im + 2, v[0][1][0], v[0][1][1]);
                send2<4u>(smask, u_ip, v[1][0][0], v[1][0][1]); send2<8u>(smask, u_ip + 2, v[1][1][0], v[1][1][1]);
                send2<16u>(smask, u_jm, v[0][0][0], v[0][0][1]); send2<32u>(smask, u_jm + 2, v[1][0][0], v[1][0][1]);
                send2<64u>(smask, u_jp, v[0][1][0], v[0][1][1]); send2<128u>(smask, u_jp + 2, v[1][1][0], v[1][1][1]);
            }
            RT_MARK(2);
            // ---- re-arm the consumed inbox cells for sweep s - 1 + SLOTS (behind the sends: those are on the neighbours'
            //      critical path; re-arming right after the wait measured 5 % slower)
            arm2<1u>(pmask, q_im); arm2<2u>(pmask, q_im + 2);
            arm2<4u>(pmask, q_ip); arm2<8u>(pmask, q_ip + 2);
            arm2<16u>(pmask, q_jm); arm2<32u>(pmask, q_jm + 2);
            arm2<64u>(pmask, q_jp); arm2<128u>(pmask, q_jp + 2);
            // ---- publish state s to the tile's threads, keep it in registers
            sts2(n00, H8, v[0][0][0], v[0][0][1]);
            sts2(n00 + KS8, H8, v[0][1][0], v[0][1][1]);
            sts2(n00 + RS8, H8, v[1][0][0], v[1][0][1]);
            sts2(n00 + RS8 + KS8, H8, v[1][1][0], v[1][1][1]);
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj) { o[di][dj][0] = v[di][dj][0]; o[di][dj][1] = v[di][dj][1]; }
        }
        RT_MARK(3);
        in_off = out_off;
        out_off += slot_sz;
        if (out_off == box_sz) out_off = 0;
        // gpu-scope fence every FENCE_EVERY sweeps, at the end of the sweep: off the critical path of the sends (in
        // front of them it measured 5 % slower)
        if ((s % FENCE_EVERY) == 0 && fencing) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        __syncthreads();
        RT_MARK(4);
        if (TRACE && s == 1) RT_STAMP(2);
        if (TRACE && s == 17) RT_STAMP(3);
    }
    RT_STAMP(4);
    if (TRACE && tracer && p.trace)
        for (int i = 0; i < 5; ++i) p.trace[(tid == 0 ? 0 : 5) + i] = tr[i];
#undef RT_MARK
    // ---- the two buffers hold the last two states of the tile: state S (even) goes to A, state S - 1 to B, like
    //      the reference leaves them (borders are untouched)
    {
        const double *last = (p.nsweeps & 1) ? buf1 : buf0, *prev = (p.nsweeps & 1) ? buf0 : buf1;
        double *g_last = (p.nsweeps & 1) ? p.B : p.A, *g_prev = (p.nsweeps & 1) ? p.A : p.B;
        for (int w = tid; w < nit * njt * nk; w += blockDim.x) {
            const int r = w / nk, k = 1 + (w - r * nk);
            const int ii = r / njt, jj = r - ii * njt;
            const long long g = (long long)(ilo + ii) * gps + (long long)(jlo + jj) * grs + k;
            const int l = ((ii + 1) * CJ + (jj + 1)) * KS + ((k & 1) ? (k + 1) / 2 : HALF + k / 2);
            g_last[g] = last[l];
            if (p.nsweeps > 1) g_prev[g] = prev[l];
        }
    }
    if (TRACE) {
        __syncthreads();
        RT_STAMP(5);
        if (tracer && tid == 0 && p.trace)
            for (int i = 0; i < 6; ++i) p.trace[10 + i] = stamp[i];
        if (tid == 0 && p.trace) { p.trace[16 + 2 * blockIdx.x] = stamp[0]; p.trace[17 + 2 * blockIdx.x] = stamp[5]; }   // every CTA: entry, exit
    }
#undef RT_STAMP
}

}  // namespace regtile
