// NOT PART OF THE BUILD.  Two heat_3d variants that were measured slower than the shipped kernels in round 1
// (numbers in DESIGN.md section 7) and are kept for reference only: heat3d_resident2_kernel (two sweeps per halo
// exchange, L 0.49 ms vs 0.37 ms) and heat3d_tb_kernel (temporally blocked shared-memory passes).  They were
// cut out of csrc/heat3d.cu verbatim; to experiment, include this file there after heat3d_resident_kernel.

// ---------------------------------------------------------------------------
// Resident variant, two sweeps per halo exchange (opt-in, mode 4; measured: L 0.49 ms vs 0.37 ms for
// the one-sweep kernel -- the redundant ring update and the 2.7x larger halo cost more than the saved
// round trips at these tile sizes).
//
// Same machinery as heat3d_resident_kernel (tiles resident in shared memory,
// sentinel-armed L2 inboxes, periodic fence), but the halo ring is TWO cells deep
// and is exchanged every second sweep: sweep 2n+1 updates the tile plus a one-cell
// ring of neighbours' cells redundantly, sweep 2n+2 updates the tile and pushes its
// two outermost layers (faces and 2x2 corners: 8 neighbours) to the inboxes.  The
// per-sweep cost of the resident kernel is one store->L2->poll round trip (~1.3 us)
// plus ~0.5 us of update; halving the round trips is worth the ~40% redundant work.
// Cell lists are per-ROW tables in shared memory (a row = all k of one (i,j)): send
// targets depend on (i,j) only.  2*(TSTEPS-1) is even, so the last pair writes state
// S-1 (own cells of its first sweep) to B and state S to A, like the reference.
// ---------------------------------------------------------------------------
constexpr int H2_SLOTS = 8;          // inbox ring depth, in exchanges
constexpr int H2_FENCE_EVERY = 2;    // gpu-scope fence cadence, in exchanges (2*cadence + 2 <= H2_SLOTS)
constexpr int H2_RECV = 10;          // inbox cells requested per thread before the first test
constexpr int H2_MAXROWS = 256;      // rows of the two-deep region (tile + halo)

struct Resident2Params {
    int n0, n1, n2;
    int PI, PJ, ti_max, tj_max;
    int npairs;                  // nsweeps / 2
    double *A, *B;
    unsigned long long *inbox;   // [PI*PJ][H2_SLOTS][(ti_max+4)*(tj_max+4)][n2-2]
    int fences;
    unsigned backoff_ns;
};

struct H2Row {                   // one (i,j) row of the CTA's region
    int soff;                    // shared offset of (row, k = 1)
    int goff;                    // global offset of (row, k = 1), or -1 if the row is not an own cell
    int tgt[8];                  // inbox cell offsets (without slot) of the receivers (up to 8 for
                                 // tiles only 2-3 cells wide), -1 = none
};

__global__ void __launch_bounds__(HR_THREADS, 1)
heat3d_resident2_kernel(Resident2Params p) {
    extern __shared__ double sm[];
    __shared__ H2Row rows1[H2_MAXROWS];          // sweep 2n+1: tile + one ring (clipped to the interior)
    __shared__ H2Row rows2[H2_MAXROWS];          // sweep 2n+2: own tile
    __shared__ int halo_rows[H2_MAXROWS];        // ring-linear index of every received row
    __shared__ int n_rows1, n_rows2, n_halo_rows;

    const int tid = threadIdx.x;
    const int ti = blockIdx.x / p.PJ, tj = blockIdx.x % p.PJ;
    int ilo, ihi, jlo, jhi;
    tile_bounds(p.n0 - 2, p.PI, ti, ilo, ihi);
    tile_bounds(p.n1 - 2, p.PJ, tj, jlo, jhi);
    const int nit = ihi - ilo, njt = jhi - jlo;
    const int n2 = p.n2, nk = n2 - 2;
    const int W = p.tj_max + 4;                          // region rows per i-plane (shared and inbox layout)
    const int rs = n2, ps = W * rs;                      // shared strides
    const size_t bufsz = (size_t)(p.ti_max + 4) * ps;
    double *const buf0 = sm, *const buf1 = sm + bufsz;   // even (A-parity) / odd (B-parity) states
    const long long grs = n2, gps = (long long)p.n1 * n2;
    const size_t slot_sz = (size_t)(p.ti_max + 4) * W * nk, box_sz = (size_t)H2_SLOTS * slot_sz;
    unsigned long long *my_box = p.inbox + (size_t)blockIdx.x * box_sz;
    // region (ri, rj) <-> global (ilo - 2 + ri, jlo - 2 + rj); my own tile is ri in [2, 2+nit), rj in [2, 2+njt)

    for (size_t w = tid; w < box_sz; w += HR_THREADS) my_box[w] = HR_SENTINEL;
    // initial state: the whole two-deep region of A -> buf0, of B -> buf1 (constant borders of each parity)
    for (int w = tid; w < (nit + 4) * (njt + 4) * n2; w += HR_THREADS) {
        const int r = w / n2, k = w - r * n2;
        const int ri = r / (njt + 4), rj = r - ri * (njt + 4);
        const int gi = ilo - 2 + ri, gj = jlo - 2 + rj;
        if (gi < 0 || gi >= p.n0 || gj < 0 || gj >= p.n1) continue;
        const long long g = (long long)gi * gps + (long long)gj * grs + k;
        buf0[ri * ps + rj * rs + k] = __ldg(p.A + g);
        buf1[ri * ps + rj * rs + k] = __ldg(p.B + g);
    }
    if (tid == 0) {
        int c1 = 0, c2 = 0, ch = 0;
        for (int ri = 0; ri < nit + 4; ++ri)
            for (int rj = 0; rj < njt + 4; ++rj) {
                const int gi = ilo - 2 + ri, gj = jlo - 2 + rj;
                if (gi < 1 || gi > p.n0 - 2 || gj < 1 || gj > p.n1 - 2) continue;     // interior cells only
                const bool own = (ri >= 2 && ri < 2 + nit && rj >= 2 && rj < 2 + njt);
                const int soff = ri * ps + rj * rs + 1;
                const int goff = (int)((long long)gi * gps + (long long)gj * grs + 1);
                if (!own) halo_rows[ch++] = ri * W + rj;                                // received every exchange
                if (ri >= 1 && ri < nit + 3 && rj >= 1 && rj < njt + 3) {                // sweep 2n+1
                    H2Row e; e.soff = soff; e.goff = own ? goff : -1;
                    for (int q = 0; q < 8; ++q) e.tgt[q] = -1;
                    rows1[c1++] = e;
                }
                if (own) {                                                               // sweep 2n+2 (+ sends)
                    H2Row e; e.soff = soff; e.goff = goff;
                    for (int q = 0; q < 8; ++q) e.tgt[q] = -1;
                    int nt = 0;
                    for (int di = -1; di <= 1; ++di)
                        for (int dj = -1; dj <= 1; ++dj) {
                            if (di == 0 && dj == 0) continue;
                            const int ti2 = ti + di, tj2 = tj + dj;
                            if (ti2 < 0 || ti2 >= p.PI || tj2 < 0 || tj2 >= p.PJ) continue;
                            int ilo2, ihi2, jlo2, jhi2;
                            tile_bounds(p.n0 - 2, p.PI, ti2, ilo2, ihi2);
                            tile_bounds(p.n1 - 2, p.PJ, tj2, jlo2, jhi2);
                            // inside that neighbour's two-deep region?
                            if (gi >= ilo2 - 2 && gi < ihi2 + 2 && gj >= jlo2 - 2 && gj < jhi2 + 2)
                                e.tgt[nt++] = (int)((size_t)(ti2 * p.PJ + tj2) * box_sz +
                                                    (size_t)((gi - (ilo2 - 2)) * W + (gj - (jlo2 - 2))) * nk);
                        }
                    rows2[c2++] = e;
                }
            }
        n_rows1 = c1; n_rows2 = c2; n_halo_rows = ch;
    }
    __threadfence();
    cooperative_groups::this_grid().sync();              // every inbox is armed, tables are built

    // receive descriptors of this thread (exchange invariant)
    const int nhalo = n_halo_rows * nk;
    int roff[H2_RECV], rdst[H2_RECV];
    unsigned rmask = 0;
#pragma unroll
    for (int u = 0; u < H2_RECV; ++u) {
        const int w = u * HR_THREADS + tid;
        roff[u] = 0; rdst[u] = 0;
        if (w < nhalo) {
            const int r = w / nk, kk = w - r * nk;
            const int lin = halo_rows[r];
            roff[u] = lin * nk + kk;
            rdst[u] = (lin / W) * ps + (lin % W) * rs + 1 + kk;
            rmask |= 1u << u;
        }
    }
    const int nc1 = n_rows1 * nk, nc2 = n_rows2 * nk;

    for (int pr = 0; pr < p.npairs; ++pr) {
        if (pr > 0) {
            // ---- receive the neighbours' state 2*pr: spin on the inbox cells themselves
            unsigned long long *slot = my_box + (size_t)(pr % H2_SLOTS) * slot_sz;
            unsigned pending = rmask;
            while (pending) {
                unsigned long long v[H2_RECV];
#pragma unroll
                for (int u = 0; u < H2_RECV; ++u)
                    if (pending & (1u << u)) v[u] = ld_relaxed_u64(slot + roff[u]);
#pragma unroll
                for (int u = 0; u < H2_RECV; ++u)
                    if ((pending & (1u << u)) && v[u] != HR_SENTINEL) {
                        buf0[rdst[u]] = __longlong_as_double((long long)v[u]);
                        st_relaxed_u64(slot + roff[u], HR_SENTINEL);       // re-arm for exchange pr + H2_SLOTS
                        pending &= ~(1u << u);
                    }
                if (pending) __nanosleep(p.backoff_ns);
            }
        }
        __syncthreads();
        // one fence every H2_FENCE_EVERY exchanges keeps "re-arm before the neighbour's rewrite" a
        // happens-before chain (see heat3d_resident_kernel)
        if (p.fences && (pr % H2_FENCE_EVERY) == 0) __threadfence();
        const bool last = (pr == p.npairs - 1);
        // ---- sweep 2*pr+1 : buf0 (even state) -> buf1, tile + one ring
        for (int w = tid; w < nc1; w += HR_THREADS) {
            const int r = w / nk, kk = w - r * nk;
            const H2Row &e = rows1[r];
            const double *c = buf0 + e.soff + kk;
            const double ce = c[0];
            const double c2 = 2.0 * ce;
            const double t1 = 0.125 * ((c[ps] - c2) + c[-ps]);
            const double t2 = 0.125 * ((c[rs] - c2) + c[-rs]);
            const double t3 = 0.125 * ((c[1] - c2) + c[-1]);
            const double v = ((t1 + t2) + t3) + ce;
            buf1[e.soff + kk] = v;
            if (last && e.goff >= 0) p.B[e.goff + kk] = v;          // state S-1
        }
        __syncthreads();
        // ---- sweep 2*pr+2 : buf1 -> buf0, own tile; outer layers go to the neighbours' inboxes
        unsigned long long *out_base = p.inbox + (size_t)((pr + 1) % H2_SLOTS) * slot_sz;
        for (int w = tid; w < nc2; w += HR_THREADS) {
            const int r = w / nk, kk = w - r * nk;
            const H2Row &e = rows2[r];
            const double *c = buf1 + e.soff + kk;
            const double ce = c[0];
            const double c2 = 2.0 * ce;
            const double t1 = 0.125 * ((c[ps] - c2) + c[-ps]);
            const double t2 = 0.125 * ((c[rs] - c2) + c[-rs]);
            const double t3 = 0.125 * ((c[1] - c2) + c[-1]);
            const double v = ((t1 + t2) + t3) + ce;
            buf0[e.soff + kk] = v;
            if (!last) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int t = e.tgt[q];
                    if (t < 0) break;
                    st_relaxed_f64((double *)(out_base + t + kk), v);
                }
            } else {
                p.A[e.goff + kk] = v;                                // state S
            }
        }
        // next iteration: receive writes halo cells of buf0 only (no update touches them); the barrier
        // after it orders this sweep's buf0 writes before the next sweep 2n+1 reads
    }
}


// Returns 1 if the two-sweep resident kernel ran, 0 if the problem is not eligible.
int try_resident2(int64_t nsweeps, int64_t n0, int64_t n1, int64_t n2, double *A, double *B) {
    if (nsweeps < 4 || (nsweeps & 1) || n0 > 4096 || n1 > 4096 || n2 > 4096) return 0;
    if (n0 * n1 * n2 >= (1LL << 31)) return 0;
    const int sms = npb::st().sm_count;
    const int in0 = (int)n0 - 2, in1 = (int)n1 - 2;
    int PI = 1, PJ = 1;
    {
        long best = -1;
        const int a_max = in0 >= 2 ? in0 / 2 : 1, b_max = in1 >= 2 ? in1 / 2 : 1;   // partitioned axes: tiles >= 2 wide
        for (int a = 1; a <= a_max && a <= sms; ++a) {
            int b = sms / a;
            if (b > b_max) b = b_max;
            if (b < 1) continue;
            const int ta = (in0 + a - 1) / a, tb = (in1 + b - 1) / b;
            // two sweeps: (ta+2)(tb+2) + ta*tb rows of update, (ta+4)(tb+4) - ta*tb rows of halo traffic
            const long cost = (long)(ta + 2) * (tb + 2) + (long)ta * tb + 2L * ((ta + 4) * (tb + 4) - ta * tb);
            if (best < 0 || cost < best) { best = cost; PI = a; PJ = b; }
        }
    }
    const int ti_max = (in0 + PI - 1) / PI, tj_max = (in1 + PJ - 1) / PJ;
    const int nk = (int)n2 - 2;
    if ((ti_max + 4) * (tj_max + 4) > H2_MAXROWS) return 0;
    if (((ti_max + 4) * (tj_max + 4) - ti_max * tj_max) * nk > HR_THREADS * H2_RECV) return 0;
    const size_t smem = (size_t)2 * (ti_max + 4) * (tj_max + 4) * n2 * sizeof(double);
    if (smem + 24576 > npb::st().smem_optin) return 0;          // + static row tables
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(heat3d_resident2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        configured = smem;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, heat3d_resident2_kernel, HR_THREADS, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if ((long)per_sm * sms < (long)PI * PJ) return 0;
    const size_t box = (size_t)H2_SLOTS * (ti_max + 4) * (tj_max + 4) * nk;
    if (box * PI * PJ >= (1ULL << 31)) return 0;                  // inbox offsets are 32-bit
    unsigned long long *inbox = (unsigned long long *)npb::workspace(1, box * PI * PJ * sizeof(unsigned long long));
    if (!inbox) return 0;
    Resident2Params rp{(int)n0, (int)n1, (int)n2, PI, PJ, ti_max, tj_max, (int)(nsweeps / 2), A, B, inbox,
                       g_resident_fences, g_backoff_ns};
    void *args[] = {&rp};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)heat3d_resident2_kernel, dim3(PI * PJ), dim3(HR_THREADS),
                                                args, smem, npb::st().stream);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    npb::count_launch();
    return 1;
}

// ---------------------------------------------------------------------------
// Temporally blocked variant for small (L2-resident) grids: NPBench S / M / L.
//
// At 70^3 a sweep is ~2 us of latency and a launch costs ~4 us, so the time loop
// is bound by the NUMBER of launches.  One launch of this kernel advances the
// grid by `nsteps` (1 or 3) sweeps: a CTA loads a TI x TJ tile of (i,j) columns
// (all k) plus an nsteps-deep halo ring into shared memory, runs the sweeps on a
// shrinking region between two shared buffers, and writes the tile centre.  As
// in jacobi2d.cu nsteps is odd, so a pass always goes A -> B or B -> A and the
// constant borders of each state's parity come from the right array.  Inside a
// sweep a thread owns one (j,k) column of the region and marches along i with a
// three-plane register window: 5 shared loads + 1 store per cell update.
// ---------------------------------------------------------------------------
constexpr int TB_H = 3;              // halo depth = max sweeps per launch
constexpr int TB_THREADS = 512;

struct TbParams {
    int n0, n1, n2;
    int T;                // tile edge (TI == TJ == T)
    int tiles_j;
    int nsteps;
    const double *src;
    double *dst;
};

__global__ void __launch_bounds__(TB_THREADS)
heat3d_tb_kernel(TbParams p) {
    extern __shared__ double sm[];
    const int n2 = p.n2, nk = n2 - 2;
    const int R = p.T + 2 * TB_H;                 // region edge (rows == cols)
    const int ps = R * n2;                        // shared stride between i-planes of the region
    double *buf0 = sm, *buf1 = sm + (size_t)R * ps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = blockIdx.x / p.tiles_j, tj = blockIdx.x % p.tiles_j;
    const int i0 = 1 + ti * p.T, j0 = 1 + tj * p.T;       // first output (i, j)
    const int gi_base = i0 - TB_H, gj_base = j0 - TB_H;    // region (r, c) <-> global (gi_base + r, gj_base + c)
    const int h = p.nsteps;
    const long long grs = n2, gps = (long long)p.n1 * n2;

    // ---- load rows (r, c) of the region that lie within h of the tile, clipped to the grid
    const int li_lo = max(0, i0 - h), li_hi = min(p.n0 - 1, i0 + p.T - 1 + h);
    const int lj_lo = max(0, j0 - h), lj_hi = min(p.n1 - 1, j0 + p.T - 1 + h);
    for (int row = warp; row < R * R; row += TB_THREADS / 32) {
        const int r = row / R, c = row - r * R;
        const int gi = gi_base + r, gj = gj_base + c;
        if (gi < li_lo || gi > li_hi || gj < lj_lo || gj > lj_hi) continue;
        const double *gs = p.src + gi * gps + gj * grs;
        const double *gd = p.dst + gi * gps + gj * grs;
        double *s0 = buf0 + r * ps + c * n2, *s1 = buf1 + r * ps + c * n2;
        const bool border = (gi == 0 || gi == p.n0 - 1 || gj == 0 || gj == p.n1 - 1);
        for (int k = lane; k < n2; k += 32) {
            s0[k] = __ldg(gs + k);
            // states of dst's parity: constant border rows and the two constant k-border cells
            if (border || k == 0 || k == n2 - 1) s1[k] = __ldg(gd + k);
        }
    }
    __syncthreads();

    // ---- nsteps sweeps on a shrinking region
    for (int s = 1; s <= p.nsteps; ++s) {
        const double *in = (s & 1) ? buf0 : buf1;
        double *out = (s & 1) ? buf1 : buf0;
        const int ui_lo = max(1, i0 - h + s), ui_hi = min(p.n0 - 2, i0 + p.T - 1 + h - s);
        const int uj_lo = max(1, j0 - h + s), uj_hi = min(p.n1 - 2, j0 + p.T - 1 + h - s);
        const int r_lo = ui_lo - gi_base, r_hi = ui_hi - gi_base;
        const int c_lo = uj_lo - gj_base, ncols = (uj_hi - uj_lo + 1) * nk;
        if (r_hi >= r_lo) {
            for (int col = tid; col < ncols; col += TB_THREADS) {
                const int cj = col / nk;
                const int k = 1 + (col - cj * nk);
                const double *q = in + r_lo * ps + (c_lo + cj) * n2 + k;
                double *o = out + r_lo * ps + (c_lo + cj) * n2 + k;
                double up = q[-ps], ce = q[0];
                for (int r = r_lo; r <= r_hi; ++r) {
                    const double dn = q[ps];
                    const double jm = q[-n2], jp = q[n2];
                    const double km = q[-1], kp = q[1];
                    const double c2 = 2.0 * ce;
                    const double t1 = 0.125 * ((dn - c2) + up);
                    const double t2 = 0.125 * ((jp - c2) + jm);
                    const double t3 = 0.125 * ((kp - c2) + km);
                    *o = ((t1 + t2) + t3) + ce;
                    up = ce; ce = dn;
                    q += ps; o += ps;
                }
            }
        }
        __syncthreads();
    }

    // ---- store the tile centre (interior cells) from the last buffer
    const double *fin = (p.nsteps & 1) ? buf1 : buf0;
    const int o_ihi = min(p.n0 - 2, i0 + p.T - 1), o_jhi = min(p.n1 - 2, j0 + p.T - 1);
    for (int row = warp; row < p.T * p.T; row += TB_THREADS / 32) {
        const int a = row / p.T, b = row - a * p.T;
        const int gi = i0 + a, gj = j0 + b;
        if (gi > o_ihi || gj > o_jhi) continue;
        const double *f = fin + (a + TB_H) * ps + (b + TB_H) * n2;
        double *g = p.dst + gi * gps + gj * grs;
        for (int k = 1 + lane; k <= nk; k += 32) g[k] = f[k];
    }
}

// Tile edge for the temporally blocked kernel, 0 if it should not be used.
int tb_pick_tile(int64_t n0, int64_t n1, int64_t n2, size_t *smem_out) {
    if (n0 * n1 * n2 > 600000 || n2 > 4096) return 0;       // large grids stream (one launch per sweep)
    const int sms = npb::st().sm_count;
    int best = 0;
    double best_cost = 0.0;
    for (int T = 8; T >= 2; --T) {
        const size_t smem = (size_t)2 * (T + 2 * TB_H) * (T + 2 * TB_H) * n2 * sizeof(double);
        if (smem + 2048 > npb::st().smem_optin) continue;
        long per_sm = (long)((npb::st().smem_optin + 1024) / (smem + 1024));
        if (per_sm > 2048 / TB_THREADS) per_sm = 2048 / TB_THREADS;
        const long tiles = (long)((n0 - 2 + T - 1) / T) * (long)((n1 - 2 + T - 1) / T);
        const long waves = (tiles + per_sm * sms - 1) / (per_sm * sms);
        double work = 0.0;
        for (int s = 1; s <= TB_H; ++s) work += (double)(T + 2 * TB_H - 2 * s) * (T + 2 * TB_H - 2 * s);
        const double cost = (double)waves * (work + 40.0);   // +40: fixed latency per launch wave
        if (best == 0 || cost < best_cost) { best = T; best_cost = cost; *smem_out = smem; }
    }
    return best;
}

int launch_tb(int T, size_t smem, int nsteps, int64_t n0, int64_t n1, int64_t n2, const double *src, double *dst) {
    static size_t configured = 0;
    if (smem > configured) {
        NPB_CUDA(cudaFuncSetAttribute(heat3d_tb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int tiles_i = (int)((n0 - 2 + T - 1) / T), tiles_j = (int)((n1 - 2 + T - 1) / T);
    TbParams p{(int)n0, (int)n1, (int)n2, T, tiles_j, nsteps, src, dst};
    heat3d_tb_kernel<<<tiles_i * tiles_j, TB_THREADS, smem, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("heat3d_tb_kernel");
    npb::count_launch();
    return 0;
}

// 2*(TSTEPS-1) sweeps as blocked passes (see jacobi2d.cu for the parity argument):
// an odd number of odd-sized passes A->B, B->A, ..., A->B, then one single sweep B->A.
int run_tb(int T, size_t smem, int64_t nsweeps, int64_t n0, int64_t n1, int64_t n2, double *A, double *B) {
    const int64_t M = nsweeps - 1;
    int64_t n = (M + TB_H - 1) / TB_H;
    if ((n & 1) == 0) ++n;
    int64_t extra_pairs = (M - n) / 2;
    const int64_t cap = (TB_H - 1) / 2;
    double *src = A, *dst = B;
    for (int64_t q = 0; q < n; ++q) {
        const int64_t left = n - q;
        int64_t take = (extra_pairs + left - 1) / left;
        if (take > cap) take = cap;
        extra_pairs -= take;
        const int rc = launch_tb(T, smem, (int)(1 + 2 * take), n0, n1, n2, src, dst);
        if (rc) return rc;
        double *t = src; src = dst; dst = t;
    }
    return launch_tb(T, smem, 1, n0, n1, n2, src, dst);
}

