// Round-1 shared-memory resident jacobi_2d kernel (T-deep halos, tile in shared memory, per-cell inbox sends).
// Measured slower than the blocked passes at S / M / L; superseded by jacobi2d_regtile.cuh.  Not built.
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

