// heat_3d, three sweeps per launch for grids that live in HBM (included by heat3d.cu).
//
// The one-launch-per-sweep kernel moves 16 B per cell update and is HBM-bound
// (reference loop: heat_3d_numpy.py:6-19).  This kernel advances the grid by THREE
// sweeps per pass over memory: a CTA owns an 18 x 58 tile of (j, k) columns plus a
// 3-deep halo ring (a 24 x 64 region) and marches along i.  The three sweeps are
// pipelined one plane apart: when source plane s arrives, the CTA computes plane s-1
// of state 1, plane s-2 of state 2 and plane s-3 of state 3, which goes to global
// memory.  Each state keeps the two most recent planes of the region in shared
// memory (for the j/k neighbours); the i neighbours of a cell are the same thread's
// registers (centre, next) and the shared slot it is about to overwrite (previous).
// One __syncthreads per plane.  Region cells whose dependency cone leaves the region
// hold garbage that never reaches the 18 x 58 centre.
//
// Traffic per cell and three sweeps: 8 B * (24*64)/(18*58) read + 8 B written
// = 19.8 B, i.e. 6.6 B per cell update against 16 B.
//
// Borders: state q's constant border cells equal A's (q even) or B's (q odd) border,
// exactly as the one-sweep launches leave them, so a pass src -> dst takes state 1's
// border from dst and state 2's from src, and never writes a border cell.
#pragma once

constexpr int HM_TJ = 18, HM_TK = 58;            // output tile
constexpr int HM_RJ = HM_TJ + 6, HM_RK = HM_TK + 6;   // region = tile + 3-deep ring: 24 x 64
constexpr int HM_THREADS = 256;
constexpr int HM_CELLS = HM_RJ * HM_RK;          // 1536
constexpr int HM_CPT = HM_CELLS / HM_THREADS;    // 6 consecutive rows of one column per thread
constexpr int HM_PAD = HM_RK;                    // one spare row above and below EVERY plane: the garbage cells of
                                                 // region rows 0 / 23 read there, never in a plane another thread
                                                 // is writing in the same plane step (racecheck-clean)
constexpr int HM_PLANE = HM_CELLS + 2 * HM_PAD;  // shared stride between planes
constexpr size_t HM_SMEM = (size_t)(6 * HM_PLANE) * sizeof(double);
static_assert(HM_RK == 64 && HM_CPT * (HM_THREADS / HM_RK) == HM_RJ, "thread <-> cell mapping");

struct HmParams {
    int n0, n1, n2;
    int tiles_k;
    int chunk;            // output planes per CTA along i
    int i_lo, i_hi;       // output planes [i_lo, i_hi) of this launch (1 <= i_lo, i_hi <= n0 - 1)
    const double *src;
    double *dst;
};

__device__ __forceinline__ double hm_update(double up, double ce, double dn, double jm, double jp,
                                            double km, double kp) {
    const double c2 = 2.0 * ce;
    const double t1 = 0.125 * ((dn - c2) + up);      // heat_3d_numpy.py:7-8 (axis 0)
    const double t2 = 0.125 * ((jp - c2) + jm);      // :9-10 (axis 1)
    const double t3 = 0.125 * ((kp - c2) + km);      // :11-12 (axis 2)
    return ((t1 + t2) + t3) + ce;
}

// EDGE = the CTA's region touches the j/k faces of the grid (or leaves it): per-cell inside /
// border predicates.  Interior CTAs (most of them) run without any per-cell predicate.
template <bool EDGE>
__device__ __forceinline__ void hm_march(const HmParams &p, double *S, int tj, int tk) {
    const int tid = threadIdx.x, rk = tid & (HM_RK - 1), rb = tid >> 6;
    const int n0 = p.n0, n1 = p.n1, n2 = p.n2;
    const int gk = tk * HM_TK - 2 + rk;                  // region (r, rk) <-> global (gj0 + q, gk)
    const int gj0 = tj * HM_TJ - 2 + rb * HM_CPT;
    const int ia = p.i_lo + blockIdx.y * p.chunk, ib = min(ia + p.chunk, p.i_hi);   // output planes [ia, ib)
    const int L = max(0, ia - 3), E = min(n0, ib + 3);                         // source planes [L, E)
    const long long ps = (long long)n1 * n2;
    const long long off0 = (long long)gj0 * n2 + gk;
    const int c0 = rb * HM_CPT * HM_RK + rk;

    unsigned m_in = 0, m_bd = 0, m_out = 0;
    {
        const bool kin = gk >= 0 && gk < n2, kb = (gk == 0 || gk == n2 - 1);
        const bool kout = rk >= 3 && rk < 3 + HM_TK && gk <= n2 - 2;
#pragma unroll
        for (int q = 0; q < HM_CPT; ++q) {
            const int gj = gj0 + q, r = rb * HM_CPT + q;
            const bool in = kin && gj >= 0 && gj < n1;
            const bool bd = in && (kb || gj == 0 || gj == n1 - 1);
            const bool out = in && !bd && kout && r >= 3 && r < 3 + HM_TJ;
            m_in |= (unsigned)in << q; m_bd |= (unsigned)bd << q; m_out |= (unsigned)out << q;
        }
    }
    if (!EDGE) { m_in = (1u << HM_CPT) - 1; m_bd = 0; }

    double v0c[HM_CPT], v0p[HM_CPT], v1c[HM_CPT], v2c[HM_CPT], b1c[HM_CPT];
    const double *gs = p.src + L * ps + off0;        // source plane being prefetched
    const double *gd = p.dst + L * ps + off0;        // dst plane `step` (border values of state 1)
    double *go = p.dst + (L - 3) * ps + off0;        // output plane step-3
#pragma unroll
    for (int q = 0; q < HM_CPT; ++q) {
        v0c[q] = 0.0; v1c[q] = 0.0; v2c[q] = 0.0; b1c[q] = 0.0;
        v0p[q] = ((m_in >> q) & 1) ? __ldg(gs + q * n2) : 0.0;
    }
    gs += ps;

    for (int step = L; step < ib + 3; ++step) {
        // ---- prefetch source plane step+1 and, for border cells, dst's value of plane `step`
        double nx[HM_CPT], nb[HM_CPT];
        const bool bpl = (step == 0 || step == n0 - 1);
        if (step + 1 < E) {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) nx[q] = ((m_in >> q) & 1) ? __ldg(gs + q * n2) : 0.0;
        } else {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) nx[q] = 0.0;
        }
        if ((EDGE || bpl) && step < E) {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q)
                nb[q] = (((m_in >> q) & 1) && (((m_bd >> q) & 1) || bpl)) ? gd[q * n2] : 0.0;
        } else {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) nb[q] = 0.0;
        }
        const int par = step & 1;
        double *S0w = S + par * HM_PLANE + c0;                 // plane step   (holds plane step-2)
        const double *S0r = S + (par ^ 1) * HM_PLANE + c0;     // plane step-1
        double *S1w = S + (2 + (par ^ 1)) * HM_PLANE + c0;     // plane step-1 (holds plane step-3)
        const double *S1r = S + (2 + par) * HM_PLANE + c0;     // plane step-2
        double *S2w = S + (4 + par) * HM_PLANE + c0;           // plane step-2 (holds plane step-4)
        const double *S2r = S + (4 + (par ^ 1)) * HM_PLANE + c0;   // plane step-3
        const int p1 = step - 1, p2 = step - 2, p3 = step - 3;
        const bool bp1 = (p1 == 0 || p1 == n0 - 1), bp2 = (p2 == 0 || p2 == n0 - 1);
        const bool st3 = (p3 >= ia && p3 < ib);

        // ---- state 1, plane step-1
        double v0m[HM_CPT], v1p[HM_CPT];
#pragma unroll
        for (int q = 0; q < HM_CPT; ++q) {
            v0m[q] = S0w[q * HM_RK];
            S0w[q * HM_RK] = v0p[q];
        }
        if (!EDGE && bp1) {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) v1p[q] = b1c[q];
        } else {
            const double top = S0r[-HM_RK], bot = S0r[HM_CPT * HM_RK];
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) {
                const double jm = q ? v0c[q - 1] : top;
                const double jp = (q < HM_CPT - 1) ? v0c[q + 1] : bot;
                const double u = hm_update(v0m[q], v0c[q], v0p[q], jm, jp, S0r[q * HM_RK - 1], S0r[q * HM_RK + 1]);
                v1p[q] = (EDGE && (((m_bd >> q) & 1) || bp1)) ? b1c[q] : u;
            }
        }
        // ---- state 2, plane step-2 (border = src's = v0m)
        double v1m[HM_CPT], v2p[HM_CPT];
#pragma unroll
        for (int q = 0; q < HM_CPT; ++q) {
            v1m[q] = S1w[q * HM_RK];
            S1w[q * HM_RK] = v1p[q];
        }
        if (!EDGE && bp2) {
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) v2p[q] = v0m[q];
        } else {
            const double top = S1r[-HM_RK], bot = S1r[HM_CPT * HM_RK];
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) {
                const double jm = q ? v1c[q - 1] : top;
                const double jp = (q < HM_CPT - 1) ? v1c[q + 1] : bot;
                const double u = hm_update(v1m[q], v1c[q], v1p[q], jm, jp, S1r[q * HM_RK - 1], S1r[q * HM_RK + 1]);
                v2p[q] = (EDGE && (((m_bd >> q) & 1) || bp2)) ? v0m[q] : u;
            }
        }
        // ---- state 3, plane step-3 -> global
        {
            double v2m[HM_CPT];
#pragma unroll
            for (int q = 0; q < HM_CPT; ++q) {
                v2m[q] = S2w[q * HM_RK];
                S2w[q * HM_RK] = v2p[q];
            }
            if (st3) {
                const double top = S2r[-HM_RK], bot = S2r[HM_CPT * HM_RK];
#pragma unroll
                for (int q = 0; q < HM_CPT; ++q) {
                    const double jm = q ? v2c[q - 1] : top;
                    const double jp = (q < HM_CPT - 1) ? v2c[q + 1] : bot;
                    const double u = hm_update(v2m[q], v2c[q], v2p[q], jm, jp, S2r[q * HM_RK - 1], S2r[q * HM_RK + 1]);
                    if ((m_out >> q) & 1) go[q * n2] = u;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < HM_CPT; ++q) {
            v0c[q] = v0p[q]; v0p[q] = nx[q]; v1c[q] = v1p[q]; v2c[q] = v2p[q]; b1c[q] = nb[q];
        }
        gs += ps; gd += ps; go += ps;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(HM_THREADS, 2)
heat3d_march_kernel(HmParams p) {
    extern __shared__ double hm_sm[];
    double *S = hm_sm + HM_PAD;
    const int tj = blockIdx.x / p.tiles_k, tk = blockIdx.x - tj * p.tiles_k;
    // region rows gj in [tj*TJ - 2, tj*TJ + 22), columns gk in [tk*TK - 2, tk*TK + 62): strictly inside the faces?
    const bool edge = (tj == 0) || (tk == 0) || (tj * HM_TJ + HM_RJ - 2 > p.n1 - 1) || (tk * HM_TK + HM_RK - 2 > p.n2 - 1);
    if (edge) hm_march<true>(p, S, tj, tk);
    else hm_march<false>(p, S, tj, tk);
}

// one pass = three sweeps src -> dst over the output planes [i_lo, i_hi) (clamped to the interior).  On a slab of a
// sharded grid planes 0 and n0 - 1 are ghost planes, not constant borders: what the kernel makes of them is garbage
// that creeps one plane per sweep, so the caller keeps >= 3 ghost planes per interior side (distributed.py).
int launch_march(int64_t n0, int64_t n1, int64_t n2, const double *src, double *dst, int64_t i_lo = 1, int64_t i_hi = -1) {
    if (i_lo < 1) i_lo = 1;
    if (i_hi < 0 || i_hi > n0 - 1) i_hi = n0 - 1;
    if (i_lo >= i_hi) return 0;
    const long span = (long)(i_hi - i_lo);
    static bool configured = false;
    if (!configured) {
        NPB_CUDA(cudaFuncSetAttribute(heat3d_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HM_SMEM));
        configured = true;
    }
    const long tiles_j = (long)((n1 - 2 + HM_TJ - 1) / HM_TJ), tiles_k = (long)((n2 - 2 + HM_TK - 1) / HM_TK);
    const long tiles = tiles_j * tiles_k, slots = 2L * npb::st().sm_count;
    // split i into chunks (each pays a 6-plane ramp) until the CTA count fills whole waves
    long best_nc = 1; double best_cost = 0.0;
    for (long nc = 1; nc <= 32 && nc <= span; ++nc) {
        const long planes = (span + nc - 1) / nc;
        const long waves = (tiles * nc + slots - 1) / slots;
        const double cost = (double)waves * (double)(planes + 6);
        if (nc == 1 || cost < best_cost * 0.98) { best_nc = nc; best_cost = cost; }
    }
    static const int env_chunk = getenv("NPB_HEAT_CHUNK") ? atoi(getenv("NPB_HEAT_CHUNK")) : 0;     // planes per chunk
    if (env_chunk > 0) best_nc = (span + env_chunk - 1) / env_chunk;
    const long chunk = (span + best_nc - 1) / best_nc;
    const long nchunks = (span + chunk - 1) / chunk;
    HmParams p{(int)n0, (int)n1, (int)n2, (int)tiles_k, (int)chunk, (int)i_lo, (int)i_hi, src, dst};
    dim3 grid((unsigned)tiles, (unsigned)nchunks);
    heat3d_march_kernel<<<grid, HM_THREADS, HM_SMEM, npb::st().stream>>>(p);
    NPB_CHECK_LAUNCH("heat3d_march_kernel");
    npb::count_launch();
    return 0;
}

// worth it when the grid is far beyond on-chip size and the tiles are mostly full
bool march_eligible(int64_t n0, int64_t n1, int64_t n2) {
    if (n0 < 8 || n1 < 3 || n2 < 3 || n0 >= (1LL << 31)) return false;
    const long tiles_j = (long)((n1 - 2 + HM_TJ - 1) / HM_TJ), tiles_k = (long)((n2 - 2 + HM_TK - 1) / HM_TK);
    if (tiles_j * tiles_k > 65535L * 1024L) return false;
    return true;
}

// 2*(TSTEPS-1) sweeps: an odd number of passes (3 sweeps or 1) ending in dst = B, then one
// single sweep B -> A (the parity argument of jacobi2d.cu's blocked passes).
int run_march(int64_t nsweeps, int64_t n0, int64_t n1, int64_t n2, double *A, double *B) {
    const int64_t M = nsweeps - 1;
    int64_t n = (M + 2) / 3;
    if ((n & 1) == 0) ++n;
    int64_t triples = (M - n) / 2;          // passes that take three sweeps instead of one
    double *src = A, *dst = B;
    for (int64_t q = 0; q < n; ++q) {
        int rc;
        if (triples > 0) { rc = launch_march(n0, n1, n2, src, dst); --triples; }
        else rc = launch_sweep(n0, n1, n2, src, dst, 1, n0 - 1);
        if (rc) return rc;
        double *t = src; src = dst; dst = t;
    }
    return launch_sweep(n0, n1, n2, src, dst, 1, n0 - 1);
}
