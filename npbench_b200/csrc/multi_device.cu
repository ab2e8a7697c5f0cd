// multi_device.cu -- column-sharded hdiff / vadv over several GPUs driven by ONE process.
//
// The reference is single device.  hdiff and vadv need no run-time exchange: split along I,
// every shard gets its rows of out / coeff (utens_stage, u_stage, u_pos, utens) and a fixed overlap of the input
// (4 rows of in_field, hdiff_numpy.py:7-28; 1 row of wcon, vadv_numpy.py:16,33-34) at copy-in.  So no NCCL and no
// second process: the entry points below select each device slot in turn (npb_mg_init / npb_mg_select, runtime.cu),
// enqueue the single-device kernel on that device's stream and return; npb_sync() waits for all of them.
// This is what NPBench's plugin reaches with NPB_B200_GPUS=N (scatter in setup_str, gather in copy_back_func) and what
// a non-Python host binds instead of torch.distributed.
#include "common.cuh"

namespace {

#define NPB_TRY(call)            \
    do {                         \
        int rc_ = (call);        \
        if (rc_) { npb_mg_select(0); return rc_; } \
    } while (0)

}  // namespace

// rows [lo, hi) of an n-row axis owned by shard `s` of `nshards` (remainder spread over the first shards;
// npbench_b200/distributed.py: slab_bounds)
extern "C" int npb_shard_bounds(int64_t n, int nshards, int s, int64_t *lo, int64_t *hi) {
    NPB_ARG(nshards >= 1 && s >= 0 && s < nshards && n >= 0 && lo && hi, "npb_shard_bounds", "bad shard index");
    const int64_t base = n / nshards, rem = n % nshards;
    *lo = s * base + (s < rem ? s : rem);
    *hi = *lo + base + (s < rem ? 1 : 0);
    return 0;
}

// Device-pointer form.  Shard s lives on device slot slots[s] and holds output rows [i_lo[s], i_lo[s+1]):
// in_shards[s] = in_field rows [i_lo[s], i_lo[s+1] + 4) (all J + 4 columns), out / coeff shards the owned rows.
extern "C" int npb_hdiff_f64_mg(int nshards, const int *slots, int64_t I, int64_t J, int64_t K,
                                const double *const *in_shards, double *const *out_shards,
                                const double *const *coeff_shards, const int64_t *i_lo) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nshards >= 1 && slots && in_shards && out_shards && coeff_shards && i_lo, "npb_hdiff_f64_mg", "null argument");
    NPB_ARG(i_lo[0] == 0 && i_lo[nshards] == I, "npb_hdiff_f64_mg", "shard bounds must cover [0, I)");
    for (int s = 0; s < nshards; ++s) {
        const int64_t rows = i_lo[s + 1] - i_lo[s];
        NPB_ARG(rows >= 0, "npb_hdiff_f64_mg", "shard bounds must be non-decreasing");
        if (rows == 0) continue;
        NPB_TRY(npb_mg_select(slots[s]));
        NPB_TRY(npb_hdiff_f64(rows, J, K, in_shards[s], out_shards[s], coeff_shards[s]));
    }
    return npb_mg_select(0);
}

// wcon_shards[s] = wcon rows [i_lo[s], i_lo[s+1] + 1); the other fields the owned rows
extern "C" int npb_vadv_f64_mg(int nshards, const int *slots, int64_t I, int64_t J, int64_t K,
                               double *const *utens_stage, const double *const *u_stage, const double *const *wcon_shards,
                               const double *const *u_pos, const double *const *utens, double dtr_stage,
                               const int64_t *i_lo) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nshards >= 1 && slots && utens_stage && u_stage && wcon_shards && u_pos && utens && i_lo, "npb_vadv_f64_mg",
            "null argument");
    NPB_ARG(i_lo[0] == 0 && i_lo[nshards] == I, "npb_vadv_f64_mg", "shard bounds must cover [0, I)");
    for (int s = 0; s < nshards; ++s) {
        const int64_t rows = i_lo[s + 1] - i_lo[s];
        NPB_ARG(rows >= 0, "npb_vadv_f64_mg", "shard bounds must be non-decreasing");
        if (rows == 0) continue;
        NPB_TRY(npb_mg_select(slots[s]));
        NPB_TRY(npb_vadv_f64(rows, J, K, utens_stage[s], u_stage[s], wcon_shards[s], u_pos[s], utens[s], dtr_stage));
    }
    return npb_mg_select(0);
}

// Host-buffer forms (the call a NumPy user or a non-Python host makes): scatter with the overlaps, run every shard
// on its device, gather, synchronise.  Shard s runs on device slot s % npb_mg_count().
extern "C" int npb_hdiff_f64_mg_host(int nshards, int64_t I, int64_t J, int64_t K, const double *in_field,
                                     double *out_field, const double *coeff) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nshards >= 1 && nshards <= 64 && I >= 0 && J >= 0 && K >= 0, "npb_hdiff_f64_mg_host", "bad extents");
    if (I == 0 || J == 0 || K == 0) return 0;
    const int ndev = npb_mg_count();
    void *din[64] = {nullptr}, *dout[64] = {nullptr}, *dco[64] = {nullptr};
    const size_t in_row = (size_t)(J + 4) * K * sizeof(double), row = (size_t)J * K * sizeof(double);
    int rc = 0;
    for (int s = 0; s < nshards && !rc; ++s) {
        int64_t lo, hi;
        npb_shard_bounds(I, nshards, s, &lo, &hi);
        if (hi == lo) continue;
        rc = npb_mg_select(s % ndev);
        if (!rc) rc = npb_malloc((size_t)(hi - lo + 4) * in_row, &din[s]);
        if (!rc) rc = npb_malloc((size_t)(hi - lo) * row, &dout[s]);
        if (!rc) rc = npb_malloc((size_t)(hi - lo) * row, &dco[s]);
        if (!rc) rc = npb_h2d(din[s], (const char *)in_field + (size_t)lo * in_row, (size_t)(hi - lo + 4) * in_row);
        if (!rc) rc = npb_h2d(dco[s], (const char *)coeff + (size_t)lo * row, (size_t)(hi - lo) * row);
        if (!rc) rc = npb_hdiff_f64(hi - lo, J, K, (const double *)din[s], (double *)dout[s], (const double *)dco[s]);
        if (!rc) rc = npb_d2h((char *)out_field + (size_t)lo * row, dout[s], (size_t)(hi - lo) * row);
    }
    npb_mg_select(0);
    const int rs = npb_sync();
    for (int s = 0; s < nshards; ++s) { npb_free(din[s]); npb_free(dout[s]); npb_free(dco[s]); }
    return rc ? rc : rs;
}

extern "C" int npb_vadv_f64_mg_host(int nshards, int64_t I, int64_t J, int64_t K, double *utens_stage,
                                    const double *u_stage, const double *wcon, const double *u_pos, const double *utens,
                                    double dtr_stage) {
    NPB_REQUIRE_INIT();
    NPB_ARG(nshards >= 1 && nshards <= 64 && I >= 0 && J >= 0 && K >= 2, "npb_vadv_f64_mg_host", "bad extents");
    if (I == 0 || J == 0) return 0;
    const int ndev = npb_mg_count();
    void *d[64][5] = {{nullptr}};
    const size_t row = (size_t)J * K * sizeof(double);
    const double *src[5] = {utens_stage, u_stage, wcon, u_pos, utens};
    int rc = 0;
    for (int s = 0; s < nshards && !rc; ++s) {
        int64_t lo, hi;
        npb_shard_bounds(I, nshards, s, &lo, &hi);
        if (hi == lo) continue;
        rc = npb_mg_select(s % ndev);
        for (int f = 0; f < 5 && !rc; ++f) {
            const size_t rows = (size_t)(hi - lo) + (f == 2 ? 1 : 0);
            rc = npb_malloc(rows * row, &d[s][f]);
            if (!rc) rc = npb_h2d(d[s][f], (const char *)src[f] + (size_t)lo * row, rows * row);
        }
        if (!rc) rc = npb_vadv_f64(hi - lo, J, K, (double *)d[s][0], (const double *)d[s][1], (const double *)d[s][2],
                                   (const double *)d[s][3], (const double *)d[s][4], dtr_stage);
        if (!rc) rc = npb_d2h((char *)utens_stage + (size_t)lo * row, d[s][0], (size_t)(hi - lo) * row);
    }
    npb_mg_select(0);
    const int rs = npb_sync();
    for (int s = 0; s < nshards; ++s)
        for (int f = 0; f < 5; ++f) npb_free(d[s][f]);
    return rc ? rc : rs;
}
