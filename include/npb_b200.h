/*
 * npb_b200.h -- C ABI of libnpb_b200.so, the B200 (sm_100a) backend for the
 * NPBench structured-grid stencil family.
 *
 * This is the drop-in boundary: plain C, raw pointers and sizes, no C++/torch
 * types.  The NPBench side binds it with ctypes (npbench_b200/_lib.py; the
 * stub a reference maintainer would add is in INTEGRATION.md).  Citations are
 * file:line inside spcl/npbench.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     npb_last_error() then returns a static, human readable message;
 *   - one process drives one GPU (npb_init(device)); library state is process
 *     global, calls are not re-entrant (the NPBench harness is single threaded:
 *     npbench/infrastructure/test.py:16-51);
 *   - kernel entry points taking DEVICE pointers only enqueue work on the
 *     library's current stream and return; npb_sync() is the blocking call
 *     (same semantics the CuPy plugin gets from stream.synchronize(),
 *     npbench/infrastructure/cupy_framework.py:47-58);
 *   - *_host entry points take HOST pointers: they copy in, run, copy the
 *     validated outputs back and return after synchronising -- the exact call
 *     a NumPy user makes (in-place mutation, bench_info/<b>.json output_args);
 *   - all arrays are C-contiguous float64 (checked by the Python side:
 *     every NPBench array_arg is);
 *   - results are bit-identical to NPBench's NumPy implementations: kernels
 *     keep NumPy's evaluation order and are compiled with -fmad=false.
 */
#ifndef NPB_B200_H
#define NPB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ---------------------------------------------------------- */
int npb_init(int device);                /* select device, create stream + pool; idempotent */
int npb_shutdown(void);
const char *npb_version(void);           /* Framework.version(): framework.py:33-35 */
const char *npb_last_error(void);
int npb_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes,
                    size_t *smem_per_block_optin, size_t *total_mem);

/* Use a caller-owned cudaStream_t (e.g. torch's current stream) for all
 * subsequent calls; NULL restores the library stream. */
int npb_set_stream(void *cuda_stream);
void *npb_get_stream(void);
int npb_sync(void);                      /* the sync appended to exec_str (cupy_framework.py:56-58); waits for every
                                          * device this process drives */

/* ---- single-process multi-device (the `_mg` variants; 8(e): hdiff / vadv column shards).
 *      The reference is single device; NPBench's harness is one process, so sharding that needs no exchange is driven
 *      from that one process: slot 0 is npb_init's device, npb_mg_init(n, devices) makes slot k drive devices[k]
 *      (a device may repeat).  npb_mg_select(k) redirects every following call (allocator, copies, kernels, timers)
 *      to slot k; select 0 to return.  NPBench reaches this with NPB_B200_GPUS=N (plugin: scatter in setup_str,
 *      framework.py:139-150; gather in copy_back_func). */
int npb_mg_init(int ndev, const int *devices);
int npb_mg_count(void);
int npb_mg_select(int slot);
int npb_mg_current(void);
int npb_shard_bounds(int64_t n, int nshards, int s, int64_t *lo, int64_t *hi);   /* rows [lo, hi) of shard s */
/* hdiff split along I: shard s = output rows [i_lo[s], i_lo[s+1]) on slot slots[s]; in_shards[s] holds in_field rows
 * [i_lo[s], i_lo[s+1] + 4) (fixed 4-row overlap, hdiff_numpy.py:7-28); i_lo has nshards + 1 entries, 0 .. I */
int npb_hdiff_f64_mg(int nshards, const int *slots, int64_t I, int64_t J, int64_t K,
                     const double *const *in_shards, double *const *out_shards,
                     const double *const *coeff_shards, const int64_t *i_lo);
/* vadv split along I: wcon_shards[s] holds wcon rows [i_lo[s], i_lo[s+1] + 1) (vadv_numpy.py:16, 33-34) */
int npb_vadv_f64_mg(int nshards, const int *slots, int64_t I, int64_t J, int64_t K,
                    double *const *utens_stage, const double *const *u_stage, const double *const *wcon_shards,
                    const double *const *u_pos, const double *const *utens, double dtr_stage, const int64_t *i_lo);
/* the same on HOST buffers with the NumPy signatures' array layout: scatter, run, gather, synchronise;
 * shard s runs on slot s % npb_mg_count() */
int npb_hdiff_f64_mg_host(int nshards, int64_t I, int64_t J, int64_t K, const double *in_field,
                          double *out_field, const double *coeff);
int npb_vadv_f64_mg_host(int nshards, int64_t I, int64_t J, int64_t K, double *utens_stage, const double *u_stage,
                         const double *wcon, const double *u_pos, const double *utens, double dtr_stage);

/* ---- device memory: what Framework.copy_func / copy_back_func need
 *      (framework.py:42-50; called once per array per repetition,
 *      framework.py:147-150) -- a caching pool, no cudaFree on the hot path */
int npb_malloc(size_t bytes, void **dptr);
int npb_free(void *dptr);
int npb_pool_trim(void);                 /* return cached blocks to the driver */
int npb_host_alloc(size_t bytes, void **hptr);   /* pinned host memory */
int npb_host_free(void *hptr);
int npb_h2d(void *dst_dev, const void *src_host, size_t bytes);   /* async on the stream */
int npb_d2h(void *dst_host, const void *src_dev, size_t bytes);   /* async on the stream */
int npb_d2d(void *dst_dev, const void *src_dev, size_t bytes);
int npb_memset(void *dst_dev, int value, size_t bytes);

/* ---- timing / accounting ---------------------------------------------- */
int npb_timer_start(void);               /* cudaEventRecord on the current stream */
int npb_timer_stop(float *ms);           /* record + synchronise + elapsed */
uint64_t npb_launch_count(void);         /* kernels launched by this library so far */
int npb_l2_flush(void);                  /* overwrite a buffer larger than L2 */

/* ---- the five kernels, DEVICE pointers -------------------------------- */

/* kernel(TSTEPS, A, B): polybench/jacobi_2d/jacobi_2d_numpy.py:4-10.
 * A, B (ni, nj); 2*(TSTEPS-1) sweeps; both A and B are outputs. */
int npb_jacobi2d_f64(int64_t tsteps, int64_t ni, int64_t nj, double *A, double *B);

/* One temporally blocked pass: `nsteps` (odd, 1..NPB_JACOBI2D_MAX_BLOCK)
 * sweeps src -> dst fused in shared memory, restricted to tile rows
 * [tile_row_lo, tile_row_hi) (pass 0, -1 for all).  Building block of the
 * slab-sharded multi-GPU driver (boundary tiles first, interior overlapped
 * with the halo exchange). */
#define NPB_JACOBI2D_MAX_BLOCK 7
int npb_jacobi2d_block_f64(int nsteps, int64_t ni, int64_t nj, const double *src, double *dst,
                           int64_t tile_row_lo, int64_t tile_row_hi);
/* The same pass, which ALSO stores the state before its last sweep into dst2 (a third array; interior cells of the
 * same rows).  The closing pass of a sharded run leaves state S in A and state S - 1 in B this way -- no separate
 * single sweep (no reference counterpart: the reference is single device, jacobi_2d_numpy.py:8-10 define the two
 * states).  Marching regime only: npb_jacobi2d_block_marches(ni, nj) (host logic, no device work) says whether an
 * (ni, nj) slab is in it. */
int npb_jacobi2d_block2_f64(int nsteps, int64_t ni, int64_t nj, const double *src, double *dst, double *dst2,
                            int64_t tile_row_lo, int64_t tile_row_hi);
int npb_jacobi2d_block_marches(int64_t ni, int64_t nj);
/* mode & 7: 0 dispatch by size (grids that fit on chip -- NPBench S / M / L -- run in ONE cooperative launch:
 * jacobi2d_regtile_kernel, cell state in registers, T sweeps per halo exchange through in-L2 inboxes; grids of
 * >= 14M cells: marching passes, jacobi2d_march_kernel, 3/5/7 sweeps per pass in registers; else blocked
 * shared-memory passes), 1 blocked passes, 2 same as 0, 3 marching passes at any size (nj >= 8);
 * mode >> 8 = rows per chunk of the marching kernel (0 = auto) */
int npb_jacobi2d_set_mode(int mode);
int npb_jacobi2d_last_path(void);        /* 1 register-tile resident kernel, 2 blocked passes, 3 marching passes */
int npb_jacobi2d_last_passes(void);      /* passes over memory of the last npb_jacobi2d_f64 call (0 for the register-tile kernel) */
/* host logic only (no device work): the passes npb_jacobi2d_f64 runs a grid in the marching regime with.  dual != 0: the
 * scratch-grid plan (an even number of odd passes, the last one stores the states S and S - 1); dual == 0: the plan
 * used without scratch memory (odd passes + a closing single sweep).  Writes min(passes, cap) entries, returns the
 * number of passes. */
int npb_jacobi2d_pass_plan(int64_t tsteps, int dual, int32_t *sweeps, int cap);
/* host logic only: rows per chunk (gridDim.y = ceil(rows / this)) of one marching launch of `ns` (1, 3, 5, 7) sweeps over
 * `rows` rows of an nj-column grid on `sms` SMs; 0 for arguments out of range */
int64_t npb_jacobi2d_march_rows_per_chunk(int ns, int64_t rows, int64_t nj, int sms);
/* configuration of the last register-tile launch: {rows, columns of cells per thread, warps per CTA, sweeps per
 * halo exchange, tiles along i, tiles along j, CTAs per SM} */
int npb_jacobi2d_regtile_config(int *out7);
/* host logic only (no device work): the configuration a grid would run with on `sms` SMs; 1 and out7 filled, or 0
 * if the grid does not run resident */
int npb_jacobi2d_regtile_plan(int64_t tsteps, int64_t ni, int64_t nj, int sms, int *out7);
int npb_jacobi2d_tile_rows(void);        /* rows per tile of the blocked kernel */

/* kernel(TSTEPS, A, B): polybench/heat_3d/heat_3d_numpy.py:4-20.  (n0,n1,n2). */
int npb_heat3d_f64(int64_t tsteps, int64_t n0, int64_t n1, int64_t n2, double *A, double *B);
/* mode & 7: 0 = dispatch by size (grids that fit on chip -- NPBench S / M / L -- run in ONE cooperative launch:
 * heat3d_regtile_kernel, cell state in registers, faces through shared memory, halos through in-L2 inboxes;
 * if its limits do not fit, heat3d_resident_kernel (state in shared memory); grids of >= 40M cells with
 * n1, n2 >= 128: three sweeps per pass over HBM, heat3d_march_kernel; else one streaming launch per sweep,
 * replayed as a CUDA graph); 1 = always streaming; 2 = the shared-memory resident kernel when eligible;
 * 5 = three-sweep marching passes at any shape with n0 >= 8; 6 = the register-tile kernel or an error.
 * +8: no graph.  +256 / +512 / +1024: timing experiments of the register-tile kernel (no fences / no polls /
 * no sends -- results are wrong by construction).  A failed cooperative launch is an error, never a silent
 * fallback to a slower path. */
int npb_heat3d_set_mode(int mode);
int npb_heat3d_last_path(void);          /* last call: 1 shared-memory resident, 2 streaming, 5 marching, 6 register-tile resident */
int npb_heat3d_set_trace(void *dev_buf); /* profiling aid: 10 int64 phase cycle counters of the register-tile kernel's centre CTA, NULL = off */
/* THREE sweeps src -> dst in one pass over memory (heat3d_march_kernel; reference loop body heat_3d_numpy.py:7-19),
 * output planes [i_lo, i_hi) (clamped to the interior).  State 1 takes its constant j / k borders from dst, state 2
 * from src, exactly as three one-sweep launches src -> dst -> src -> dst would.  On a slab, planes closer than 3 to
 * an edge that is not a grid edge come out as garbage: the caller keeps >= 3 ghost planes. */
int npb_heat3d_march_f64(int64_t n0, int64_t n1, int64_t n2, const double *src, double *dst,
                         int64_t i_lo, int64_t i_hi);
/* one sweep src -> dst over planes [i_lo, i_hi) (clamped to the interior) */
int npb_heat3d_sweep_f64(int64_t n0, int64_t n1, int64_t n2, const double *src, double *dst,
                         int64_t i_lo, int64_t i_hi);

/* kernel(TMAX, ex, ey, hz, _fict_): polybench/fdtd_2d/fdtd_2d_numpy.py:4-11. */
int npb_fdtd2d_f64(int64_t tmax, int64_t nx, int64_t ny, double *ex, double *ey, double *hz,
                   const double *fict);
/* mode & 3: 0 = dispatch by size (grids that fit on chip -- NPBench S / M / L -- run in ONE cooperative launch:
 * fdtd2d_regtile_kernel, the three fields in registers, T steps per halo exchange through in-L2 inboxes; grids of
 * >= 4M cells: marching passes of up to five time steps, fdtd2d_march_kernel; else one launch per step),
 * 1 = always one launch per step, 2 = marching passes at any size (TMAX >= 2), 3 = same as 0;
 * mode >> 8 = rows per chunk of the marching kernel (0 = automatic). */
int npb_fdtd2d_set_mode(int mode);
int npb_fdtd2d_last_path(void);          /* last call: 1 one launch per step, 2 marching passes, 3 register-tile resident kernel */
/* configuration of the last register-tile launch: {rows, columns of cells per thread, warps per CTA, steps per
 * halo exchange, tiles along i, tiles along j} */
int npb_fdtd2d_regtile_config(int *out6);
int npb_fdtd2d_regtile_plan(int64_t tmax, int64_t nx, int64_t ny, int sms, int *out6);   /* host logic only, as above */
/* host logic only: the steps-per-pass plan of npb_fdtd2d_f64 (march != 0: up to five steps per pass, an even
 * number of passes when TMAX allows); writes min(passes, cap) entries, returns the number of passes */
int npb_fdtd2d_pass_plan(int64_t tmax, int march, int32_t *steps, int cap);
/* one fused time step on a row slab: local rows [0, nrows) are global rows
 * [row0, row0+nrows) of an nx_global-row grid; src fields -> dst fields
 * (out of place); `fict_t` is _fict_[t].  Rows whose stencil leaves the slab
 * are copied (they are ghost rows of the sharded driver).  Only local rows
 * [row_lo, row_hi) are written (0, -1 for all): boundary rows first, interior
 * overlapped with the halo exchange. */
int npb_fdtd2d_step_f64(int64_t nx_global, int64_t row0, int64_t nrows, int64_t ny,
                        const double *ex, const double *ey, const double *hz,
                        double *ex_out, double *ey_out, double *hz_out, double fict_t,
                        int64_t row_lo, int64_t row_hi);

/* ns (2..5) fused time steps on a row slab in ONE pass over memory (fdtd2d_march_kernel; the reference loop body is
 * fdtd_2d_numpy.py:6-11): src fields -> dst fields, output rows [row_lo, row_hi) (0, -1 for all).  `fict_dev` points
 * at _fict_[t] of the first step IN DEVICE MEMORY (ns consecutive values are read).  Rows closer than ns to a slab
 * edge that is not a grid edge come out as garbage: the caller keeps >= ns ghost rows (npbench_b200/distributed.py). */
int npb_fdtd2d_march_f64(int ns, int64_t nx_global, int64_t row0, int64_t nrows, int64_t ny,
                         const double *ex, const double *ey, const double *hz,
                         double *ex_out, double *ey_out, double *hz_out, const double *fict_dev,
                         int64_t row_lo, int64_t row_hi);

/* hdiff(in_field, out_field, coeff): weather_stencils/hdiff/hdiff_numpy.py:5-29.
 * in (I+4, J+4, K); out, coeff (I, J, K). */
int npb_hdiff_f64(int64_t I, int64_t J, int64_t K, const double *in_field, double *out_field,
                  const double *coeff);
/* 0 = dispatch by size; 1 = register-marching kernel; 2 = TMA-bulk ring kernel when legal */
int npb_hdiff_set_mode(int mode);
int npb_hdiff_last_path(void);           /* 1 marching, 2 ring */

/* vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage):
 * weather_stencils/vadv/vadv_numpy.py:9-78.  All (I,J,K) but wcon (I+1,J,K). K >= 2. */
int npb_vadv_f64(int64_t I, int64_t J, int64_t K, double *utens_stage, const double *u_stage,
                 const double *wcon, const double *u_pos, const double *utens, double dtr_stage);

int npb_vadv_set_mode(int mode);         /* 0 dispatch (streaming TMEM/TMA solver for even K <= 256, else tile kernel), 1 tile kernel,
                                            2 TMA-fed tile kernel (experimental), 3/4/5 streaming solver variants */
int npb_vadv_last_path(void);            /* 1 tile kernel, 2 TMA-fed tile kernel, 3 streaming solver */
int npb_vadv_set_trace(void *dev_buf);   /* profiling aid: per-group phase timestamps (ngroups*8 u64), NULL = off */

/* ---- the same five calls on HOST buffers (copy in, run, copy outputs back,
 *      synchronise): the NumPy-signature call of bench_info/<b>.json ------- */
int npb_jacobi2d_f64_host(int64_t tsteps, int64_t ni, int64_t nj, double *A, double *B);
int npb_heat3d_f64_host(int64_t tsteps, int64_t n0, int64_t n1, int64_t n2, double *A, double *B);
int npb_fdtd2d_f64_host(int64_t tmax, int64_t nx, int64_t ny, double *ex, double *ey, double *hz,
                        const double *fict);
int npb_hdiff_f64_host(int64_t I, int64_t J, int64_t K, const double *in_field, double *out_field,
                       const double *coeff);
int npb_vadv_f64_host(int64_t I, int64_t J, int64_t K, double *utens_stage, const double *u_stage,
                      const double *wcon, const double *u_pos, const double *utens,
                      double dtr_stage);

/* ---- widening row: the next two structured-grid kernels of the suite
 *      on the same plugin path.
 * kernel(TSTEPS, A, B): polybench/jacobi_1d/jacobi_1d_numpy.py:4-8.  A, B of length n; end cells untouched. */
int npb_jacobi1d_f64(int64_t tsteps, int64_t n, double *A, double *B);
int npb_jacobi1d_f64_host(int64_t tsteps, int64_t n, double *A, double *B);
/* kernel(TSTEPS, N, A): polybench/seidel_2d/seidel_2d_numpy.py:4-13.  A is (n, n), updated in place
 * (Gauss-Seidel order: row by row, west to east, TSTEPS-1 sweeps). */
int npb_seidel2d_f64(int64_t tsteps, int64_t n, double *A);
int npb_seidel2d_f64_host(int64_t tsteps, int64_t n, double *A);
int npb_seidel2d_set_mode(int mode);     /* 0 dispatch (distributed-shared-memory kernel when the grid fits in one cluster), 1 L2 wavefront kernel */
int npb_seidel2d_last_path(void);        /* 1 distributed-shared-memory kernel, 2 L2 wavefront kernel */

/* widening row, rank 2 -- kernel(TSTEPS, N, u): polybench/adi/adi_numpy.py:6-54.  u is (n, n), updated in place;
 * TSTEPS >= 1 (the reference divides by it). */
int npb_adi_f64(int64_t tsteps, int64_t n, double *u);
int npb_adi_f64_host(int64_t tsteps, int64_t n, double *u);

/* widening row, rank 3 -- cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu):
 * cavity_flow/cavity_flow_numpy.py:46-89.  u, v, p are (ny, nx), updated in place; nx, ny >= 3. */
int npb_cavity_flow_f64(int64_t nx, int64_t ny, int64_t nt, int64_t nit, double *u, double *v, double dt, double dx,
                        double dy, double *p, double rho, double nu);
int npb_cavity_flow_f64_host(int64_t nx, int64_t ny, int64_t nt, int64_t nit, double *u, double *v, double dt, double dx,
                             double dy, double *p, double rho, double nu);

/* channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F) -> stepcount: channel_flow/channel_flow_numpy.py:74-170.
 * u, v, p are (ny, nx), updated in place; the call blocks (the convergence test needs one double per step). */
int npb_channel_flow_f64(int64_t nit, int64_t nx, int64_t ny, double *u, double *v, double dt, double dx, double dy,
                         double *p, double rho, double nu, double F, int64_t *stepcount);
int npb_channel_flow_f64_host(int64_t nit, int64_t nx, int64_t ny, double *u, double *v, double dt, double dx, double dy,
                              double *p, double rho, double nu, double F, int64_t *stepcount);

/* ---- device-side initialisers (NPBench `initialize`, closed forms):
 *      jacobi_2d.py:6-10, heat_3d.py:6-11, fdtd_2d.py:6-15.  Rows
 *      [row0, row0+nrows) of the global grid, for the scaled / sharded grids. */
int npb_init_jacobi2d_f64(int64_t n_global, int64_t row0, int64_t nrows, int64_t ncols, double *A,
                          double *B);
int npb_init_heat3d_f64(int64_t n_global, int64_t row0, int64_t nrows, double *A, double *B);
int npb_init_fdtd2d_f64(int64_t tmax, int64_t nx_global, int64_t ny, int64_t row0, int64_t nrows,
                        double *ex, double *ey, double *hz, double *fict);

#ifdef __cplusplus
}
#endif
#endif /* NPB_B200_H */
