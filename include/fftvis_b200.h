/*
 * fftvis_b200 -- C ABI of the B200-native (sm_100a) visibility hot path.
 *
 * Drop-in boundary for the reference's stubbed ``backend="gpu"`` slot
 * (/root/reference/src/fftvis/gpu/{gpu_simulate,beams,nufft,utils}.py; selected at
 * /root/reference/src/fftvis/wrapper.py:77-80).  The reference has no FFI of its own (it is pure
 * Python over third-party natives), so every entry point below cites the *Python* interface whose
 * arithmetic it replaces.  The host-side mirror (fftvis_b200/gpu/) binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative fv_status on a library error, or a
 *     positive cudaError_t; fv_last_error_string() describes the last failure on this thread.
 *   - no exceptions, no torch types: plain pointers, sizes and a cudaStream_t (as void*).
 *   - all array pointers are DEVICE pointers owned by the caller unless the name ends in _host.
 *   - prec: 1 = float32 / complex64, 2 = float64 / complex128 ("real"/"cplx" below).
 *   - complex numbers are interleaved (re, im); sign convention exp(+i ...) (finufft isign=+1,
 *     reference cpu/nufft.py:48,105,162).
 *   - `n_dev` is a device int32 holding the live number of sources (written by fv_rotate_cut), so
 *     that a whole time step can be enqueued without a host synchronisation; `n_cap` is the
 *     capacity every per-source buffer was allocated with.
 */
#ifndef FFTVIS_B200_H
#define FFTVIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FV_OK = 0,
  FV_ERR_INVALID = -1,   /* bad argument */
  FV_ERR_ALLOC = -2,     /* device allocation failed */
  FV_ERR_CUFFT = -3,     /* cuFFT failure */
  FV_ERR_UNSUPPORTED = -4,
  FV_ERR_NO_DEVICE = -5
} fv_status;

const char* fv_last_error_string(void);
int fv_version(void);
/* number of CUDA devices visible; fails with FV_ERR_NO_DEVICE when there is none */
int fv_device_count(int* count_host);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t fv_launch_count(void);

/* Strided host <-> device copy on `stream` (direction 0: device -> host, 1: host -> device; pitches and
 * width in bytes).  Replaces the host scatter `vis[tc][..., fc] = future` (cpu_simulate.py:846-847):
 * each finished time slab of the (nf, nt, P, nbls) result is streamed into the caller's page-locked
 * array while later time steps are still being computed. */
int fv_memcpy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width,
                      int64_t height, int direction, void* stream);

/* ---- kernel / grid parameter rules (host only; finufft's published rules) */
int fv_kernel_params(double eps, double upsampfac, int prec, int* w_host, double* beta_host);
int64_t fv_next235even(int64_t n);

/* ---- a1+a2+a3: rotate, horizon cut, compaction, az/za, array-plane rotation ------------------
 * Replaces coord_mgr.rotate(ti) + select_chunk (cpu_simulate.py:937-946), enu_to_az_za
 * (:957-959), inplace_rot (cpu/utils.py:5-24; calls :961-965) and topo *= 2*pi (:967).
 *   eq_xyz     (3, nsrc) fp64 equatorial unit vectors
 *   enu_mat    9 fp64, row-major: enu = M @ p        (per time; core/astrometry.py: latitude tilt .
 *              polar motion . R3(local Earth rotation angle) . celestial-to-intermediate matrix)
 *   astrom     NULL, or 10 fp64 per time (matvis CoordinateRotationERFA's apco block, cpu_simulate.py:
 *              693-709): Sun->observer unit vector [3], its length in au, observer barycentric
 *              velocity / c [3], sqrt(1 - v^2), deflection limiter, flag (0: skip).  When present every
 *              source direction gets the Sun's light deflection and the aberration before M
 *              (erfa ldsun + ab arithmetic, in fp64), i.e. p above is the proper direction.
 *   plane_mat  9 fp64, row-major: rotation R (type 3) or basis_matrix^T / c (type 1);
 *              applied in working precision in the reference's operation order, then * 2 pi
 *   src_lo/hi  catalogue slice [lo, hi) handled by this call (the reference's `nchunks`)
 * outputs (capacity n_cap each): xyz (3, n_cap) real = 2 pi * plane_mat @ enu, az, za real,
 *   src_idx int32 (catalogue index of each kept source, ascending), n_dev int32 count.
 *   If more than n_cap sources are above the horizon, n_dev is set to -(count) and nothing else
 *   is valid (the reference raises "increase source_buffer").
 *   scratch: at least fv_rotate_cut_scratch_bytes(nsrc) bytes.
 */
int64_t fv_rotate_cut_scratch_bytes(int64_t nsrc);
int fv_rotate_cut(int prec, const double* eq_xyz, int64_t nsrc, int64_t src_lo, int64_t src_hi,
                  const double* enu_mat_host, const double* astrom_host, const double* plane_mat_host,
                  void* xyz, void* az, void* za, int32_t* src_idx, int64_t n_cap, int32_t* n_dev,
                  void* scratch, void* stream);

/* in-place b[:, s] <- rot @ b[:, s]  (gpu/utils.py:8 `inplace_rot`; cpu/utils.py:5-24) */
int fv_inplace_rot(int prec, const double* rot_host, void* b /* (3, n) real */, int64_t n,
                   void* stream);

/* ---- a4+a5: beam evaluation and apparent coherency -> NUFFT strengths ------------------------
 * Replaces CPUBeamEvaluator.evaluate_beam (cpu/beams.py:12-89), _evaluate_beam_list
 * (cpu_simulate.py:38-87) and _compute_apparent_coherency with its four numba kernels
 * (cpu_simulate.py:90-202, cpu/beams.py:129-246), batched over `nf` frequencies.
 */
typedef struct {
  int32_t kind;        /* 0 Gaussian, 1 Airy, 2 uniform, 3 az/za table */
  int32_t is_power;    /* 1: scalar power beam (unpolarised path); 0: 2x2 E-field */
  double diameter;     /* analytic beams (metres) */
  /* table: power (nfreq_table, nza, naz_ext) real; E-field (nfreq_table, 4, nza, naz_ext) cplx with
   * the four Jones entries in [vec*2+feed] order; azimuth axis already wrap-extended by the host
   * when the grid is periodic */
  const void* table;
  int32_t nza, naz;    /* naz: extended length */
  int32_t az_wrap_period; /* >0: az index is taken modulo this before adding az_pad */
  int32_t az_pad;
  double az0, daz, za0, dza;
  int32_t order;       /* interpolation order: 0 nearest, 1 bilinear, 3 cubic B-spline (the table then
                          holds spline COEFFICIENTS: scipy.ndimage.spline_filter of the edge-padded
                          grid, exactly what map_coordinates(order=3, mode="nearest") evaluates) */
  int32_t freq_offset; /* table frequency index of the batch's first frequency */
  int32_t spline_pad;  /* order 3: edge padding (12) added on every side before prefiltering; nza/naz
                          count the padded table */
  int32_t reserved;
} fv_beam;

/* mode: 0 unpolarised  W = sqrt(B_i B_j) F                     (cpu_simulate.py:183-186)
 *       1 polarised beam, unpolarised sky  A_i^H diag(F) A_j   (cpu/beams.py:129-145,182-212)
 *       2 polarised beam, polarised sky    A'_i^H C A'_j, A' = axis-0 flipped (cpu/beams.py:147-180,
 *         215-246; flip at cpu_simulate.py:146-147,153)
 * flux: mode 0/1 (nfreq_total, nsrc_total) cplx frequency-major; mode 2 (nfreq_total, 4, nsrc_total).
 * out:  (nf, P, n_cap) cplx, P = 1 (mode 0) or 4, row order [a*2+p] (cpu_simulate.py:191).
 * out_beam_i (optional, may be NULL): the evaluated beam i, (nf, 4 | 1, n_cap) cplx.
 */
int fv_weights(int prec, int mode, const fv_beam* beam_i_host, const fv_beam* beam_j_host,
               const void* az, const void* za, const int32_t* src_idx, const int32_t* n_dev,
               int64_t n_cap, const double* freqs /* device, all frequencies */, int nf,
               int64_t freq_index0,
               const void* flux, int64_t nsrc_total, void* out, void* out_beam_i, void* stream);

/* Basis path (_compute_basis_visibilities, cpu_simulate.py:303-470, weights part :416-450): the K basis
 * beams (beams_host[0..K), E-field, K <= 8) are evaluated once per (source, frequency) and all
 * K (K + 1) / 2 pair products (k <= l, row-major) are written as the strengths of one batched NUFFT:
 * out (nf, npairs * 4, n_cap) cplx, pair q at transforms 4 q .. 4 q + 3.  mode 1 or 2 as in fv_weights. */
int fv_weights_basis(int prec, int mode, const fv_beam* beams_host, int K, const void* az, const void* za,
                     const int32_t* src_idx, const int32_t* n_dev, int64_t n_cap, const double* freqs, int nf,
                     int64_t freq_index0, const void* flux, int64_t nsrc_total, void* out, void* stream);

/* ---- az/za table beams staged in shared memory (north_star (2): "the beam grid staged in shared memory,
 * TMA where it tiles"; same arithmetic as fv_weights / fv_weights_basis for table beams of order 0 / 1,
 * i.e. pyuvdata's az/za interpolation behind evaluate_beam, cpu/beams.py:12-89).
 * fv_tiles_sort: once per (time step, source chunk), after fv_rotate_cut: sorts the live set by the 16 x 16-cell
 *   tile of the beam grid each direction falls in and re-orders xyz / az / za / src_idx IN PLACE (stable: slots of
 *   a tile keep their catalogue order), so that slot == sorted position for everything downstream.
 * fv_weights_tiled: one CTA per (tile, group of frequencies) copies the tile's 17 x 17-point patch of every Jones
 *   entry into shared memory (cp.async.bulk rows + mbarrier, double-buffered over the frequencies) and evaluates
 *   all of the tile's sources from it.  basis = 0: the pair (beams[0], beams[K-1]), K = 1 or 2, out as fv_weights;
 *   basis = 1: all pairs k <= l of the K <= 6 beams, out as fv_weights_basis.
 * fv_tiles_supported: 1 when the K beams are tables of order 0 / 1 on one common grid. */
typedef struct fv_tiles fv_tiles;
int fv_tiles_create(fv_tiles** tiles, void* stream);
int fv_tiles_destroy(fv_tiles* tiles);
int fv_tiles_supported(const fv_beam* beams_host, int K);
int fv_tiles_sort(fv_tiles* tiles, int prec, const fv_beam* beam_host, void* xyz /* (3, n_cap) */, void* az, void* za,
                  int32_t* src_idx, const int32_t* n_dev, int64_t n_cap);
int fv_weights_tiled(fv_tiles* tiles, int prec, int mode, const fv_beam* beams_host, int K, int basis,
                     const void* az, const void* za, const int32_t* src_idx, int64_t n_cap, const double* freqs,
                     int nf, int64_t freq_index0, const void* flux, int64_t nsrc_total, void* out);

/* stand-alone apparent-coherency products on caller-supplied beam values: the four methods of
 * CPUBeamEvaluator (cpu/beams.py:129-246).  mode 1: A_i^H diag(F) A_j; mode 4: A_i^H C A_j;
 * mode 2: as 4 with both beams flipped along the vector axis (cpu_simulate.py:146-147,153).
 * beam_i / beam_j / coherency / out: (4, n) cplx rows [vec*2+feed]; flux: (n) cplx. */
int fv_coherency(int prec, int mode, const void* beam_i, const void* beam_j,
                 const void* flux_or_coh, int64_t n, void* out, void* stream);

/* ---- a6-a9: the NUFFT ----------------------------------------------------------------------- */
typedef struct fv_plan fv_plan; /* opaque: cuFFT plan cache + work grids, one per GPU/stream */
int fv_plan_create(fv_plan** plan, void* stream);
int fv_plan_destroy(fv_plan* plan);
/* per-stage device time: with timing enabled every stage launch inside fv_nufft2d1 / fv_nufft3 is
 * bracketed by a CUDA event pair on the plan's stream (north_star: "the inner uniform FFT ... is
 * timed separately"; bench.py's roofline block uses the spread / interp entries).
 * fv_plan_stage_ms returns the cumulative milliseconds and launch count of one stage since the
 * last reset (it synchronises on the outstanding events). */
typedef enum {
  FV_STAGE_ZERO = 0,    /* clearing the fine grid */
  FV_STAGE_SPREAD = 1,  /* spread kernel (type 1 and type 3 step 1) */
  FV_STAGE_FFT = 2,     /* cuFFT exec */
  FV_STAGE_GATHER = 3,  /* type 1 deconvolve + mode gather + epilogue */
  FV_STAGE_DECONV = 4,  /* type 3 deconvolve + zero-pad */
  FV_STAGE_INTERP = 5,  /* type 3 interpolation + post-phase + epilogue */
  FV_STAGE_COUNT = 6
} fv_stage;
int fv_plan_set_timing(fv_plan* plan, int enable);
int fv_plan_reset_timing(fv_plan* plan);
int fv_plan_stage_ms(fv_plan* plan, int stage, double* ms_host, int64_t* count_host);
/* bytes of device memory the plan currently holds (grids + cuFFT work areas) */
int64_t fv_plan_bytes(fv_plan* plan);

/* common output epilogue: value for (batch b, transform p, target k) goes to
 *   out[b*out_stride_b + pmap[p]*out_stride_p + (kmap ? kmap[k] : k)], conjugated first when
 *   conj_flag[k] (flipped baselines, cpu_simulate.py:298), added when accumulate != 0
 *   (cpu_simulate.py:1024,1069), stored otherwise. */
typedef struct {
  void* out;
  int64_t out_stride_b, out_stride_p;
  int32_t pmap[4];            /* e.g. {0,2,1,3}: the feed-axis swap of cpu_simulate.py:300 */
  const int32_t* kmap;        /* device, may be NULL */
  const uint8_t* conj_flag;   /* device, may be NULL */
  int32_t accumulate;
} fv_epilogue;

/* type 1, 2-D (cpu_nufft2d_type1 -> finufft.nufft2d1 modeord=1 + integer mode gather,
 * cpu/nufft.py:120-175), batched over nb frequencies that share the grid:
 *   x_s = fl(bx_s * scale[b]),  y_s = fl(by_s * scale[b])     (cpu_simulate.py:990-992)
 *   out[b,p,k] = sum_s W[b,p,s] exp(i (m1_k x_s + m2_k y_s))
 * W (nb, ntr, n_cap) cplx.  m1/m2 int32 signed mode numbers with |m| <= (n_modes-1)/2; a flipped
 * baseline passes (-m1,-m2) and conj_flag=1 (cpu_simulate.py:259,298). */
int fv_nufft2d1(fv_plan* plan, int prec, const void* bx, const void* by, const int32_t* n_dev,
                int64_t n_cap, const double* scale_host, int nb, int ntr, const void* W,
                int n_modes, const int32_t* m1, const int32_t* m2, int64_t nk, double eps,
                double upsampfac, const fv_epilogue* epi_host);

/* type 1, 2-D, fused shared-memory form (same transform, kernel, grid size and deconvolution as
 * fv_nufft2d1, i.e. finufft.nufft2d1 + the mode gather of cpu/nufft.py:120-175), for arrays whose
 * baselines are known up front: the fine grid lives only in shared memory (spread + FFT along x per
 * strip of grid rows; FFT along y + deconvolve + gather per group of needed columns; the inner FFT
 * is this library's own shared-memory mixed-radix transform, not cuFFT).
 * An fv_modeset holds the (m1, m2) integer modes of one beam pair's baselines (flip already
 * applied) bucketed by m1; it is built once on the host. */
typedef struct fv_modeset fv_modeset;
int fv_modeset_create(fv_modeset** modes, const int32_t* m1_host, const int32_t* m2_host, int64_t nk,
                      int n_modes);
int fv_modeset_destroy(fv_modeset* modes);
int fv_nufft2d1_fused(fv_plan* plan, int prec, const void* bx, const void* by, const int32_t* n_dev,
                      int64_t n_cap, const double* scale_host, int nb, int ntr, const void* W,
                      fv_modeset* modes, double eps, double upsampfac, const fv_epilogue* epi_host);
/* tuning knobs: "t1_rows" (strip height of the fused type-1 path, 0 = automatic), "t1_cols"
 * (columns per CTA of its second pass), "max_grid_bytes" */
int fv_plan_set_option(fv_plan* plan, const char* name, int64_t value);
/* Geometry of the plan's last type-3 transform, for the roofline of bench.py (finufft keeps the same
 * numbers inside its plan object: nf1..3, the inner type-2 plan's grid; cpu/nufft.py:48,105 never
 * sees them): out12 = {dim, w, nf[3] spread grid, ng[3] FFT grid, tiled spreader?, own pruned FFT?,
 * frequencies per sub-batch, ntr}. */
int fv_plan_last_geometry(fv_plan* plan, int64_t* out12_host);

/* type 3, 2-D / 3-D (cpu_nufft2d / cpu_nufft3d -> finufft.nufft2d3 / nufft3d3, cpu/nufft.py:11-118),
 * batched over nb frequencies:   s_k(b) = fl(base_k * scale[b])   (uvw = bls*freq, :973)
 *   out[b,p,k] = sum_s W[b,p,s] exp(i (u_k x_s + v_k y_s [+ w_k z_s]))
 * x/y/z (n_cap) real NU points (shared by the batch); u/v/w (nk) real per-unit-scale targets.
 * xlim_host: 2*dim doubles {min,max} of each NU coordinate over the live points, or NULL to have
 * the library reduce them on the device (one small D2H). */
int fv_nufft3(fv_plan* plan, int prec, int dim, const void* x, const void* y, const void* z,
              const int32_t* n_dev, int64_t n_cap, const double* xlim_host, const void* u,
              const void* v, const void* w, int64_t nk, const double* ulim_host,
              const double* scale_host, int nb, int ntr, const void* W, double eps,
              double upsampfac, const fv_epilogue* epi_host);

/* {min, max} of the first *n_dev (or n_fixed when n_dev is NULL) entries of each of `dim` device
 * arrays, returned to the host as lim_host[2*d], lim_host[2*d+1] (one small D2H + stream sync).
 * The engine calls it once per time step for the type-3 NU-point extents (finufft's arraywidcen
 * pass inside nufft2d3/nufft3d3, cpu/nufft.py:48,105); min > max means "no live entries". */
int fv_minmax(fv_plan* plan, int prec, int dim, const void* x, const void* y, const void* z,
              const int32_t* n_dev, int64_t n_fixed, double* lim_host);

/* on-GPU direct fp64-accumulated sum (validation aid and crossover baseline) */
int fv_direct_sum(int prec, int dim, const void* x, const void* y, const void* z,
                  const int32_t* n_dev, int64_t n_cap, const void* u, const void* v, const void* w,
                  int64_t nk, const double* scale_host, int nb, int ntr, const void* W,
                  const fv_epilogue* epi_host, void* stream);

/* ---- a10: beam-basis contraction (cpu_simulate.py:416-468) ------------------------------------
 * vis[b, :, :, k] += conj(c[a1_k, kk, f_b]) c[a2_k, ll, f_b] V_kl[b, :, :, k]
 *                 (+ conj(c[a1_k, ll]) c[a2_k, kk] V_kl^T  when ll != kk)
 * vkl: (nb, 4, nk) cplx, ALREADY in output feed order (written by a NUFFT call whose epilogue
 * used pmap {0,2,1,3}); coefs (nant, K, nfreq_total) cplx; result goes through the epilogue. */
int fv_basis_contract(int prec, const void* vkl, int nb, int64_t nk, const void* coefs,
                      int64_t nant, int K, int64_t nfreq_total, int64_t freq_index0, int kk, int ll,
                      const int32_t* ant1, const int32_t* ant2, const fv_epilogue* epi_host,
                      void* stream);

/* The same contraction over ALL K (K + 1) / 2 pairs in one pass: vkl (nb, npairs * 4, nk), pair q =
 * (kk, ll), kk <= ll in row-major order, at transforms 4 q .. 4 q + 3 -- the layout one batched NUFFT
 * call with ntr = 4 npairs writes (cpu_simulate.py:416-468, the double loop over k <= l). */
int fv_basis_contract_all(int prec, const void* vkl, int nb, int64_t nk, const void* coefs,
                          int64_t nant, int K, int64_t nfreq_total, int64_t freq_index0,
                          const int32_t* ant1, const int32_t* ant2, const fv_epilogue* epi_host,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFTVIS_B200_H */
