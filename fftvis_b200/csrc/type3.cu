// Type-3 path: host side (grid sizing, bin sort, pruned FFT passes) around the kernels of nufft_internal.cuh,
// type3_tiles.cuh and type3_fft.cuh.
#include "nufft_internal.cuh"
#include "type3_tiles.cuh"
#include "type3_fft.cuh"

namespace fv {

// ---- type 3 ------------------------------------------------------------------------------------
static void arraywidcen(double lo, double hi, double* w, double* c) {
  *w = (hi - lo) / 2.0; *c = (hi + lo) / 2.0;
  if (fabs(*c) < 0.1 * (*w)) { *w += fabs(*c); *c = 0.0; }
}

static void set_nhg_type3(double S, double X, double upsampfac, int w, int64_t* nf, double* h, double* gam) {
  double Xs = X, Ss = S;
  if (X == 0.0) { if (S == 0.0) { Xs = 1.0; Ss = 1.0; } else Xs = std::max(Xs, 1.0 / S); }
  else Ss = std::max(Ss, 1.0 / X);
  double nfd = 2.0 * upsampfac * Ss * Xs / M_PI + (w + 1);
  if (!std::isfinite(nfd)) nfd = 0.0;
  int64_t n = (int64_t)nfd;
  if (n < 2 * w) n = 2 * w;
  n = next235even(n);
  *nf = n; *h = 2.0 * M_PI / (double)n; *gam = (double)n / (2.0 * upsampfac * Ss);
}

template <typename T>
static int device_limits(fv_plan* P, const T* const* arr, int dim, const int32_t* n_dev, int64_t n_fixed, double* lim) {
  if (!P->lim_dev) FV_CUDA(cudaMalloc((void**)&P->lim_dev, sizeof(double) * 6));
  auto enc = [](double d) { long long i; memcpy(&i, &d, 8); return i >= 0 ? i : i ^ 0x7fffffffffffffffLL; };
  long long init[6];
  for (int d = 0; d < 3; ++d) { init[2 * d] = enc(INFINITY); init[2 * d + 1] = enc(-INFINITY); }
  FV_CUDA(cudaMemcpyAsync(P->lim_dev, init, sizeof(init), cudaMemcpyHostToDevice, P->stream));
  for (int d = 0; d < dim; ++d) {
    minmax_kernel<T><<<kNumSMs, 256, 0, P->stream>>>(arr[d], n_dev, n_fixed, P->lim_dev + 2 * d);
    FV_LAUNCH_CHECK();
  }
  long long raw[6];
  FV_CUDA(cudaMemcpyAsync(raw, P->lim_dev, sizeof(raw), cudaMemcpyDeviceToHost, P->stream));
  FV_CUDA(cudaStreamSynchronize(P->stream));
  for (int i = 0; i < 2 * dim; ++i) { long long v = raw[i] >= 0 ? raw[i] : raw[i] ^ 0x7fffffffffffffffLL; memcpy(&lim[i], &v, 8); }
  return FV_OK;
}

// half-length strided pass (y: d = 1, z: d = 2): vectors per CTA and CTAs per SM from the shared memory one CTA needs
template <typename T>
static int launch_half_strided(fv_plan* P, T3FftArgs<T> fa, fv_plan::SmemFft* Fh, int64_t nh, int d, int64_t ninner,
                               int64_t nouter, int q, int vfit) {
  using C = cplx_t<T>;
  int v = vfit;
  if (P->t3_v[d] > 0) v = std::min(v, P->t3_v[d]);
  if (v >= 4) v -= v % 2;
  int minb = P->t3_minb[d];
  if (minb == 0) minb = d == 2 ? 4 : 1;      // z: short vectors, occupancy-bound (cfg4: 26 -> 14 ms per 4 transforms); y: no gain measured
  // several CTAs per SM: each gets its share of shared memory
  if (minb > 1) {
    const size_t budget = (size_t)(220 * 1024) / minb;
    while (v > 1 && sizeof(C) * ((size_t)v * (nh + 1) + nh) > budget) --v;
  }
  fa.nvec_cta = v; fa.tw = (const C*)Fh->tw; fa.st = Fh->st; fa.pos = Fh->pos_dev;
  const size_t smem = sizeof(C) * ((size_t)v * (nh + 1) + nh);
  dim3 g((unsigned)(nouter * ceil_div(ninner, v)), q);
  const C* wn = (const C*)Fh->wn;
#define FV_T3H(MB, THR)                                                                                        \
  {                                                                                                            \
    auto kern = t3_fft_half_strided_kernel<T, MB>;                                                             \
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<g, THR, smem, P->stream>>>(fa, wn);                                                                 \
  }
  if (minb >= 4) FV_T3H(4, 256) else if (minb >= 2) FV_T3H(2, 256) else FV_T3H(1, (P->t3_thr[d] > 0 ? P->t3_thr[d] : 512))
#undef FV_T3H
  FV_LAUNCH_CHECK();
  return FV_OK;
}

template <typename T>
static int nufft3_impl(fv_plan* P, int prec, int dim, const void* x, const void* y, const void* z,
                       const int32_t* n_dev, int64_t n_cap, const double* xlim_in, const void* u,
                       const void* v, const void* wv, int64_t nk, const double* ulim_in,
                       const double* scale, int nb, int ntr, const void* W, double eps,
                       double upsampfac, const fv_epilogue* epi) {
  using C = cplx_t<T>;
  int w; double beta;
  kernel_params(eps, upsampfac, prec, &w, &beta);
  const T* xs[3] = {(const T*)x, (const T*)y, (const T*)z};
  const T* us[3] = {(const T*)u, (const T*)v, (const T*)wv};
  double xlim[6], ulim[6];
  int rc;
  if (xlim_in) memcpy(xlim, xlim_in, sizeof(double) * 2 * dim);
  else { rc = device_limits<T>(P, xs, dim, n_dev, 0, xlim); if (rc) return rc; }
  if (ulim_in) memcpy(ulim, ulim_in, sizeof(double) * 2 * dim);
  else { rc = device_limits<T>(P, us, dim, nullptr, nk, ulim); if (rc) return rc; }
  EpiDev ed = make_epi(epi);
  if (!(xlim[0] <= xlim[1])) {
    // no live sources: the transform is identically zero
    if (!epi->accumulate) {
      // the direct kernel stores zeros through the epilogue map when n == 0
      std::vector<BatchParams> bp(nb);
      for (int b = 0; b < nb; ++b) { bp[b] = BatchParams{}; bp[b].smul = 1.0; bp[b].tmul = scale[b]; }
      rc = upload_bp(P, bp); if (rc) return rc;
      dim3 grid(ceil_div(nk, 128), nb);
      if (dim == 2) direct_sum_kernel<T, 2><<<grid, 128, 0, P->stream>>>(xs[0], xs[1], xs[2], n_dev, n_cap, us[0], us[1], us[2], nk, P->bp_dev, ntr, (const C*)W, ed);
      else direct_sum_kernel<T, 3><<<grid, 128, 0, P->stream>>>(xs[0], xs[1], xs[2], n_dev, n_cap, us[0], us[1], us[2], nk, P->bp_dev, ntr, (const C*)W, ed);
      FV_LAUNCH_CHECK();
    }
    return FV_OK;
  }
  double X[3], Cc[3];
  for (int d = 0; d < dim; ++d) arraywidcen(xlim[2 * d], xlim[2 * d + 1], &X[d], &Cc[d]);

  // One grid shape for the whole batch: the frequencies of a batch differ by a few per cent, so the
  // grid is sized (finufft's set_nhg_type3 rule) for the widest target extent of the batch and every
  // frequency uses that rescaling.  Smaller extents only sit further inside the kernel's accurate
  // range; the sources then fall on the SAME cells for every frequency (one bin sort per batch).
  std::vector<BatchParams> bp(nb);
  bool prephase = false, postphase = false;
  double Smax[3] = {0, 0, 0};
  std::vector<double> Dv(3 * (size_t)nb, 0.0);
  for (int b = 0; b < nb; ++b)
    for (int d = 0; d < dim; ++d) {
      // targets are fl(base * scale) in working precision; min/max commute with that (monotone)
      const double lo = (double)((T)ulim[2 * d] * (T)scale[b]), hi = (double)((T)ulim[2 * d + 1] * (T)scale[b]);
      double S, D;
      arraywidcen(std::min(lo, hi), std::max(lo, hi), &S, &D);
      Smax[d] = std::max(Smax[d], S);
      Dv[3 * (size_t)b + d] = D;
    }
  int64_t nf[3] = {1, 1, 1};
  double hh[3] = {0, 0, 0}, gam[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) set_nhg_type3(Smax[d], X[d], upsampfac, w, &nf[d], &hh[d], &gam[d]);
  for (int b = 0; b < nb; ++b) {
    bp[b] = BatchParams{};
    bp[b].smul = 1.0;          // type 3 scales the targets (uvw = bls * freq), not the sources
    bp[b].tmul = scale[b];
    for (int d = 0; d < 3; ++d) bp[b].invgam[d] = 1.0;
    for (int d = 0; d < dim; ++d) {
      bp[b].C[d] = Cc[d]; bp[b].invgam[d] = 1.0 / gam[d]; bp[b].D[d] = Dv[3 * (size_t)b + d]; bp[b].hgam[d] = hh[d] * gam[d];
      if (bp[b].D[d] != 0.0) prephase = true;
      if (Cc[d] != 0.0) postphase = true;
    }
  }
  rc = upload_bp(P, bp);
  if (rc) return rc;
  const Quad Q = make_quad(w, beta);
  int64_t ng[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) ng[d] = next235even(std::max<int64_t>((int64_t)(upsampfac * nf[d]), 2 * w));
  const size_t cells1 = (size_t)nf[0] * nf[1] * nf[2], cells2 = (size_t)ng[0] * ng[1] * ng[2];
  const size_t per_b = sizeof(C) * ntr * (cells1 + cells2 + (dim == 3 ? (size_t)nf[2] * ng[1] * ng[0] : (size_t)nf[1] * ng[0]));
  if (per_b > P->max_grid_bytes) { set_error("a single type-3 transform needs " + std::to_string(per_b) + " bytes of grids"); return FV_ERR_ALLOC; }
  const int sub_max = (int)std::min<size_t>(nb, std::max<size_t>(1, P->max_grid_bytes / per_b));

  // thin 3-D grids: bin-sort the sources into column tiles once for the whole batch
  const bool tiled = dim == 3 && P->t3_tiles && nf[2] <= T3_NZMAX && n_cap > 0;
  T3Geom<T> geo{};
  int32_t *bin_counts = nullptr, *bin_offsets = nullptr, *bin_cursor = nullptr, *bin_list = nullptr;
  T* t3_rows = nullptr; int32_t* t3_cells = nullptr;
  int ntiles = 0;
  if (tiled) {
    geo.x = xs[0]; geo.y = xs[1]; geo.z = xs[2]; geo.n_dev = n_dev; geo.w = w;
    for (int d = 0; d < 3; ++d) { geo.C[d] = Cc[d]; geo.invgam[d] = 1.0 / gam[d]; geo.nf[d] = (int)nf[d]; }
    geo.ntx = ceil_div(nf[0], T3_TILE); geo.nty = ceil_div(nf[1], T3_TILE);
    ntiles = geo.ntx * geo.nty;
    const size_t nt1 = (size_t)ntiles + 1;
    const size_t lcap = 16 * (size_t)n_cap;                     // <= 4 x 4 tiles per source
    const size_t need = sizeof(int32_t) * (3 * nt1 + 2 * lcap);
    rc = ensure(&P->bins, &P->bins_bytes, need); if (rc) return rc;
    bin_counts = (int32_t*)P->bins; bin_offsets = bin_counts + nt1; bin_cursor = bin_offsets + nt1; bin_list = bin_cursor + nt1;
    StageScope ts(P, FV_STAGE_ZERO);
    FV_CUDA(cudaMemsetAsync(bin_counts, 0, sizeof(int32_t) * 3 * nt1, P->stream));
    const int blocks = ceil_div(n_cap, 256);
    t3_bin_kernel<T, 0><<<blocks, 256, 0, P->stream>>>(geo, bin_counts, nullptr, nullptr, nullptr);
    FV_LAUNCH_CHECK();
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, bin_counts, bin_offsets, (int)nt1, P->stream);
    rc = ensure(&P->scan_tmp, &P->scan_tmp_bytes, std::max<size_t>(tmp, 16)); if (rc) return rc;
    FV_CUDA(cub::DeviceScan::ExclusiveSum(P->scan_tmp, tmp, bin_counts, bin_offsets, (int)nt1, P->stream));
    ++fv::g_launches;
    t3_bin_kernel<T, 1><<<blocks, 256, 0, P->stream>>>(geo, nullptr, bin_offsets, bin_cursor, bin_list);
    FV_LAUNCH_CHECK();
    // ascending source index inside every tile: a fixed summation order (the atomic cursors above are served
    // in any order)
    int32_t* sorted = bin_list + lcap;
    const int nitems = (int)std::min<size_t>(lcap, (size_t)INT_MAX);
    tmp = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, tmp, bin_list, sorted, nitems, ntiles, bin_offsets, bin_offsets + 1, P->stream);
    rc = ensure(&P->scan_tmp, &P->scan_tmp_bytes, std::max<size_t>(tmp, 16)); if (rc) return rc;
    FV_CUDA(cub::DeviceSegmentedSort::SortKeys(P->scan_tmp, tmp, bin_list, sorted, nitems, ntiles, bin_offsets, bin_offsets + 1, P->stream));
    ++fv::g_launches;
    bin_list = sorted;
    // kernel rows + first cells of every live source (shared by every frequency and product of the batch)
    const int wmax_rec = (w == 7 || w == 9 || w == 11 || w == 13 || w == 14) ? w : kMaxW;
    rc = ensure(&P->rec, &P->rec_bytes, (size_t)n_cap * 3 * (wmax_rec * sizeof(T) + sizeof(int32_t)) + 64);
    if (rc) return rc;
    t3_rows = (T*)P->rec;
    t3_cells = (int32_t*)(t3_rows + (size_t)n_cap * 3 * wmax_rec);
    const unsigned rblocks = (unsigned)ceil_div(3 * n_cap, 256);
    FV_DISPATCH_W(w, (t3_records_kernel<T, WT><<<rblocks, 256, 0, P->stream>>>(geo, (T)beta, (T)(4.0 / ((double)w * w)),
                                                                              (T)(w / 2.0), t3_rows, t3_cells)));
    FV_LAUNCH_CHECK();
  }

  // own pruned FFT passes when every padded dimension's vectors fit shared memory
  const size_t smem_fft_max = 200 * 1024;
  auto vec_fit = [&](int64_t n, int cap) {
    const int64_t room = (int64_t)smem_fft_max / (int64_t)sizeof(C) - n;      // minus the twiddle table
    return (int)std::max<int64_t>(0, std::min<int64_t>(cap, room / (n + 1)));
  };
  int vx = vec_fit(ng[0], 8), vy = vec_fit(ng[1], 16), vz = dim == 3 ? vec_fit(ng[2], 64) : 1;
  if (vy >= 4) vy -= vy % 4;
  if (vz >= 4) vz -= vz % 4;
  if (P->t3_v[0] > 0) vx = std::min(vx, P->t3_v[0]);
  if (P->t3_v[1] > 0) vy = std::min(vy, P->t3_v[1]);
  if (P->t3_v[2] > 0) vz = std::min(vz, P->t3_v[2]);
  const int thx = P->t3_thr[0] > 0 ? P->t3_thr[0] : 512, thy = P->t3_thr[1] > 0 ? P->t3_thr[1] : 512,
            thz = P->t3_thr[2] > 0 ? P->t3_thr[2] : 256;
  // automatic choice: own pruned passes for 3-D grids, except in single precision below sigma = 2, where the
  // deconvolution factors span > 1e4 and the three separately rounded passes measured ~2x the error of one
  // cuFFT transform (tools/parity_probe.py); every BASELINE config runs at sigma = 2
  const bool lowsig32 = prec == 1 && upsampfac < 2.0;
  bool own_fft = (P->t3_fft == 2 || (P->t3_fft == 1 && dim == 3 && !lowsig32)) && vx >= 1 && vy >= 1 && vz >= 1;
  fv_plan::SmemFft* F[3] = {nullptr, nullptr, nullptr};
  if (own_fft) {
    for (int d = 0; d < dim && own_fft; ++d) {
      if (get_smem_fft(P, prec, ng[d], &F[d]) != FV_OK) own_fft = false;     // not 2-3-5 smooth etc.
      else if (!F[d]->pos_dev) {
        FV_CUDA(cudaMalloc((void**)&F[d]->pos_dev, sizeof(int) * ng[d]));
        FV_CUDA(cudaMemcpyAsync(F[d]->pos_dev, F[d]->pos.data(), sizeof(int) * ng[d], cudaMemcpyHostToDevice, P->stream));
        FV_CUDA(cudaStreamSynchronize(P->stream));
      }
    }
  }
  // half-length passes (type3_fft.cuh): ng = 2 nf with nf even, and a shared-memory plan for nf
  fv_plan::SmemFft* Fh[3] = {nullptr, nullptr, nullptr};
  bool halfd[3] = {false, false, false};
  if (own_fft && P->t3_half) {
    for (int d = 0; d < dim; ++d) {
      if (!((P->t3_half >> d) & 1) || ng[d] != 2 * nf[d] || (nf[d] & 1)) continue;
      if (get_smem_fft(P, prec, nf[d], &Fh[d]) != FV_OK) { Fh[d] = nullptr; continue; }
      if (!Fh[d]->pos_dev) {
        FV_CUDA(cudaMalloc((void**)&Fh[d]->pos_dev, sizeof(int) * nf[d]));
        FV_CUDA(cudaMemcpyAsync(Fh[d]->pos_dev, Fh[d]->pos.data(), sizeof(int) * nf[d], cudaMemcpyHostToDevice, P->stream));
        FV_CUDA(cudaStreamSynchronize(P->stream));
      }
      if (!Fh[d]->wn) {
        std::vector<C> h(nf[d]);
        for (int64_t j = 0; j < nf[d]; ++j) {
          const double ang = M_PI * (double)j / (double)nf[d];                 // 2 pi j / (2 nf)
          h[j] = make_c<T>((T)cos(ang), (T)sin(ang));
        }
        FV_CUDA(cudaMalloc(&Fh[d]->wn, sizeof(C) * nf[d]));
        FV_CUDA(cudaMemcpyAsync(Fh[d]->wn, h.data(), sizeof(C) * nf[d], cudaMemcpyHostToDevice, P->stream));
        FV_CUDA(cudaStreamSynchronize(P->stream));
      }
      halfd[d] = true;
    }
  }
  auto vec_fit_half = [&](int64_t nh, int cap) {
    const int64_t room = (int64_t)smem_fft_max / (int64_t)sizeof(C) - nh;
    int v = (int)std::max<int64_t>(0, std::min<int64_t>(cap, room / (nh + 1)));
    while (v > 1 && (int64_t)v * nh >= 65536) --v;                            // exact 32-bit index division in the kernels
    return v;
  };
  const size_t cells3 = dim == 3 ? (size_t)nf[2] * ng[1] * ng[0] : (size_t)nf[1] * ng[0];
  {
    int64_t* g = P->last_geo;
    g[0] = dim; g[1] = w;
    for (int d = 0; d < 3; ++d) { g[2 + d] = nf[d]; g[5 + d] = ng[d]; }
    g[8] = tiled ? 1 : 0; g[9] = own_fft ? 1 : 0; g[10] = sub_max; g[11] = ntr;
  }

  int b0 = 0;
  while (b0 < nb) {
    const int sub = std::min(sub_max, nb - b0);
    const int b1 = b0 + sub;
    rc = ensure(&P->grid, &P->grid_bytes, sizeof(C) * sub_max * ntr * cells1); if (rc) return rc;
    rc = ensure(&P->grid2, &P->grid2_bytes, sizeof(C) * sub_max * ntr * cells2); if (rc) return rc;
    if (own_fft) { rc = ensure(&P->grid3, &P->grid3_bytes, sizeof(C) * sub_max * ntr * cells3); if (rc) return rc; }
    if (tiled) {
      T3SpreadArgs<T> ta{};
      ta.g = geo; ta.n_cap = n_cap; ta.beta = (T)beta; ta.c = (T)(4.0 / ((double)w * w)); ta.halfw = (T)(w / 2.0);
      ta.ntr = ntr; ta.prephase = prephase ? 1 : 0;
      ta.W = (const C*)W + (int64_t)b0 * ntr * n_cap; ta.bp = P->bp_dev + b0;
      ta.offsets = bin_offsets; ta.list = bin_list; ta.rows = t3_rows; ta.cells = t3_cells; ta.grid = (C*)P->grid;
      StageScope ts(P, FV_STAGE_SPREAD);
      dim3 tg(ntiles, sub * ntr);
      const size_t ssm = t3_spread_smem<T>((w == 7 || w == 9 || w == 11 || w == 13 || w == 14) ? w : kMaxW);
      FV_DISPATCH_W(w, (cudaFuncSetAttribute(t3_col_spread_kernel<T, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm),
                        t3_col_spread_kernel<T, WT><<<tg, T3_TILE * T3_TILE, ssm, P->stream>>>(ta)));
      FV_LAUNCH_CHECK();
    } else {
      StageScope ts(P, FV_STAGE_ZERO);
      FV_CUDA(cudaMemsetAsync(P->grid, 0, sizeof(C) * sub * ntr * cells1, P->stream));
    }

    SpreadArgs<T> a{};
    for (int d = 0; d < 3; ++d) { a.x[d] = xs[d]; a.nf[d] = (int)nf[d]; }
    a.n_dev = n_dev; a.n_cap = n_cap; a.w = w; a.beta = (T)beta; a.c = (T)(4.0 / ((double)w * w)); a.halfw = (T)(w / 2.0);
    a.ntr = ntr; a.prephase = prephase ? 1 : 0;
    a.W = (const C*)W + (int64_t)b0 * ntr * n_cap; a.grid = (C*)P->grid; a.bp = P->bp_dev + b0;
    if (!tiled) { rc = launch_spread<T>(P, dim, a, sub); if (rc) return rc; }

    const T *inv1, *inv2, *inv3 = nullptr;
    rc = get_invphi<T>(P, prec, nf[0], ng[0], w, beta, true, &inv1); if (rc) return rc;
    rc = get_invphi<T>(P, prec, nf[1], ng[1], w, beta, true, &inv2); if (rc) return rc;
    if (dim == 3) { rc = get_invphi<T>(P, prec, nf[2], ng[2], w, beta, true, &inv3); if (rc) return rc; }
    if (own_fft) {
      // pruned inner FFT: deconvolve + transform x on the non-zero rows, then y, then z
      StageScope ts(P, FV_STAGE_FFT);
      const int q = sub * ntr;
      C* A1 = dim == 3 ? (C*)P->grid2 : (C*)P->grid3;
      {
        T3FftArgs<T> fa{};
        fa.in = (const C*)P->grid; fa.out = A1; fa.nin = (int)nf[0]; fa.n = (int)ng[0]; fa.nvec_cta = vx;
        fa.nvec = nf[1] * nf[2]; fa.in_q = (int64_t)cells1; fa.out_q = dim == 3 ? (int64_t)cells2 : nf[1] * ng[0];
        fa.inv1 = inv1; fa.inv2 = inv2; fa.inv3 = inv3; fa.nf2 = (int)nf[1];
        fa.tw = (const C*)F[0]->tw; fa.st = F[0]->st; fa.pos = F[0]->pos_dev;
        if (halfd[0]) {
          int v = vec_fit_half(nf[0], 16);
          if (P->t3_v[0] > 0) v = std::min(v, P->t3_v[0]);
          fa.nvec_cta = v; fa.tw = (const C*)Fh[0]->tw; fa.st = Fh[0]->st; fa.pos = Fh[0]->pos_dev;
          const size_t smem = sizeof(C) * ((size_t)v * (nf[0] + 1) + nf[0]);
          FV_CUDA(cudaFuncSetAttribute(t3_fft_half_contig_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          dim3 g(ceil_div(fa.nvec, v), q);
          t3_fft_half_contig_kernel<T><<<g, thx, smem, P->stream>>>(fa, (const C*)Fh[0]->wn);
          FV_LAUNCH_CHECK();
        } else {
        const size_t smem = sizeof(C) * ((size_t)vx * (ng[0] + 1) + ng[0]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_contig_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g(ceil_div(fa.nvec, vx), q);
        t3_fft_contig_kernel<T><<<g, thx, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
        }
      }
      {
        T3FftArgs<T> fa{};
        fa.in = A1; fa.out = dim == 3 ? (C*)P->grid3 : (C*)P->grid2; fa.nin = (int)nf[1]; fa.n = (int)ng[1]; fa.nvec_cta = vy;
        fa.ninner = (int)ng[0]; fa.nouter = (int)nf[2];
        fa.in_q = dim == 3 ? (int64_t)cells2 : nf[1] * ng[0]; fa.in_a = nf[1] * ng[0]; fa.in_k = ng[0];
        fa.out_q = dim == 3 ? nf[2] * ng[1] * ng[0] : (int64_t)cells2; fa.out_a = ng[1] * ng[0]; fa.out_k = ng[0];
        fa.tw = (const C*)F[1]->tw; fa.st = F[1]->st; fa.pos = F[1]->pos_dev;
        if (halfd[1]) {
          rc = launch_half_strided<T>(P, fa, Fh[1], nf[1], 1, (int64_t)ng[0], (int64_t)nf[2], q, vec_fit_half(nf[1], 16));
          if (rc) return rc;
        } else {
        const size_t smem = sizeof(C) * ((size_t)vy * (ng[1] + 1) + ng[1]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g((unsigned)(nf[2] * ceil_div(ng[0], vy)), q);
        t3_fft_strided_kernel<T><<<g, thy, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
        }
      }
      if (dim == 3) {
        T3FftArgs<T> fa{};
        fa.in = (const C*)P->grid3; fa.out = (C*)P->grid2; fa.nin = (int)nf[2]; fa.n = (int)ng[2]; fa.nvec_cta = vz;
        fa.ninner = (int)(ng[1] * ng[0]); fa.nouter = 1;
        fa.in_q = nf[2] * ng[1] * ng[0]; fa.in_a = 0; fa.in_k = ng[1] * ng[0];
        fa.out_q = (int64_t)cells2; fa.out_a = 0; fa.out_k = ng[1] * ng[0];
        fa.tw = (const C*)F[2]->tw; fa.st = F[2]->st; fa.pos = F[2]->pos_dev;
        if (halfd[2]) {
          rc = launch_half_strided<T>(P, fa, Fh[2], nf[2], 2, (int64_t)(ng[1] * ng[0]), 1, q, vec_fit_half(nf[2], 64));
          if (rc) return rc;
        } else {
        const size_t smem = sizeof(C) * ((size_t)vz * (ng[2] + 1) + ng[2]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g((unsigned)ceil_div(ng[1] * ng[0], vz), q);
        t3_fft_strided_kernel<T><<<g, thz, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
        }
      }
    } else {
      {
        StageScope ts(P, FV_STAGE_DECONV);
        dim3 g2(ceil_div((int64_t)cells2, 256), sub * ntr);
        if (dim == 2) deconv_pad_kernel<T, 2><<<g2, 256, 0, P->stream>>>((const C*)P->grid, (C*)P->grid2, (int)nf[0], (int)nf[1], 1, (int)ng[0], (int)ng[1], 1, inv1, inv2, inv3);
        else deconv_pad_kernel<T, 3><<<g2, 256, 0, P->stream>>>((const C*)P->grid, (C*)P->grid2, (int)nf[0], (int)nf[1], (int)nf[2], (int)ng[0], (int)ng[1], (int)ng[2], inv1, inv2, inv3);
        FV_LAUNCH_CHECK();
      }
      cufftHandle h;
      rc = get_fft(P, prec, dim, ng[0], ng[1], ng[2], (int64_t)sub * ntr, &h); if (rc) return rc;
      rc = run_fft(P, h, prec, P->grid2); if (rc) return rc;
    }

    InterpArgs<T> ia{};
    for (int d = 0; d < 3; ++d) { ia.u[d] = us[d]; ia.ng[d] = (int)ng[d]; }
    ia.nk = nk; ia.w = w; ia.beta = a.beta; ia.c = a.c; ia.halfw = a.halfw; ia.ntr = ntr;
    ia.postphase = postphase ? 1 : 0; ia.fw2 = (const C*)P->grid2; ia.bp = P->bp_dev + b0; ia.quad = Q;
    ia.epi = ed;
    ia.epi.out = (C*)ed.out + (int64_t)b0 * ed.sb;
    rc = launch_interp<T>(P, dim, ia, sub); if (rc) return rc;
    b0 = b1;
  }
  return FV_OK;
}


int nufft3_entry(fv_plan* P, int prec, int dim, const void* x, const void* y, const void* z, const int32_t* n_dev,
                 int64_t n_cap, const double* xlim_in, const void* u, const void* v, const void* wv, int64_t nk,
                 const double* ulim_in, const double* scale, int nb, int ntr, const void* W, double eps,
                 double upsampfac, const fv_epilogue* epi) {
  if (prec == 1) return nufft3_impl<float>(P, prec, dim, x, y, z, n_dev, n_cap, xlim_in, u, v, wv, nk, ulim_in, scale, nb, ntr, W, eps, upsampfac, epi);
  return nufft3_impl<double>(P, prec, dim, x, y, z, n_dev, n_cap, xlim_in, u, v, wv, nk, ulim_in, scale, nb, ntr, W, eps, upsampfac, epi);
}

int minmax_entry(fv_plan* P, int prec, int dim, const void* x, const void* y, const void* z, const int32_t* n_dev,
                 int64_t n_fixed, double* lim_host) {
  if (prec == 1) {
    const float* a[3] = {(const float*)x, (const float*)y, (const float*)z};
    return device_limits<float>(P, a, dim, n_dev, n_fixed, lim_host);
  }
  const double* a[3] = {(const double*)x, (const double*)y, (const double*)z};
  return device_limits<double>(P, a, dim, n_dev, n_fixed, lim_host);
}

}  // namespace fv
