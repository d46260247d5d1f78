// Fused type-1 path: host side (mode-set tables, launch configuration) of the kernels in type1_fused.cuh.
#include "nufft_internal.cuh"

namespace fv {

template <typename T>
static int get_modeset_tables(fv_plan* P, fv_modeset* M, int prec, int64_t nf, int w, double beta,
                              const fv_plan::SmemFft& F, fv_modeset::Tables** out) {
  auto key = std::make_tuple(prec, nf, w, beta);
  auto it = M->tables.find(key);
  if (it != M->tables.end()) { *out = &it->second; return FV_OK; }
  const int64_t nk = (int64_t)M->m1.size();
  const int half = M->n_modes / 2;
  Quad Q = make_quad(w, beta);
  std::vector<double> ph = kernel_ft_series(nf, Q);
  // columns = sorted unique first mode numbers
  std::vector<int32_t> order(nk);
  for (int64_t k = 0; k < nk; ++k) order[k] = (int32_t)k;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return M->m1[a] < M->m1[b]; });
  std::vector<int32_t> col_pos, col_off, col_k, s_k(nk), s_pos(nk);
  std::vector<T> s_scale(nk), s_scale_y(nk);
  for (int64_t i = 0; i < nk; ++i) {
    const int32_t k = order[i];
    const int a1 = M->m1[k], a2 = M->m2[k];
    if (i == 0 || a1 != M->m1[order[i - 1]]) {
      col_off.push_back((int32_t)i);
      col_pos.push_back(F.pos[a1 < 0 ? a1 + nf : a1]);
      col_k.push_back(a1);
    }
    s_k[i] = k;
    s_pos[i] = F.pos[a2 < 0 ? a2 + nf : a2];
    s_scale[i] = (T)(1.0 / (ph[abs(a1)] * ph[abs(a2)]));
    s_scale_y[i] = (T)(1.0 / ph[abs(a2)]);
    (void)half;
  }
  col_off.push_back((int32_t)nk);
  fv_modeset::Tables t;
  t.ncols = (int)col_pos.size();
  auto up = [&](const void* src, size_t bytes, void** dst) -> int {
    FV_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
    if (bytes) FV_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, P->stream));
    return FV_OK;
  };
  int rc;
  if ((rc = up(col_pos.data(), col_pos.size() * 4, (void**)&t.col_pos))) return rc;
  if ((rc = up(col_off.data(), col_off.size() * 4, (void**)&t.col_off))) return rc;
  if ((rc = up(s_k.data(), s_k.size() * 4, (void**)&t.s_k))) return rc;
  if ((rc = up(s_pos.data(), s_pos.size() * 4, (void**)&t.s_pos))) return rc;
  if ((rc = up(s_scale.data(), s_scale.size() * sizeof(T), &t.s_scale))) return rc;
  if ((rc = up(col_k.data(), col_k.size() * 4, (void**)&t.col_k))) return rc;
  if ((rc = up(s_scale_y.data(), s_scale_y.size() * sizeof(T), &t.s_scale_y))) return rc;
  FV_CUDA(cudaStreamSynchronize(P->stream));
  auto res = M->tables.emplace(key, t);
  *out = &res.first->second;
  return FV_OK;
}

template <typename T, int WT, int NP>
static int launch_t1_spread(fv_plan* P, T1SpreadArgs<T>& a, dim3 grid, int threads, size_t smem) {
  auto kern = t1_spread_fftx_kernel<T, WT, NP>;
  FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, threads, smem, P->stream>>>(a);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

template <typename T>
static int nufft2d1_fused_impl(fv_plan* P, int prec, const void* bx, const void* by, const int32_t* n_dev,
                               int64_t n_cap, const double* scale, int nb, int ntr, const void* W,
                               fv_modeset* M, double eps, double upsampfac, const fv_epilogue* epi) {
  using C = cplx_t<T>;
  int w; double beta;
  kernel_params(eps, upsampfac, prec, &w, &beta);
  const int n_modes = M->n_modes;
  const int64_t nf = next235even(std::max<int64_t>((int64_t)(upsampfac * n_modes), 2 * w));
  fv_plan::SmemFft* F;
  int rc = get_smem_fft(P, prec, nf, &F);
  if (rc) return rc;
  fv_modeset::Tables* tab;
  rc = get_modeset_tables<T>(P, M, prec, nf, w, beta, *F, &tab);
  if (rc) return rc;
  const int ncols = tab->ncols;
  const int pitch = (int)nf + 1;
  const size_t tneed = sizeof(C) * (size_t)nb * ntr * (8 * (size_t)((ncols + 7) / 8)) * nf;   // blocks of 8 columns (x-direct)
  rc = ensure(&P->tbuf, &P->tbuf_bytes, tneed);
  if (rc) return rc;
  std::vector<BatchParams> bp(nb);
  for (int b = 0; b < nb; ++b) {
    bp[b] = BatchParams{};
    bp[b].smul = scale[b]; bp[b].tmul = 1.0;
    for (int d = 0; d < 3; ++d) bp[b].invgam[d] = 1.0;
  }
  rc = upload_bp(P, bp);
  if (rc) return rc;

  // ---- pass 1: spread + FFT along x ------------------------------------------------------------
  const int wmax = (w == 7 || w == 9 || w == 11 || w == 13 || w == 14) ? w : kMaxW;
  const int bank_mod = 128 / (int)sizeof(C);                 // elements per shared-memory bank sweep
  const size_t row_bytes = sizeof(C) * (nf + bank_mod);      // pass 1: upper bound of the padded row (pitch1 below)
  const size_t row_bytes2 = sizeof(C) * pitch;               // pass 2 column pitch
  const size_t smem_max = 227 * 1024 - 1024;
  // strip height R and CTA size: whole grid in one CTA when it fits; otherwise 16 rows x 512 threads
  // (one row per warp) if that fits, else 8 rows x 256 threads
  // strip height R: the whole grid when it fits one CTA, else as many rows as shared memory holds
  // (<= 32); one warp per strip row (256..768 threads): the row FFTs are warp tasks
  auto thr_for = [](int64_t rows) { return (int)std::min<int64_t>(t1_limits<T>::spread_threads, std::max<int64_t>(256, 32 * rows)); };
  int R;
  const int np = 1;   // products per CTA (the kernel also supports 4 per CTA; it measured slower and is not built)
  const bool whole = t1_spread_fixed_smem<T>((int)nf, wmax, thr_for(nf)) + row_bytes * nf <= 200 * 1024;
  if (P->t1_rows > 0) R = (int)std::min<int64_t>(P->t1_rows, nf);
  else if (whole) R = (int)nf;
  else {
    R = 32;
    while (R > 1 && t1_spread_fixed_smem<T>((int)nf, wmax, thr_for(R)) + row_bytes * R > smem_max) --R;
    if (R > 8) R -= R % 8;
  }
  int threads = thr_for(R);
  // Row pitch of the pass-1 strip: the spreader's lanes are (footprint row, column group) pairs that touch
  // cell (row * pitch + column group); with pitch = G (segment path) or w (row-block path) modulo one bank
  // sweep, the 32 lanes of one instruction fall in 32 / 16 distinct banks -- no conflicts (pitch = nf + 1
  // measured 3-4 wavefronts per access in the spreader).
  int pitch1, nseg;
  {
    const int nwarps = threads / 32, G = std::max(1, 32 / w);
    nseg = std::min(std::min(nwarps, T1_MAXSEG), (int)(nf / (4 * w)));
    const bool use_seg = nseg >= std::min(nwarps, 8);
    const int want = (use_seg ? G : w) % bank_mod;
    pitch1 = (int)nf + 1;
    while (pitch1 % bank_mod != want) ++pitch1;
  }
  size_t fixed1 = t1_spread_fixed_smem<T>((int)nf, wmax, threads, np);
  while (R > 1 && fixed1 + np * row_bytes * R > smem_max) --R;
  if (fixed1 + np * row_bytes * R > smem_max) { set_error("fine-grid row does not fit shared memory: use the cuFFT type-1 path"); return FV_ERR_UNSUPPORTED; }
  // x-direct pass 1 (type1_xdirect.cuh): single precision, a grid of several strips
  const bool use_xd = sizeof(T) == 4 && P->t1_xdirect && !whole && P->t1_rows == 0 && t1_xdirect_built(w) &&
                      nf >= 2 * (t1_xdirect_rows() + w) &&
                      ceil_div(nf, t1_xdirect_rows()) <= 1024;   // per-warp strip counters of the prep pass in shared memory
  if (use_xd) R = t1_xdirect_rows();
  // fold every (frequency, source) point once
  const size_t per = (size_t)nb * n_cap;
  const int nstrips = ceil_div(nf, R);
  const bool masks = nstrips > 1 && nstrips <= 64;           // strip membership of every source as two 32-bit masks
  if (!use_xd) rc = ensure(&P->prep, &P->prep_bytes, per * (4 * sizeof(int32_t) + 2 * sizeof(T)));
  if (rc) return rc;
  int32_t* ix0 = (int32_t*)P->prep;
  int32_t* iy0 = ix0 + per;
  uint32_t* hm0 = (uint32_t*)(iy0 + per);
  uint32_t* hm1 = hm0 + per;
  T* zx = (T*)(hm1 + per);
  T* zy = zx + per;
  if (!use_xd) {
    StageScope ts(P, FV_STAGE_ZERO);
    dim3 grid(ceil_div(n_cap, 256), nb);
    t1_prep_kernel<T><<<grid, 256, 0, P->stream>>>((const T*)bx, (const T*)by, n_dev, n_cap, P->bp_dev, (int)nf, w, ix0, iy0, zx, zy,
                                                   R, nstrips, masks ? hm0 : nullptr, masks ? hm1 : nullptr);
    FV_LAUNCH_CHECK();
  }
  // small grid held whole in one CTA and many sources: bins of sources + register windows (type1_small.cuh)
  bool small_ok = !use_xd && whole && R == (int)nf && t1s_width_built(w) && (int64_t)nb * n_cap < (1ll << 31) / 40;
  if (small_ok) {
    const int B = 4 * ((w + 4) / 4) - w + 1 < 2 ? 0 : [&] { for (int q = std::min(4 * ((w + 4) / 4) - w + 1, 6); q >= 2; --q) if (nf % q == 0) return q; return 0; }();
    const int64_t nbd = B ? nf / B : 0;
    small_ok = B >= 2 && nbd / ((w + 2 * B - 2) / B) >= 2 && nbd * nbd < 65536 &&
               sizeof(C) * ((size_t)nf * (nf + 1) + nf) + sizeof(int) * (nf + 4096 + nbd * nbd) + 2 * nbd * nbd +
                       4 * 2 * (8 * 32 * sizeof(T) + 8 * sizeof(C)) + 256 <= smem_max;
  }
  const bool use_small = small_ok && (P->t1_small == 2 || (P->t1_small == 1 && n_cap >= 4096));
  if (use_xd) {
    rc = t1_xdirect_pass1_entry(P, bx, by, n_dev, n_cap, nb, ntr, W, nf, w, beta, tab);
    if (rc) return rc;
  }
  if (use_small) {
    rc = t1_small_pass1_entry(P, prec, n_dev, n_cap, nb, ntr, W, nf, w, beta, ix0, iy0, zx, zy, F, tab);
    if (rc) return rc;
  }
  T1SpreadArgs<T> a{};
  a.n_dev = n_dev; a.n_cap = n_cap; a.ix0 = ix0; a.iy0 = iy0; a.zx = zx; a.zy = zy;
  a.hm0 = masks ? hm0 : nullptr; a.hm1 = masks ? hm1 : nullptr;
  a.nf = (int)nf; a.R = R; a.pitch = pitch1; a.w = w;
  a.nseg = nseg; a.seg = nseg > 0 ? (int)((nf + nseg - 1) / nseg) : (int)nf;
  a.inv_ntr = ntr > 1 ? 0xFFFFFFFFu / (unsigned)ntr + 1u : 0u;
  a.beta = (T)beta; a.c = (T)(4.0 / ((double)w * w)); a.halfw = (T)(w / 2.0);
  a.ntr = ntr; a.W = (const C*)W; a.tw = (const C*)F->tw; a.st = F->st;
  a.ncols = ncols; a.col_pos = tab->col_pos; a.Tbuf = (C*)P->tbuf;
  if (!use_small && !use_xd) {
    StageScope ts(P, FV_STAGE_SPREAD);
    dim3 grid(ceil_div(nf, R), np == 4 ? nb : nb * ntr);
    const size_t smem = fixed1 + np * sizeof(C) * pitch1 * R;
    static const bool dbg = getenv("FV_DEBUG") != nullptr;
    static long long* dbg_dev = nullptr;
    if (dbg) {
      if (!dbg_dev) { cudaMalloc((void**)&dbg_dev, 96); }
      cudaMemsetAsync(dbg_dev, 0, 96, P->stream);
      a.dbg = dbg_dev;
    }
    if (dbg) fprintf(stderr, "[fv] t1 fused: nf=%lld w=%d ncols=%d R=%d threads=%d smem=%zu grid=(%u,%u)\n",
                     (long long)nf, w, ncols, R, threads, smem, grid.x, grid.y);
    FV_DISPATCH_W(w, (rc = launch_t1_spread<T, WT, 1>(P, a, grid, threads, smem)));
    if (rc) return rc;
    if (dbg) {
      long long hcyc[12];
      cudaMemcpyAsync(hcyc, dbg_dev, 96, cudaMemcpyDeviceToHost, P->stream);
      cudaStreamSynchronize(P->stream);
      const double nw = 8.0 * (threads / 32);
      fprintf(stderr, "[fv] t1 pass-1 cycles/warp: zero %.0f scan %.0f fill %.0f spread %.0f fft %.0f fftwait %.0f write %.0f\n",
              hcyc[0] / nw, hcyc[1] / nw, hcyc[2] / nw, hcyc[3] / nw, hcyc[4] / nw, hcyc[5] / nw, hcyc[6] / nw);
    }
  }
  // ---- pass 2: FFT along y + deconvolve + gather -----------------------------------------------
  const size_t fixed2 = sizeof(C) * nf;
  const size_t gbudget = (t1_limits<T>::gather_blocks >= 3 ? 74 : 100) * 1024;   // shared memory per CTA: 3 or 2 CTAs per SM
  int cpc = P->t1_cols > 0 ? P->t1_cols : (int)std::max<size_t>(1, (gbudget - std::min<size_t>(fixed2, gbudget - 1024)) / row_bytes2);
  cpc = std::min(cpc, 16);
  if (cpc >= 8) cpc -= cpc % 8;
  cpc = std::min(cpc, ncols);
  while (cpc > 1 && fixed2 + row_bytes2 * cpc > smem_max) --cpc;
  T1GatherArgs<T> g{};
  g.Tbuf = (const C*)P->tbuf; g.nf = (int)nf; g.pitch = pitch; g.ncols = ncols; g.cols_per_cta = cpc; g.ntr = ntr;
  g.tw = (const C*)F->tw; g.st = F->st; g.col_off = tab->col_off; g.s_k = tab->s_k; g.s_pos = tab->s_pos;
  g.s_scale = (const T*)(use_xd ? tab->s_scale_y : tab->s_scale); g.epi = make_epi(epi);
  g.t_blocked = use_xd ? 1 : 0;
  // blocked T (x-direct pass 1): one CTA per block of 8 columns, transformed in the block's own layout
  const size_t smem8 = sizeof(float2) * (size_t)nf * 9 + 16;  // the block + twiddles (<= nf entries, rounded up to 16 bytes)
  if (use_xd && P->t1_cols == 0 && smem8 <= 74 * 1024) {
    if constexpr (sizeof(T) == 4) {
      StageScope ts(P, FV_STAGE_GATHER);
      auto kern8 = t1_ffty8_gather_kernel<T>;
      FV_CUDA(cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
      dim3 grid(ceil_div(ncols, 8), nb * ntr);
      kern8<<<grid, T1_THREADS, smem8, P->stream>>>(g);
      FV_LAUNCH_CHECK();
      return FV_OK;
    }
  }
  {
    StageScope ts(P, FV_STAGE_GATHER);
    auto kern = t1_ffty_gather_kernel<T>;
    const size_t smem = fixed2 + row_bytes2 * cpc;
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(ncols, cpc), nb * ntr);
    const int gthreads = std::min(T1_THREADS, std::max(64, 32 * cpc));     // one warp per column
    kern<<<grid, gthreads, smem, P->stream>>>(g);
    FV_LAUNCH_CHECK();
  }
  return FV_OK;
}


int nufft2d1_fused_entry(fv_plan* P, int prec, const void* bx, const void* by, const int32_t* n_dev, int64_t n_cap,
                         const double* scale, int nb, int ntr, const void* W, fv_modeset* M, double eps,
                         double upsampfac, const fv_epilogue* epi) {
  if (prec == 1) return nufft2d1_fused_impl<float>(P, prec, bx, by, n_dev, n_cap, scale, nb, ntr, W, M, eps, upsampfac, epi);
  return nufft2d1_fused_impl<double>(P, prec, bx, by, n_dev, n_cap, scale, nb, ntr, W, M, eps, upsampfac, epi);
}

}  // namespace fv
