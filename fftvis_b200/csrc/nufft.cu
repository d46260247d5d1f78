// Stages a6-a9 of the hot path: the non-uniform FFT, hand-written for sm_100a around a cuFFT plan.
//   type 1 (gridded arrays):  spread -> cuFFT -> deconvolve + integer-mode gather
//   type 3 (2-D / 3-D):       pre-phase + spread -> deconvolve + zero-pad -> cuFFT ->
//                             interpolate at the rescaled baselines + post-phase / deconvolve
// Replaces finufft.nufft2d1 / nufft2d3 / nufft3d3 as called from the reference
// (cpu/nufft.py:48,105,162) and the dispatch around them (_run_nufft, cpu_simulate.py:205-300).
// Not a port of finufft/cufinufft: the transform is *batched over frequency* -- all frequencies of
// a batch share the source set and differ only by a scalar on the coordinates -- so one launch
// covers (sources x frequencies), one cuFFT call covers (frequencies x polarisation products), and
// the batch is sized so the fine grids stay resident in B200's 126 MB L2 between the stages.
// Same kernel (exponential of semicircle), width/beta rules and grid sizes as the published
// algorithm, so the error behaves like the CPU backend's.
#include "nufft_internal.cuh"

namespace fv {
// defined in type1_fused.cu / type3.cu (one translation unit per path keeps the build parallel)
int nufft2d1_fused_entry(fv_plan* P, int prec, const void* bx, const void* by, const int32_t* n_dev, int64_t n_cap,
                         const double* scale, int nb, int ntr, const void* W, fv_modeset* M, double eps,
                         double upsampfac, const fv_epilogue* epi);
int nufft3_entry(fv_plan* P, int prec, int dim, const void* x, const void* y, const void* z, const int32_t* n_dev,
                 int64_t n_cap, const double* xlim_in, const void* u, const void* v, const void* wv, int64_t nk,
                 const double* ulim_in, const double* scale, int nb, int ntr, const void* W, double eps,
                 double upsampfac, const fv_epilogue* epi);
int minmax_entry(fv_plan* P, int prec, int dim, const void* x, const void* y, const void* z, const int32_t* n_dev,
                 int64_t n_fixed, double* lim_host);
}  // namespace fv

namespace fv {

// ---- type 1 ------------------------------------------------------------------------------------
template <typename T>
static int nufft2d1_impl(fv_plan* P, int prec, const void* bx, const void* by, const int32_t* n_dev,
                         int64_t n_cap, const double* scale, int nb, int ntr, const void* W,
                         int n_modes, const int32_t* m1, const int32_t* m2, int64_t nk, double eps,
                         double upsampfac, const fv_epilogue* epi) {
  using C = cplx_t<T>;
  int w; double beta;
  kernel_params(eps, upsampfac, prec, &w, &beta);
  const int64_t nf = next235even(std::max<int64_t>((int64_t)(upsampfac * n_modes), 2 * w));
  const size_t need = sizeof(C) * (size_t)nb * ntr * nf * nf;
  if (need > P->max_grid_bytes) { set_error("type-1 batch needs " + std::to_string(need) + " bytes of grid; reduce the frequency batch"); return FV_ERR_ALLOC; }
  int rc = ensure(&P->grid, &P->grid_bytes, need);
  if (rc) return rc;
  { StageScope ts(P, FV_STAGE_ZERO); FV_CUDA(cudaMemsetAsync(P->grid, 0, need, P->stream)); }
  std::vector<BatchParams> bp(nb);
  for (int b = 0; b < nb; ++b) {
    bp[b] = BatchParams{};
    bp[b].smul = scale[b];
    bp[b].tmul = 1.0;
    for (int d = 0; d < 3; ++d) { bp[b].invgam[d] = 1.0; }
  }
  rc = upload_bp(P, bp);
  if (rc) return rc;
  const T* invphi;
  rc = get_invphi<T>(P, prec, n_modes / 2 + 1, nf, w, beta, false, &invphi);
  if (rc) return rc;
  cufftHandle h;
  rc = get_fft(P, prec, 2, nf, nf, 1, (int64_t)nb * ntr, &h);
  if (rc) return rc;

  SpreadArgs<T> a{};
  a.x[0] = (const T*)bx; a.x[1] = (const T*)by; a.x[2] = nullptr;
  a.n_dev = n_dev; a.n_cap = n_cap;
  a.nf[0] = (int)nf; a.nf[1] = (int)nf; a.nf[2] = 1;
  a.w = w; a.beta = (T)beta; a.c = (T)(4.0 / ((double)w * w)); a.halfw = (T)(w / 2.0);
  a.ntr = ntr; a.prephase = 0; a.W = (const C*)W; a.grid = (C*)P->grid; a.bp = P->bp_dev;
  rc = launch_spread<T>(P, 2, a, nb);
  if (rc) return rc;
  rc = run_fft(P, h, prec, P->grid);
  if (rc) return rc;
  StageScope ts(P, FV_STAGE_GATHER);
  dim3 grid(ceil_div(nk, 256), nb * ntr);
  gather_modes_kernel<T><<<grid, 256, 0, P->stream>>>((const C*)P->grid, (int)nf, ntr, n_modes / 2, invphi, m1, m2, nk, make_epi(epi));
  FV_LAUNCH_CHECK();
  return FV_OK;
}


}  // namespace fv

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int fv_kernel_params(double eps, double upsampfac, int prec, int* w_host, double* beta_host) {
  FV_REQUIRE(w_host && beta_host, "null pointer");
  FV_REQUIRE(upsampfac > 1.0, "upsampfac must exceed 1");
  fv::kernel_params(eps, upsampfac, prec, w_host, beta_host);
  return FV_OK;
}

extern "C" int64_t fv_next235even(int64_t n) { return fv::next235even(n); }

extern "C" int fv_plan_create(fv_plan** plan, void* stream) {
  FV_REQUIRE(plan, "null pointer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    fv::set_error("no CUDA device: fftvis_b200 has no CPU fallback");
    return FV_ERR_NO_DEVICE;
  }
  *plan = new fv_plan();
  (*plan)->stream = (cudaStream_t)stream;
  return FV_OK;
}

extern "C" int fv_plan_destroy(fv_plan* P) {
  if (!P) return FV_OK;
  for (auto& kv : P->ffts) cufftDestroy(kv.second);
  for (auto& kv : P->invphi) cudaFree(kv.second);
  for (auto& ev : P->pending) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
  for (auto& ev : P->event_pool) cudaEventDestroy(ev);
  if (P->grid) cudaFree(P->grid);
  if (P->grid2) cudaFree(P->grid2);
  if (P->bp_dev) cudaFree(P->bp_dev);
  if (P->lim_dev) cudaFree(P->lim_dev);
  if (P->tbuf) cudaFree(P->tbuf);
  if (P->prep) cudaFree(P->prep);
  if (P->bins) cudaFree(P->bins);
  if (P->grid3) cudaFree(P->grid3);
  if (P->scan_tmp) cudaFree(P->scan_tmp);
  if (P->small) cudaFree(P->small);
  if (P->rec) cudaFree(P->rec);
  for (auto& kv : P->small_scheds) { cudaFree(kv.second.ph_off); cudaFree(kv.second.ph_bins); }
  for (auto& kv : P->smem_ffts) { cudaFree(kv.second.tw); if (kv.second.pos_dev) cudaFree(kv.second.pos_dev); if (kv.second.wn) cudaFree(kv.second.wn); }
  delete P;
  return FV_OK;
}

namespace fv {
static int drain_timing(fv_plan* P) {
  for (auto& ev : P->pending) {
    FV_CUDA(cudaEventSynchronize(ev.e1));
    float ms = 0;
    FV_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    P->stage_ms[ev.stage] += ms;
    P->stage_n[ev.stage] += 1;
    P->event_pool.push_back(ev.e0); P->event_pool.push_back(ev.e1);
  }
  P->pending.clear();
  return FV_OK;
}
}  // namespace fv

extern "C" int fv_plan_set_timing(fv_plan* P, int enable) {
  FV_REQUIRE(P, "null plan");
  P->timing = enable != 0;
  return FV_OK;
}

extern "C" int fv_plan_reset_timing(fv_plan* P) {
  FV_REQUIRE(P, "null plan");
  int rc = fv::drain_timing(P);
  if (rc) return rc;
  for (int i = 0; i < FV_STAGE_COUNT; ++i) { P->stage_ms[i] = 0.0; P->stage_n[i] = 0; }
  return FV_OK;
}

extern "C" int fv_plan_stage_ms(fv_plan* P, int stage, double* ms_host, int64_t* count_host) {
  FV_REQUIRE(P && ms_host && count_host, "null pointer");
  FV_REQUIRE(stage >= 0 && stage < FV_STAGE_COUNT, "unknown stage");
  int rc = fv::drain_timing(P);
  if (rc) return rc;
  *ms_host = P->stage_ms[stage];
  *count_host = P->stage_n[stage];
  return FV_OK;
}

extern "C" int64_t fv_plan_bytes(fv_plan* P) {
  if (!P) return 0;
  return (int64_t)(P->grid_bytes + P->grid2_bytes + P->tbuf_bytes + P->prep_bytes + P->bins_bytes + P->grid3_bytes + P->scan_tmp_bytes + P->fft_work_bytes + P->table_bytes + P->small_bytes + P->rec_bytes);
}

extern "C" int fv_nufft2d1(fv_plan* plan, int prec, const void* bx, const void* by, const int32_t* n_dev,
                           int64_t n_cap, const double* scale_host, int nb, int ntr, const void* W,
                           int n_modes, const int32_t* m1, const int32_t* m2, int64_t nk, double eps,
                           double upsampfac, const fv_epilogue* epi_host) {
  FV_REQUIRE(plan && bx && by && n_dev && scale_host && W && m1 && m2 && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(ntr >= 1 && ntr <= 4, "ntr must be 1..4");
  FV_REQUIRE(n_modes >= 1, "n_modes must be positive");
  FV_REQUIRE(nb >= 0 && (int64_t)nb * ntr <= 65535, "batch too large");
  FV_REQUIRE(upsampfac > 1.0 && eps > 0, "bad eps / upsampfac");
  if (nb == 0 || nk == 0) return FV_OK;
  if (prec == 1) return fv::nufft2d1_impl<float>(plan, prec, bx, by, n_dev, n_cap, scale_host, nb, ntr, W, n_modes, m1, m2, nk, eps, upsampfac, epi_host);
  return fv::nufft2d1_impl<double>(plan, prec, bx, by, n_dev, n_cap, scale_host, nb, ntr, W, n_modes, m1, m2, nk, eps, upsampfac, epi_host);
}


extern "C" int fv_modeset_create(fv_modeset** ms, const int32_t* m1_host, const int32_t* m2_host, int64_t nk,
                                 int n_modes) {
  FV_REQUIRE(ms && (nk == 0 || (m1_host && m2_host)), "null pointer");
  FV_REQUIRE(n_modes >= 1 && nk >= 0, "bad n_modes / nk");
  const int half = n_modes / 2;
  for (int64_t k = 0; k < nk; ++k)
    FV_REQUIRE(abs(m1_host[k]) <= half && abs(m2_host[k]) <= half, "mode number outside [-n_modes/2, n_modes/2]");
  fv_modeset* M = new fv_modeset();
  M->m1.assign(m1_host, m1_host + nk);
  M->m2.assign(m2_host, m2_host + nk);
  M->n_modes = n_modes;
  *ms = M;
  return FV_OK;
}

extern "C" int fv_modeset_destroy(fv_modeset* M) {
  if (!M) return FV_OK;
  for (auto& kv : M->tables) {
    cudaFree(kv.second.col_pos); cudaFree(kv.second.col_off); cudaFree(kv.second.s_k);
    cudaFree(kv.second.s_pos); cudaFree(kv.second.s_scale);
    cudaFree(kv.second.col_k); cudaFree(kv.second.s_scale_y);
  }
  delete M;
  return FV_OK;
}

extern "C" int fv_plan_last_geometry(fv_plan* P, int64_t* out12_host) {
  FV_REQUIRE(P && out12_host, "null pointer");
  for (int i = 0; i < 12; ++i) out12_host[i] = P->last_geo[i];
  return FV_OK;
}

extern "C" int fv_plan_set_option(fv_plan* P, const char* name, int64_t value) {
  FV_REQUIRE(P && name, "null pointer");
  const std::string n(name);
  if (n == "t1_rows") P->t1_rows = (int)value;
  else if (n == "t1_cols") P->t1_cols = (int)value;
  else if (n == "max_grid_bytes") P->max_grid_bytes = (size_t)value;
  else if (n == "t3_tiles") P->t3_tiles = (int)value;
  else if (n == "t1_small") P->t1_small = (int)value;
  else if (n == "t1_xdirect") P->t1_xdirect = (int)value;
  else if (n == "timing_mask") P->timing_mask = (int)value;
  else if (n == "t3_half") P->t3_half = (int)value;
  else if (n == "t3_minby") P->t3_minb[1] = (int)value;
  else if (n == "t3_minbz") P->t3_minb[2] = (int)value;
  else if (n == "t3_fft") P->t3_fft = (int)value;
  else if (n.rfind("t3_v", 0) == 0 && n.size() == 5 && n[4] >= 'x' && n[4] <= 'z') P->t3_v[n[4] - 'x'] = (int)value;
  else if (n.rfind("t3_thr", 0) == 0 && n.size() == 7 && n[6] >= 'x' && n[6] <= 'z') P->t3_thr[n[6] - 'x'] = (int)value;
  else { fv::set_error("unknown option " + n); return FV_ERR_INVALID; }
  return FV_OK;
}

extern "C" int fv_nufft2d1_fused(fv_plan* plan, int prec, const void* bx, const void* by, const int32_t* n_dev,
                                 int64_t n_cap, const double* scale_host, int nb, int ntr, const void* W,
                                 fv_modeset* modes, double eps, double upsampfac, const fv_epilogue* epi_host) {
  FV_REQUIRE(plan && bx && by && n_dev && scale_host && W && modes && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(ntr >= 1 && ntr <= 64, "ntr must be 1..64");
  FV_REQUIRE(nb >= 0 && (int64_t)nb * ntr <= 65535, "batch too large");
  FV_REQUIRE(upsampfac > 1.0 && eps > 0, "bad eps / upsampfac");
  if (nb == 0 || modes->m1.empty()) return FV_OK;
  return fv::nufft2d1_fused_entry(plan, prec, bx, by, n_dev, n_cap, scale_host, nb, ntr, W, modes, eps, upsampfac, epi_host);
}

extern "C" int fv_nufft3(fv_plan* plan, int prec, int dim, const void* x, const void* y, const void* z,
                         const int32_t* n_dev, int64_t n_cap, const double* xlim_host, const void* u,
                         const void* v, const void* w, int64_t nk, const double* ulim_host,
                         const double* scale_host, int nb, int ntr, const void* W, double eps,
                         double upsampfac, const fv_epilogue* epi_host) {
  FV_REQUIRE(plan && x && y && n_dev && u && v && scale_host && W && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  FV_REQUIRE(dim == 2 || (z && w), "3-D transform needs z and w");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(ntr >= 1 && ntr <= 4, "ntr must be 1..4");
  FV_REQUIRE(nb >= 0 && (int64_t)nb * ntr <= 65535, "batch too large");
  FV_REQUIRE(upsampfac > 1.0 && eps > 0, "bad eps / upsampfac");
  if (nb == 0 || nk == 0) return FV_OK;
  return fv::nufft3_entry(plan, prec, dim, x, y, z, n_dev, n_cap, xlim_host, u, v, w, nk, ulim_host, scale_host, nb, ntr, W, eps, upsampfac, epi_host);
}

extern "C" int fv_minmax(fv_plan* plan, int prec, int dim, const void* x, const void* y, const void* z,
                         const int32_t* n_dev, int64_t n_fixed, double* lim_host) {
  FV_REQUIRE(plan && x && lim_host, "null pointer");
  FV_REQUIRE(dim >= 1 && dim <= 3 && (dim < 2 || y) && (dim < 3 || z), "bad dim / arrays");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  return fv::minmax_entry(plan, prec, dim, x, y, z, n_dev, n_fixed, lim_host);
}

extern "C" int fv_direct_sum(int prec, int dim, const void* x, const void* y, const void* z,
                             const int32_t* n_dev, int64_t n_cap, const void* u, const void* v,
                             const void* w, int64_t nk, const double* scale_host, int nb, int ntr,
                             const void* W, const fv_epilogue* epi_host, void* stream) {
  FV_REQUIRE(x && y && n_dev && u && v && scale_host && W && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  FV_REQUIRE(dim == 2 || (z && w), "3-D sum needs z and w");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(ntr >= 1 && ntr <= 4 && nb >= 0 && nb <= 65535, "bad ntr / nb");
  if (nb == 0 || nk == 0) return FV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<fv::BatchParams> bp(nb);
  for (int b = 0; b < nb; ++b) { bp[b] = fv::BatchParams{}; bp[b].smul = 1.0; bp[b].tmul = scale_host[b]; }
  fv::BatchParams* bp_dev = nullptr;
  FV_CUDA(cudaMallocAsync((void**)&bp_dev, sizeof(fv::BatchParams) * nb, st));
  FV_CUDA(cudaMemcpyAsync(bp_dev, bp.data(), sizeof(fv::BatchParams) * nb, cudaMemcpyHostToDevice, st));
  fv::EpiDev ed = fv::make_epi(epi_host);
  dim3 grid(fv::ceil_div(nk, 128), nb);
  if (prec == 1) {
    if (dim == 2) fv::direct_sum_kernel<float, 2><<<grid, 128, 0, st>>>((const float*)x, (const float*)y, (const float*)z, n_dev, n_cap, (const float*)u, (const float*)v, (const float*)w, nk, bp_dev, ntr, (const float2*)W, ed);
    else fv::direct_sum_kernel<float, 3><<<grid, 128, 0, st>>>((const float*)x, (const float*)y, (const float*)z, n_dev, n_cap, (const float*)u, (const float*)v, (const float*)w, nk, bp_dev, ntr, (const float2*)W, ed);
  } else {
    if (dim == 2) fv::direct_sum_kernel<double, 2><<<grid, 128, 0, st>>>((const double*)x, (const double*)y, (const double*)z, n_dev, n_cap, (const double*)u, (const double*)v, (const double*)w, nk, bp_dev, ntr, (const double2*)W, ed);
    else fv::direct_sum_kernel<double, 3><<<grid, 128, 0, st>>>((const double*)x, (const double*)y, (const double*)z, n_dev, n_cap, (const double*)u, (const double*)v, (const double*)w, nk, bp_dev, ntr, (const double2*)W, ed);
  }
  FV_LAUNCH_CHECK();
  FV_CUDA(cudaFreeAsync(bp_dev, st));
  return FV_OK;
}

extern "C" int fv_basis_contract(int prec, const void* vkl, int nb, int64_t nk, const void* coefs,
                                 int64_t nant, int K, int64_t nfreq_total, int64_t freq_index0, int kk,
                                 int ll, const int32_t* ant1, const int32_t* ant2,
                                 const fv_epilogue* epi_host, void* stream) {
  FV_REQUIRE(vkl && coefs && ant1 && ant2 && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(kk >= 0 && ll >= kk && ll < K, "need 0 <= kk <= ll < K");
  FV_REQUIRE(nant > 0 && nb >= 0 && nb <= 65535, "bad nant / nb");
  if (nb == 0 || nk == 0) return FV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  fv::EpiDev ed = fv::make_epi(epi_host);
  dim3 grid(fv::ceil_div(nk, 256), nb);
  if (prec == 1) fv::basis_contract_kernel<float><<<grid, 256, 0, st>>>((const float2*)vkl, nk, (const float2*)coefs, K, nfreq_total, freq_index0, kk, ll, ant1, ant2, ed);
  else fv::basis_contract_kernel<double><<<grid, 256, 0, st>>>((const double2*)vkl, nk, (const double2*)coefs, K, nfreq_total, freq_index0, kk, ll, ant1, ant2, ed);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_basis_contract_all(int prec, const void* vkl, int nb, int64_t nk, const void* coefs,
                                     int64_t nant, int K, int64_t nfreq_total, int64_t freq_index0,
                                     const int32_t* ant1, const int32_t* ant2, const fv_epilogue* epi_host,
                                     void* stream) {
  FV_REQUIRE(vkl && coefs && ant1 && ant2 && epi_host && epi_host->out, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(K >= 1 && K <= 8, "1 <= K <= 8 basis beams");
  FV_REQUIRE(nant > 0 && nb >= 0 && nb <= 65535, "bad nant / nb");
  if (nb == 0 || nk == 0) return FV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  fv::EpiDev ed = fv::make_epi(epi_host);
  dim3 grid(fv::ceil_div(nk, 256), nb);
  if (prec == 1) fv::basis_contract_all_kernel<float><<<grid, 256, 0, st>>>((const float2*)vkl, nk, (const float2*)coefs, K, nfreq_total, freq_index0, ant1, ant2, ed);
  else fv::basis_contract_all_kernel<double><<<grid, 256, 0, st>>>((const double2*)vkl, nk, (const double2*)coefs, K, nfreq_total, freq_index0, ant1, ant2, ed);
  FV_LAUNCH_CHECK();
  return FV_OK;
}
