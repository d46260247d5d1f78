// Small-grid type-1 path: host side (bin sort, phase schedule, launches) of the kernels in type1_small.cuh.
#include "nufft_internal.cuh"
#include "type1_small.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace fv {

// ---- small-grid path (type1_small.cuh) ------------------------------------------------------------
bool t1s_width_built(int w) { return w == 7 || w == 9 || w == 11 || w == 12 || w == 13 || w == 14; }

#define FV_DISPATCH_WS(WV, CALL)                        \
  switch (WV) {                                         \
    case 7: { constexpr int WT = 7; CALL; } break;      \
    case 9: { constexpr int WT = 9; CALL; } break;      \
    case 11: { constexpr int WT = 11; CALL; } break;    \
    case 12: { constexpr int WT = 12; CALL; } break;    \
    case 13: { constexpr int WT = 13; CALL; } break;    \
    default: { constexpr int WT = 14; CALL; } break;    \
  }

// Phases of bins with pairwise disjoint (w + B - 1)-cell windows on the periodic grid.  Per dimension the
// nbd = nf / B bins are split into classes whose members are >= D = ceil((w + B - 1) / B) bins apart
// cyclically: m = nbd / D members per class at stride nbd / m, plus one single-bin class for each of the
// nbd % m left-over bins; a phase is the product of an x class and a y class.
static int get_small_sched(fv_plan* P, int64_t nf, int w, fv_plan::SmallSched** out) {
  auto key = std::make_pair(nf, w);
  auto it = P->small_scheds.find(key);
  if (it != P->small_scheds.end()) { *out = &it->second; return FV_OK; }
  const int B = t1s_bin_size(nf, w);
  const int nbd = (int)(nf / B), D = (w + B - 1 + B - 1) / B;
  const int m = nbd / D, stride = nbd / m, rem = nbd - m * stride;
  std::vector<std::vector<int>> cls;
  for (int c = 0; c < stride; ++c) {
    std::vector<int> v;
    for (int k = 0; k < m; ++k) v.push_back(c + k * stride);
    cls.push_back(v);
  }
  for (int r = 0; r < rem; ++r) cls.push_back({m * stride + r});
  std::vector<int32_t> ph_off{0};
  std::vector<uint16_t> ph_bins;
  for (auto& cy : cls)
    for (auto& cx : cls) {
      for (int by : cy) for (int bx : cx) ph_bins.push_back((uint16_t)(by * nbd + bx));
      ph_off.push_back((int32_t)ph_bins.size());
    }
  fv_plan::SmallSched sc;
  sc.nphase = (int)ph_off.size() - 1;
  FV_CUDA(cudaMalloc((void**)&sc.ph_off, sizeof(int32_t) * ph_off.size()));
  FV_CUDA(cudaMalloc((void**)&sc.ph_bins, sizeof(uint16_t) * ph_bins.size()));
  FV_CUDA(cudaMemcpyAsync(sc.ph_off, ph_off.data(), sizeof(int32_t) * ph_off.size(), cudaMemcpyHostToDevice, P->stream));
  FV_CUDA(cudaMemcpyAsync(sc.ph_bins, ph_bins.data(), sizeof(uint16_t) * ph_bins.size(), cudaMemcpyHostToDevice, P->stream));
  FV_CUDA(cudaStreamSynchronize(P->stream));
  auto res = P->small_scheds.emplace(key, sc);
  *out = &res.first->second;
  return FV_OK;
}

template <typename T, int WT>
static int launch_t1s(fv_plan* P, T1SmallArgs<T>& a, int nb, int ntr, int64_t nitems) {
  // as many warps (<= 16) as the per-warp record buffers leave room for beside the grid
  int nwarps = 16;
  while (nwarps > 4 && t1s_smem_bytes<T>(a.nf, a.nbd, a.nphase, nwarps, T1Small<WT>::REC) > 226 * 1024) nwarps -= 4;
  const size_t smem = t1s_smem_bytes<T>(a.nf, a.nbd, a.nphase, nwarps, T1Small<WT>::REC);
  if (smem > 226 * 1024) { set_error("small-grid type-1 path: grid does not fit shared memory"); return FV_ERR_UNSUPPORTED; }
  {
    const int64_t nt = nitems * 2;
    t1s_records_kernel<T, WT><<<(unsigned)((nt + 255) / 256), 256, 0, P->stream>>>(a, nitems);
    FV_LAUNCH_CHECK();
  }
  auto kern = t1s_spread_kernel<T, WT>;
  FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nb * ntr, 32 * nwarps, smem, P->stream>>>(a);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

// pass 1 of the small-grid path: fills Tbuf exactly like t1_spread_fftx_kernel does
template <typename T>
static int t1_small_pass1(fv_plan* P, const int32_t* n_dev, int64_t n_cap, int nb, int ntr, const void* W, int64_t nf,
                          int w, double beta, const int32_t* ix0, const int32_t* iy0, const T* zx, const T* zy,
                          const fv_plan::SmemFft* F, const fv_modeset::Tables* tab) {
  using C = cplx_t<T>;
  fv_plan::SmallSched* sc;
  int rc = get_small_sched(P, nf, w, &sc);
  if (rc) return rc;
  const int64_t nitems = (int64_t)nb * n_cap;
  const int B = t1s_bin_size(nf, w);
  const int nbd = (int)(nf / B), nbins = nbd * nbd;
  const int64_t nkeys = (int64_t)nb * nbins;
  int end_bit = 1;
  while ((1ll << end_bit) <= nkeys) ++end_bit;
  const int rec_len = 8 * ((w + 4) / 4);                 // T1Small<w>::REC
  rc = ensure(&P->rec, &P->rec_bytes, sizeof(T) * (size_t)nitems * rec_len);
  if (rc) return rc;
  const size_t off_bytes = sizeof(int32_t) * ((size_t)nkeys + 2);
  rc = ensure(&P->small, &P->small_bytes, 16 * (size_t)nitems + off_bytes + 64);
  if (rc) return rc;
  T1SmallArgs<T> a{};
  a.keys = (uint32_t*)P->small;
  a.vals = (int32_t*)(a.keys + nitems);
  uint32_t* skeys = a.keys + 2 * nitems;
  int32_t* svals = (int32_t*)(skeys + nitems);
  a.skeys = skeys; a.svals = svals;
  a.off = (int32_t*)(svals + nitems);
  a.rec = (T*)P->rec;
  a.n_dev = n_dev; a.n_cap = n_cap; a.nf = (int)nf; a.pitch = (int)nf + 1; a.w = w; a.nbd = nbd; a.B = B;
  a.beta = (T)beta; a.c = (T)(4.0 / ((double)w * w)); a.halfw = (T)(w / 2.0);
  a.ntr = ntr; a.W = (const C*)W; a.ix0 = ix0; a.iy0 = iy0; a.zx = zx; a.zy = zy;
  a.nphase = sc->nphase; a.ph_off = sc->ph_off; a.ph_bins = sc->ph_bins;
  a.tw = (const C*)F->tw; a.st = F->st; a.ncols = tab->ncols; a.col_pos = tab->col_pos; a.Tbuf = (C*)P->tbuf;
  a.nb = nb;
  {
    StageScope ts(P, FV_STAGE_ZERO);                     // reported with the prep pass: keys, sort, bin bounds
    dim3 grid(ceil_div(n_cap, 256), nb);
    t1s_key_kernel<T><<<grid, 256, 0, P->stream>>>(a);
    FV_LAUNCH_CHECK();
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, a.keys, skeys, a.vals, svals, (int)nitems, 0, end_bit, P->stream);
    rc = ensure(&P->scan_tmp, &P->scan_tmp_bytes, std::max<size_t>(tmp, 16));
    if (rc) return rc;
    FV_CUDA(cub::DeviceRadixSort::SortPairs(P->scan_tmp, tmp, a.keys, skeys, a.vals, svals, (int)nitems, 0, end_bit, P->stream));
    ++fv::g_launches;
    t1s_bounds_kernel<<<(unsigned)((nitems + 256) / 256), 256, 0, P->stream>>>(skeys, nitems, (uint32_t)nkeys, a.off);
    FV_LAUNCH_CHECK();
  }
  {
    StageScope ts(P, FV_STAGE_SPREAD);
    FV_DISPATCH_WS(w, (rc = launch_t1s<T, WT>(P, a, nb, ntr, nitems)));
    if (rc) return rc;
  }
  return FV_OK;
}


int t1_small_pass1_entry(fv_plan* P, int prec, const int32_t* n_dev, int64_t n_cap, int nb, int ntr, const void* W,
                         int64_t nf, int w, double beta, const int32_t* ix0, const int32_t* iy0, const void* zx,
                         const void* zy, const fv_plan::SmemFft* F, const fv_modeset::Tables* tab) {
  if (prec == 1) return t1_small_pass1<float>(P, n_dev, n_cap, nb, ntr, W, nf, w, beta, ix0, iy0, (const float*)zx, (const float*)zy, F, tab);
  return t1_small_pass1<double>(P, n_dev, n_cap, nb, ntr, W, nf, w, beta, ix0, iy0, (const double*)zx, (const double*)zy, F, tab);
}

}  // namespace fv
