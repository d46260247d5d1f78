// Stages a4 + a5 of the hot path fused: per-frequency beam evaluation (analytic or az/za table
// interpolation) and the apparent-coherency product, emitting the NUFFT strengths directly.
// Replaces: CPUBeamEvaluator.evaluate_beam (reference cpu/beams.py:12-89; pyuvdata compute_response),
//   _evaluate_beam_list (cpu_simulate.py:38-87), _compute_apparent_coherency (cpu_simulate.py:90-202)
//   and the four numba kernels (cpu/beams.py:129-246).
// One thread per (source, frequency); beam arithmetic in fp64 (the reference evaluates beams in
// fp64 and casts, cpu_simulate.py:84-86), products in the working precision.  Catalogue fluxes are
// stored frequency-major so that the gather through src_idx (ascending) is coalesced.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace fv {

constexpr double kC = 299792458.0;

struct BeamVal { double re[4], im[4]; };   // efield: [vec*2+feed]; power: re[0]

template <typename T> struct tab_elem;      // how table entries are stored
template <> struct tab_elem<float> { using real = float; using cplx = float2; };
template <> struct tab_elem<double> { using real = double; using cplx = double2; };

// grid index of a direction on an az/za table (fractional; azimuth wrapped and shifted into the host's
// wrap-extended axis)
__device__ __forceinline__ void table_index(const fv_beam& b, double az, double za, double& zi, double& ai) {
  zi = (za - b.za0) / b.dza;
  ai = (az - b.az0) / b.daz;
  if (b.az_wrap_period > 0) {
    ai = fmod(ai, (double)b.az_wrap_period);
    if (ai < 0) ai += (double)b.az_wrap_period;
    ai += (double)b.az_pad;
  }
}

// the (up to) four taps and weights of orders 0 / 1, mode 'nearest' outside the grid
__device__ __forceinline__ void table_taps(const fv_beam& b, double zi, double ai, int& z0, int& z1, int& a0, int& a1,
                                           double& wz, double& wa) {
  if (b.order == 0) {
    z0 = z1 = min(max((int)floor(zi + 0.5), 0), b.nza - 1);
    a0 = a1 = min(max((int)floor(ai + 0.5), 0), b.naz - 1);
    wz = wa = 0.0;
  } else {
    const double zf = floor(zi), af = floor(ai);
    wz = zi - zf; wa = ai - af;
    z0 = min(max((int)zf, 0), b.nza - 1); z1 = min(max((int)zf + 1, 0), b.nza - 1);
    a0 = min(max((int)af, 0), b.naz - 1); a1 = min(max((int)af + 1, 0), b.naz - 1);
  }
}

template <typename T>
__device__ inline void eval_beam(const fv_beam& b, double az, double za, double freq, int fb,
                                 BeamVal& v) {
  if (b.kind != 3) {
    double e;
    if (b.kind == 0) {
      const double s = asin(2.2150894 * (kC / freq) / (M_PI * b.diameter)) * 2.0 / 2.355;
      e = exp(-(za * za) / (2.0 * s * s));
    } else if (b.kind == 1) {
      const double x = M_PI * b.diameter * sin(za) * freq / kC;
      if (sizeof(T) == 4) {
        // single precision: the argument in fp64 (it reaches ~30 rad), the Bessel function itself in fp32 -- its
        // ~1e-7 absolute error is below the rounding of the fp32 strengths it is folded into
        const float xf = (float)x;
        e = (x == 0.0) ? 1.0 : (double)(2.0f * j1f(xf) / xf);
      } else {
        e = (x == 0.0) ? 1.0 : 2.0 * j1(x) / x;
      }
    } else {
      e = 1.0;
    }
    if (b.is_power) {
      v.re[0] = e * e; v.im[0] = 0.0;
    } else {
      const double q = e / 1.4142135623730951;
#pragma unroll
      for (int c = 0; c < 4; ++c) { v.re[c] = q; v.im[c] = 0.0; }
    }
    return;
  }
  // ---- az/za table, mode 'nearest' outside the grid (scipy.ndimage.map_coordinates semantics)
  double zi, ai;
  table_index(b, az, za, zi, ai);
  const int ncomp = b.is_power ? 1 : 4;
  const int64_t plane = (int64_t)b.nza * b.naz;
  const int fi = b.freq_offset + fb;
  if (b.order == 3) {
    // cubic B-spline on prefiltered coefficients (scipy.ndimage.map_coordinates order=3, mode
    // 'nearest': coordinates clamped to the original grid, coefficients of the 12-cell edge-padded
    // grid; every tap of the 4 x 4 stencil then lies inside the padded table)
    const int pad = b.spline_pad;
    const int nza0 = b.nza - 2 * pad, naz0 = b.naz - 2 * pad;
    zi = fmin(fmax(zi, 0.0), (double)(nza0 - 1)) + pad;
    ai = fmin(fmax(ai, 0.0), (double)(naz0 - 1)) + pad;
    const double zf = floor(zi), af = floor(ai);
    const double tz = zi - zf, ta = ai - af;
    double wz[4], wa[4];
    auto bsp = [](double t, double* wgt) {
      const double t2 = t * t, t3 = t2 * t, u = 1.0 - t;
      wgt[0] = u * u * u / 6.0;
      wgt[1] = (3.0 * t3 - 6.0 * t2 + 4.0) / 6.0;
      wgt[2] = (-3.0 * t3 + 3.0 * t2 + 3.0 * t + 1.0) / 6.0;
      wgt[3] = t3 / 6.0;
    };
    bsp(tz, wz); bsp(ta, wa);
    const int zb = (int)zf - 1, ab = (int)af - 1;
    for (int c = 0; c < ncomp; ++c) {
      double re = 0.0, im = 0.0;
      if (b.is_power) {
        const typename tab_elem<T>::real* t = (const typename tab_elem<T>::real*)b.table + ((int64_t)fi) * plane;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) re += wz[i] * wa[j] * (double)t[(int64_t)(zb + i) * b.naz + ab + j];
      } else {
        const typename tab_elem<T>::cplx* t = (const typename tab_elem<T>::cplx*)b.table + ((int64_t)fi * 4 + c) * plane;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const auto pv = t[(int64_t)(zb + i) * b.naz + ab + j];
            const double ww = wz[i] * wa[j];
            re += ww * (double)pv.x; im += ww * (double)pv.y;
          }
      }
      v.re[c] = re; v.im[c] = im;
    }
    return;
  }
  int z0, z1, a0, a1;
  double wz, wa;
  table_taps(b, zi, ai, z0, z1, a0, a1, wz, wa);
  const double w00 = (1 - wz) * (1 - wa), w01 = (1 - wz) * wa, w10 = wz * (1 - wa), w11 = wz * wa;
  // nfreq_table is implied by the caller's indexing: plane stride per (comp, freq)
  for (int c = 0; c < ncomp; ++c) {
    if (b.is_power) {
      const typename tab_elem<T>::real* t =
          (const typename tab_elem<T>::real*)b.table + ((int64_t)fi) * plane;
      v.re[0] = w00 * t[(int64_t)z0 * b.naz + a0] + w01 * t[(int64_t)z0 * b.naz + a1] +
                w10 * t[(int64_t)z1 * b.naz + a0] + w11 * t[(int64_t)z1 * b.naz + a1];
      v.im[0] = 0.0;
    } else {
      // layout (nfreq_table, 4, nza, naz): the four Jones entries of one frequency are adjacent
      const typename tab_elem<T>::cplx* t =
          (const typename tab_elem<T>::cplx*)b.table + ((int64_t)fi * 4 + c) * plane;
      const auto p00 = t[(int64_t)z0 * b.naz + a0], p01 = t[(int64_t)z0 * b.naz + a1];
      const auto p10 = t[(int64_t)z1 * b.naz + a0], p11 = t[(int64_t)z1 * b.naz + a1];
      v.re[c] = w00 * p00.x + w01 * p01.x + w10 * p10.x + w11 * p11.x;
      v.im[c] = w00 * p00.y + w01 * p01.y + w10 * p10.y + w11 * p11.y;
    }
  }
}

// The four 2x2 products of the reference (cpu/beams.py:129-246), A[b, feed] at index b*2+feed:
//   mode 1: out[a,p] = sum_b conj(Ai[b,a]) Aj[b,p] F                 (F = Cm[0])
//   mode 2: out[a,p] = sum_{b,k} conj(Ai'[b,a]) C[b,k] Aj'[k,p],  A'[b] = A[1-b]  (the axis-0 flip
//           of the polarised-sky branch, cpu_simulate.py:146-147,153)
//   mode 4: as mode 2 without the flip (the bare numba kernel, cpu/beams.py:147-180,215-246)
template <typename T>
__device__ __forceinline__ void coherency_product(int mode, const cplx_t<T>* Ai, const cplx_t<T>* Aj,
                                                  const cplx_t<T>* Cm, cplx_t<T>* out) {
  using C = cplx_t<T>;
  if (mode == 1 || mode == 3) {
#pragma unroll
    for (int aa = 0; aa < 2; ++aa)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const C t = cadd(cmulc(Ai[0 * 2 + aa], Aj[0 * 2 + pp]), cmulc(Ai[1 * 2 + aa], Aj[1 * 2 + pp]));
        out[aa * 2 + pp] = cmul(t, Cm[0]);
      }
    return;
  }
  const int fl = mode == 2 ? 1 : 0;
  C tmp[4];   // tmp[a,k] = sum_b conj(Fi[b,a]) C[b,k]
#pragma unroll
  for (int aa = 0; aa < 2; ++aa)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      tmp[aa * 2 + kk] = cadd(cmulc(Ai[(0 ^ fl) * 2 + aa], Cm[0 * 2 + kk]), cmulc(Ai[(1 ^ fl) * 2 + aa], Cm[1 * 2 + kk]));
#pragma unroll
  for (int aa = 0; aa < 2; ++aa)
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
      out[aa * 2 + pp] = cadd(cmul(tmp[aa * 2 + 0], Aj[(0 ^ fl) * 2 + pp]), cmul(tmp[aa * 2 + 1], Aj[(1 ^ fl) * 2 + pp]));
}

template <typename T>
__global__ void __launch_bounds__(256, 3)
weights_kernel(int mode, fv_beam bi, fv_beam bj, int same_beam, const T* __restrict__ az,
               const T* __restrict__ za, const int32_t* __restrict__ src_idx,
               const int32_t* __restrict__ n_dev, int64_t n_cap, const double* __restrict__ freqs,
               int64_t f0, const cplx_t<T>* __restrict__ flux, int64_t nsrc_total,
               cplx_t<T>* __restrict__ out, cplx_t<T>* __restrict__ out_beam) {
  using C = cplx_t<T>;
  const int n = *n_dev;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int fb = blockIdx.y;
  const double freq = freqs[f0 + fb];
  const double a = (double)az[s], z = (double)za[s];
  BeamVal vi, vj;
  eval_beam<T>(bi, a, z, freq, fb, vi);
  if (same_beam) vj = vi; else eval_beam<T>(bj, a, z, freq, fb, vj);
  const int64_t src = src_idx[s];
  const int P = mode == 0 ? 1 : 4;
  C* o = out + ((int64_t)fb * P) * n_cap + s;
  if (out_beam) {
    const int nc = bi.is_power ? 1 : 4;
    for (int c = 0; c < nc; ++c)
      out_beam[((int64_t)fb * nc + c) * n_cap + s] = make_c<T>((T)vi.re[c], (T)vi.im[c]);
  }
  if (mode == 0) {
    // sqrt(B_i B_j) F with the complex principal root (beams were cast to complex first)
    const T bi_ = (T)vi.re[0], bj_ = (T)vj.re[0];
    const T p = bi_ * bj_;
    const C r = p >= T(0) ? make_c<T>(sqrt(p), T(0)) : make_c<T>(T(0), sqrt(-p));
    o[0] = cmul(r, flux[(int64_t)(f0 + fb) * nsrc_total + src]);
    return;
  }
  C Ai[4], Aj[4], Cm[4], res[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    Ai[c] = make_c<T>((T)vi.re[c], (T)vi.im[c]);
    Aj[c] = make_c<T>((T)vj.re[c], (T)vj.im[c]);
  }
  if (mode == 1) {
    Cm[0] = flux[(int64_t)(f0 + fb) * nsrc_total + src];
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) Cm[c] = flux[((int64_t)(f0 + fb) * 4 + c) * nsrc_total + src];
  }
  coherency_product<T>(mode, Ai, Aj, Cm, res);
#pragma unroll
  for (int c = 0; c < 4; ++c) o[(int64_t)c * n_cap] = res[c];
}

// stand-alone form of the products (API parity with the CPU evaluator's four methods)
template <typename T>
__global__ void __launch_bounds__(256)
coherency_kernel(int mode, const cplx_t<T>* __restrict__ bi, const cplx_t<T>* __restrict__ bj,
                 const cplx_t<T>* __restrict__ fc, int64_t n, cplx_t<T>* __restrict__ out) {
  using C = cplx_t<T>;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  C Ai[4], Aj[4], Cm[4], res[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { Ai[c] = bi[c * n + s]; Aj[c] = bj[c * n + s]; }
  if (mode == 1 || mode == 3) Cm[0] = fc[s];
  else {
#pragma unroll
    for (int c = 0; c < 4; ++c) Cm[c] = fc[c * n + s];
  }
  coherency_product<T>(mode, Ai, Aj, Cm, res);
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c * n + s] = res[c];
}

// Basis path (cpu_simulate.py:416-468): every basis beam is evaluated ONCE per (source, frequency) and the
// K (K + 1) / 2 pair products (k <= l) are formed from the K Jones matrices in registers; the strengths of
// pair q go to transforms 4 q .. 4 q + 3 of the batched NUFFT.
constexpr int kMaxBasis = 8;
struct BeamSet { fv_beam b[kMaxBasis]; int K; };

template <typename T>
__global__ void __launch_bounds__(256, 2)
weights_basis_kernel(int mode, BeamSet bs, const T* __restrict__ az, const T* __restrict__ za,
                     const int32_t* __restrict__ src_idx, const int32_t* __restrict__ n_dev, int64_t n_cap,
                     const double* __restrict__ freqs, int64_t f0, const cplx_t<T>* __restrict__ flux,
                     int64_t nsrc_total, cplx_t<T>* __restrict__ out) {
  using C = cplx_t<T>;
  const int n = *n_dev;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int fb = blockIdx.y;
  const double freq = freqs[f0 + fb];
  const double a = (double)az[s], z = (double)za[s];
  C A[kMaxBasis][4];
#pragma unroll
  for (int k = 0; k < kMaxBasis; ++k) {
    if (k < bs.K) {
      BeamVal v;
      eval_beam<T>(bs.b[k], a, z, freq, fb, v);
#pragma unroll
      for (int c = 0; c < 4; ++c) A[k][c] = make_c<T>((T)v.re[c], (T)v.im[c]);
    }
  }
  const int64_t src = src_idx[s];
  C Cm[4];
  if (mode == 1) {
    Cm[0] = flux[(int64_t)(f0 + fb) * nsrc_total + src];
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) Cm[c] = flux[((int64_t)(f0 + fb) * 4 + c) * nsrc_total + src];
  }
  const int npairs = bs.K * (bs.K + 1) / 2;
  C* o = out + ((int64_t)fb * npairs * 4) * n_cap + s;
  int q = 0;
#pragma unroll
  for (int k = 0; k < kMaxBasis; ++k)
#pragma unroll
    for (int l = k; l < kMaxBasis; ++l) {
      if (l < bs.K) {
        C res[4];
        coherency_product<T>(mode, A[k], A[l], Cm, res);
#pragma unroll
        for (int c = 0; c < 4; ++c) o[(int64_t)(q * 4 + c) * n_cap] = res[c];
        ++q;
      }
    }
}

}  // namespace fv

#include "weights_tiled.cuh"

// sorted-by-tile state of one time step's live set (fv_tiles_*)
struct fv_tiles {
  cudaStream_t stream = nullptr;
  void* sortbuf = nullptr; size_t sort_bytes = 0;      // keys, vals, sorted keys, sorted vals (n_cap each)
  void* scratch = nullptr; size_t scratch_bytes = 0;   // permuted copy of the live set
  void* cub_tmp = nullptr; size_t cub_bytes = 0;
  int32_t* tile_off = nullptr; int off_cap = 0;
  fv_beam geom{};                                      // grid the tiles were built for
  int ntiles = 0; int64_t n_cap = 0; bool valid = false;
};

namespace fv {

static int grow(void** p, size_t* have, size_t need) {
  if (*have >= need) return FV_OK;
  if (*p) FV_CUDA(cudaFree(*p));
  *p = nullptr; *have = 0;
  FV_CUDA(cudaMalloc(p, need));
  *have = need;
  return FV_OK;
}

static bool same_table_grid(const fv_beam& a, const fv_beam& b) {
  return a.kind == 3 && b.kind == 3 && a.nza == b.nza && a.naz == b.naz && a.az_wrap_period == b.az_wrap_period &&
         a.az_pad == b.az_pad && a.az0 == b.az0 && a.daz == b.daz && a.za0 == b.za0 && a.dza == b.dza &&
         a.order == b.order && a.is_power == b.is_power;
}

template <typename T, typename E, int NC>
static int launch_weights_tiled(const WtArgs<T>& a_in, int ntiles, cudaStream_t st) {
  WtArgs<T> a = a_in;
  const size_t per_stage = (size_t)a.K * NC * WT_PTS * WT_PTS * sizeof(E);
  a.nstage = (int)std::max<size_t>(2, std::min<size_t>(WT_MAXSTAGE, (size_t)(100 * 1024) / per_stage));
  const size_t smem = a.nstage * per_stage;
  dim3 grid(ntiles, ceil_div(a.nf, a.freqs_per_cta));
#define FV_WT_LAUNCH(KM)                                                                               \
  {                                                                                                    \
    auto kern = weights_tiled_kernel<T, E, NC, KM>;                                                    \
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    kern<<<grid, WT_THREADS, smem, st>>>(a);                                                           \
  }
  if (a.K <= 2) FV_WT_LAUNCH(2) else if (a.K <= 5) FV_WT_LAUNCH(5) else FV_WT_LAUNCH(6)
#undef FV_WT_LAUNCH
  FV_LAUNCH_CHECK();
  return FV_OK;
}

template <typename T>
static int weights_tiled_impl(fv_tiles* h, int mode, const fv_beam* beams, int K, int basis, const void* az,
                              const void* za, const int32_t* src_idx, int64_t n_cap, const double* freqs, int nf,
                              int64_t f0, const void* flux, int64_t nsrc_total, void* out) {
  WtArgs<T> a{};
  a.mode = mode; a.K = K; a.basis = basis;
  for (int k = 0; k < 8; ++k) a.b[k] = beams[k < K ? k : 0];
  a.tile_off = h->tile_off; a.az = (const T*)az; a.za = (const T*)za; a.src_idx = src_idx; a.n_cap = n_cap;
  a.freqs = freqs; a.f0 = f0; a.nf = nf;
  // enough CTAs to fill the GPU a few times over, long enough runs of frequencies to pipeline the patches
  a.freqs_per_cta = std::max(4, (int)ceil_div((int64_t)nf * h->ntiles, 8 * kNumSMs));
  a.flux = (const cplx_t<T>*)flux; a.nsrc_total = nsrc_total; a.out = (cplx_t<T>*)out;
  if (beams[0].is_power) {
    a.bulk = 0;
    return launch_weights_tiled<T, T, 1>(a, h->ntiles, h->stream);
  }
  a.bulk = sizeof(cplx_t<T>) == 16 ? 1 : 0;
  return launch_weights_tiled<T, cplx_t<T>, 4>(a, h->ntiles, h->stream);
}

static bool same_beam_desc(const fv_beam& a, const fv_beam& b) {
  return a.kind == b.kind && a.is_power == b.is_power && a.diameter == b.diameter &&
         a.table == b.table && a.order == b.order && a.freq_offset == b.freq_offset;
}

}  // namespace fv

extern "C" int fv_weights(int prec, int mode, const fv_beam* beam_i_host, const fv_beam* beam_j_host,
                          const void* az, const void* za, const int32_t* src_idx,
                          const int32_t* n_dev, int64_t n_cap, const double* freqs, int nf,
                          int64_t freq_index0, const void* flux, int64_t nsrc_total, void* out,
                          void* out_beam_i, void* stream) {
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  FV_REQUIRE(beam_i_host && beam_j_host && az && za && src_idx && n_dev && freqs && flux && out,
             "null pointer");
  FV_REQUIRE((mode == 0) == (beam_i_host->is_power != 0) && (mode == 0) == (beam_j_host->is_power != 0),
             "mode 0 needs power beams, modes 1/2 need E-field beams");
  for (const fv_beam* b : {beam_i_host, beam_j_host}) {
    FV_REQUIRE(b->kind >= 0 && b->kind <= 3, "unknown beam kind");
    if (b->kind == 3) {
      FV_REQUIRE(b->table && b->nza > 0 && b->naz > 0, "table beam without table");
      if (!(b->order == 0 || b->order == 1 || b->order == 3)) {
        fv::set_error("beam interpolation order must be 0, 1 or 3");
        return FV_ERR_UNSUPPORTED;
      }
      FV_REQUIRE(b->order != 3 || (b->spline_pad >= 2 && b->nza > 2 * b->spline_pad && b->naz > 2 * b->spline_pad),
                 "order-3 table needs spline_pad >= 2 and a padded coefficient grid");
    }
  }
  if (nf == 0 || n_cap == 0) return FV_OK;
  FV_REQUIRE(nf <= 65535, "at most 65535 frequencies per call");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(fv::ceil_div(n_cap, 256), nf);
  const int same = fv::same_beam_desc(*beam_i_host, *beam_j_host);
  if (prec == 1)
    fv::weights_kernel<float><<<grid, 256, 0, st>>>(
        mode, *beam_i_host, *beam_j_host, same, (const float*)az, (const float*)za, src_idx, n_dev,
        n_cap, freqs, freq_index0, (const float2*)flux, nsrc_total, (float2*)out, (float2*)out_beam_i);
  else
    fv::weights_kernel<double><<<grid, 256, 0, st>>>(
        mode, *beam_i_host, *beam_j_host, same, (const double*)az, (const double*)za, src_idx, n_dev,
        n_cap, freqs, freq_index0, (const double2*)flux, nsrc_total, (double2*)out,
        (double2*)out_beam_i);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_weights_basis(int prec, int mode, const fv_beam* beams_host, int K, const void* az, const void* za,
                                const int32_t* src_idx, const int32_t* n_dev, int64_t n_cap, const double* freqs,
                                int nf, int64_t freq_index0, const void* flux, int64_t nsrc_total, void* out,
                                void* stream) {
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(mode == 1 || mode == 2, "the basis path is polarised: mode must be 1 or 2");
  FV_REQUIRE(beams_host && az && za && src_idx && n_dev && freqs && flux && out, "null pointer");
  FV_REQUIRE(K >= 1 && K <= fv::kMaxBasis, "1 <= K <= 8 basis beams");
  fv::BeamSet bs;
  bs.K = K;
  for (int k = 0; k < fv::kMaxBasis; ++k) bs.b[k] = beams_host[k < K ? k : 0];
  for (int k = 0; k < K; ++k) {
    const fv_beam* b = &bs.b[k];
    FV_REQUIRE(!b->is_power, "basis beams must be E-field beams");
    FV_REQUIRE(b->kind >= 0 && b->kind <= 3, "unknown beam kind");
    if (b->kind == 3) {
      FV_REQUIRE(b->table && b->nza > 0 && b->naz > 0, "table beam without table");
      if (!(b->order == 0 || b->order == 1 || b->order == 3)) {
        fv::set_error("beam interpolation order must be 0, 1 or 3");
        return FV_ERR_UNSUPPORTED;
      }
    }
  }
  if (nf == 0 || n_cap == 0) return FV_OK;
  FV_REQUIRE(nf <= 65535, "at most 65535 frequencies per call");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(fv::ceil_div(n_cap, 256), nf);
  if (prec == 1)
    fv::weights_basis_kernel<float><<<grid, 256, 0, st>>>(mode, bs, (const float*)az, (const float*)za, src_idx, n_dev, n_cap,
                                                         freqs, freq_index0, (const float2*)flux, nsrc_total, (float2*)out);
  else
    fv::weights_basis_kernel<double><<<grid, 256, 0, st>>>(mode, bs, (const double*)az, (const double*)za, src_idx, n_dev, n_cap,
                                                          freqs, freq_index0, (const double2*)flux, nsrc_total, (double2*)out);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

/* stand-alone apparent-coherency products on caller-supplied beam values (host API parity with
 * CPUBeamEvaluator.get_apparent_flux_polarized{,_beam,_beam_pair,_pair}, cpu/beams.py:129-246).
 * mode 1: A_i^H diag(F) A_j;  mode 4: A_i^H C A_j (no flip);  mode 2: with the axis-0 flip.
 * beams / coherency / out: (4, n) cplx rows [b*2+feed]; flux (n) cplx. */
extern "C" int fv_coherency(int prec, int mode, const void* beam_i, const void* beam_j,
                            const void* flux_or_coh, int64_t n, void* out, void* stream) {
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(mode == 1 || mode == 2 || mode == 4, "mode must be 1, 2 or 4");
  FV_REQUIRE(beam_i && beam_j && flux_or_coh && out, "null pointer");
  if (n == 0) return FV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = fv::ceil_div(n, 256);
  if (prec == 1)
    fv::coherency_kernel<float><<<blocks, 256, 0, st>>>(mode, (const float2*)beam_i, (const float2*)beam_j, (const float2*)flux_or_coh, n, (float2*)out);
  else
    fv::coherency_kernel<double><<<blocks, 256, 0, st>>>(mode, (const double2*)beam_i, (const double2*)beam_j, (const double2*)flux_or_coh, n, (double2*)out);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

/* ---- beam tables staged in shared memory (weights_tiled.cuh) ----------------------------------- */
extern "C" int fv_tiles_create(fv_tiles** h, void* stream) {
  FV_REQUIRE(h, "null pointer");
  *h = new fv_tiles();
  (*h)->stream = (cudaStream_t)stream;
  return FV_OK;
}

extern "C" int fv_tiles_destroy(fv_tiles* h) {
  if (!h) return FV_OK;
  if (h->sortbuf) cudaFree(h->sortbuf);
  if (h->scratch) cudaFree(h->scratch);
  if (h->cub_tmp) cudaFree(h->cub_tmp);
  if (h->tile_off) cudaFree(h->tile_off);
  delete h;
  return FV_OK;
}

extern "C" int fv_tiles_supported(const fv_beam* beams_host, int K) {
  if (!beams_host || K < 1 || K > 6) return 0;
  for (int k = 0; k < K; ++k) {
    const fv_beam& b = beams_host[k];
    if (b.kind != 3 || !(b.order == 0 || b.order == 1) || !fv::same_table_grid(beams_host[0], b)) return 0;
  }
  return 1;
}

template <typename T>
static int tiles_sort_impl(fv_tiles* h, const fv_beam& b, void* xyz, void* az, void* za, int32_t* src_idx,
                           const int32_t* n_dev, int64_t n_cap) {
  using namespace fv;
  const TileGrid g = tile_grid(b);
  const int ntiles = g.ntz * g.nta;
  int rc = grow(&h->sortbuf, &h->sort_bytes, 16 * (size_t)n_cap);
  if (rc) return rc;
  rc = grow(&h->scratch, &h->scratch_bytes, (5 * sizeof(T) + 4) * (size_t)n_cap);
  if (rc) return rc;
  if (h->off_cap < ntiles + 1) {
    if (h->tile_off) FV_CUDA(cudaFree(h->tile_off));
    FV_CUDA(cudaMalloc((void**)&h->tile_off, sizeof(int32_t) * (ntiles + 1)));
    h->off_cap = ntiles + 1;
  }
  uint32_t* keys = (uint32_t*)h->sortbuf;
  int32_t* vals = (int32_t*)(keys + n_cap);
  uint32_t* skeys = (uint32_t*)(vals + n_cap);
  int32_t* svals = (int32_t*)(skeys + n_cap);
  const int blocks = ceil_div(n_cap, 256);
  tile_key_kernel<T><<<blocks, 256, 0, h->stream>>>(b, (const T*)az, (const T*)za, n_dev, n_cap, keys, vals);
  FV_LAUNCH_CHECK();
  int end_bit = 1;
  while ((1ll << end_bit) <= ntiles) ++end_bit;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, skeys, vals, svals, (int)n_cap, 0, end_bit, h->stream);
  rc = grow(&h->cub_tmp, &h->cub_bytes, std::max<size_t>(tmp, 16));
  if (rc) return rc;
  FV_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, keys, skeys, vals, svals, (int)n_cap, 0, end_bit, h->stream));
  ++g_launches;
  tile_bounds_kernel<<<ceil_div(n_cap + 1, 256), 256, 0, h->stream>>>(skeys, n_cap, (uint32_t)ntiles, h->tile_off);
  FV_LAUNCH_CHECK();
  T* o_xyz = (T*)h->scratch;
  T* o_az = o_xyz + 3 * n_cap;
  T* o_za = o_az + n_cap;
  int32_t* o_src = (int32_t*)(o_za + n_cap);
  tile_permute_kernel<T><<<blocks, 256, 0, h->stream>>>(svals, n_dev, n_cap, (const T*)xyz, (const T*)az, (const T*)za,
                                                     src_idx, o_xyz, o_az, o_za, o_src);
  FV_LAUNCH_CHECK();
  // dead slots keep whatever they held (nothing reads them): copy back the whole arrays
  FV_CUDA(cudaMemcpyAsync(xyz, o_xyz, sizeof(T) * 3 * n_cap, cudaMemcpyDeviceToDevice, h->stream));
  FV_CUDA(cudaMemcpyAsync(az, o_az, sizeof(T) * n_cap, cudaMemcpyDeviceToDevice, h->stream));
  FV_CUDA(cudaMemcpyAsync(za, o_za, sizeof(T) * n_cap, cudaMemcpyDeviceToDevice, h->stream));
  FV_CUDA(cudaMemcpyAsync(src_idx, o_src, sizeof(int32_t) * n_cap, cudaMemcpyDeviceToDevice, h->stream));
  h->geom = b; h->ntiles = ntiles; h->n_cap = n_cap; h->valid = true;
  return FV_OK;
}

extern "C" int fv_tiles_sort(fv_tiles* h, int prec, const fv_beam* beam_host, void* xyz, void* az, void* za,
                             int32_t* src_idx, const int32_t* n_dev, int64_t n_cap) {
  FV_REQUIRE(h && beam_host && xyz && az && za && src_idx && n_dev, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(fv_tiles_supported(beam_host, 1), "tiles need an az/za table beam of order 0 or 1");
  FV_REQUIRE(n_cap > 0 && n_cap < (1ll << 31), "bad n_cap");
  if (prec == 1) return tiles_sort_impl<float>(h, *beam_host, xyz, az, za, src_idx, n_dev, n_cap);
  return tiles_sort_impl<double>(h, *beam_host, xyz, az, za, src_idx, n_dev, n_cap);
}

extern "C" int fv_weights_tiled(fv_tiles* h, int prec, int mode, const fv_beam* beams_host, int K, int basis,
                                const void* az, const void* za, const int32_t* src_idx, int64_t n_cap,
                                const double* freqs, int nf, int64_t freq_index0, const void* flux, int64_t nsrc_total,
                                void* out) {
  FV_REQUIRE(h && beams_host && az && za && src_idx && freqs && flux && out, "null pointer");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  FV_REQUIRE(h->valid && h->n_cap == n_cap, "fv_tiles_sort has not been run for this live set");
  FV_REQUIRE(fv_tiles_supported(beams_host, K) && fv::same_table_grid(h->geom, beams_host[0]),
             "beams do not share the grid the tiles were built for");
  FV_REQUIRE(!basis || (mode != 0 && K >= 1), "the basis form is polarised");
  FV_REQUIRE(basis || K <= 2, "the pair form takes one or two beams");
  FV_REQUIRE((mode == 0) == (beams_host[0].is_power != 0), "mode 0 needs power beams, modes 1/2 need E-field beams");
  if (nf == 0) return FV_OK;
  if (prec == 1)
    return fv::weights_tiled_impl<float>(h, mode, beams_host, K, basis, az, za, src_idx, n_cap, freqs, nf, freq_index0, flux, nsrc_total, out);
  return fv::weights_tiled_impl<double>(h, mode, beams_host, K, basis, az, za, src_idx, n_cap, freqs, nf, freq_index0, flux, nsrc_total, out);
}
