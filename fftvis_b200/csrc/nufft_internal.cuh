// Shared internals of the NUFFT translation units (nufft.cu, type1_fused.cu, type3.cu): parameter
// rules, the kernels several of them launch, the plan object and its helpers.  Everything here is
// static or a template, so each translation unit instantiates only what it uses.
#pragma once
#include <math.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace fv {

// ------------------------------------------------------------------------------------------------
// host-side parameter rules
// ------------------------------------------------------------------------------------------------
static void kernel_params(double eps, double upsampfac, int prec, int* w, double* beta) {
  const double mach = prec == 2 ? 1.1e-16 : 6e-8;
  eps = std::max(eps, mach);
  int ns;
  if (upsampfac == 2.0) ns = (int)ceil(-log10(eps / 10.0));
  else ns = (int)ceil(-log(eps) / (M_PI * sqrt(1.0 - 1.0 / upsampfac)));
  ns = std::min(std::max(ns, 2), kMaxW);
  double bon = 2.30;
  if (upsampfac == 2.0) {
    if (ns == 2) bon = 2.20;
    if (ns == 3) bon = 2.26;
    if (ns == 4) bon = 2.38;
  } else {
    bon = 0.97 * M_PI * (1.0 - 1.0 / (2.0 * upsampfac));
  }
  *w = ns;
  *beta = bon * ns;
}

static int64_t next235even(int64_t n) {
  if (n <= 2) return 2;
  if (n % 2) ++n;
  for (;; n += 2) {
    int64_t m = n;
    while (m % 2 == 0) m /= 2;
    while (m % 3 == 0) m /= 3;
    while (m % 5 == 0) m /= 5;
    if (m == 1) return n;
  }
}

// Gauss-Legendre nodes on (-1,1) by Newton iteration (host, fp64)
static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double z1 = z; z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-15) break;
    }
    x[i] = -z; x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

struct Quad { int q; double z[32], f[32]; };   // nodes on (0, w/2) and weight*phi(node)

static Quad make_quad(int w, double beta) {
  Quad Q;
  Q.q = (int)(2 + 3.0 * (w / 2.0));
  std::vector<double> x, wt;
  gauss_legendre(2 * Q.q, x, wt);
  const double J2 = w / 2.0;
  for (int n = 0; n < Q.q; ++n) {
    const double z = x[Q.q + n] * J2;                       // positive half
    const double a = 1.0 - (2.0 * z / w) * (2.0 * z / w);
    Q.z[n] = z;
    Q.f[n] = wt[Q.q + n] * J2 * (a > 0 ? exp(beta * (sqrt(a) - 1.0)) : 0.0);
  }
  return Q;
}

// phihat(k), k = 0..nf/2, including the (-1)^k of the half-grid fold shift
static std::vector<double> kernel_ft_series(int64_t nf, const Quad& Q) {
  std::vector<double> ph(nf / 2 + 1);
  for (int64_t k = 0; k <= nf / 2; ++k) {
    double s = 0;
    for (int n = 0; n < Q.q; ++n) s += Q.f[n] * 2.0 * cos(2.0 * M_PI * (double)k * Q.z[n] / (double)nf);
    ph[k] = (k % 2) ? -s : s;
  }
  return ph;
}

// ------------------------------------------------------------------------------------------------
// device argument blocks
// ------------------------------------------------------------------------------------------------
struct BatchParams {      // one per frequency of a batch (device array)
  double smul;            // scalar on the source coordinates, applied in working precision (type 1: freq)
  double tmul;            // scalar on the target coordinates, applied in working precision (type 3: freq)
  double C[3];            // centre of the NU points            (type 3)
  double invgam[3];       // 1/gamma_d                           (type 3; 1 for type 1)
  double D[3];            // centre of the targets               (type 3)
  double hgam[3];         // h_d * gamma_d                       (type 3)
};

struct EpiDev {
  void* out; int64_t sb, sp; int32_t pmap[4]; const int32_t* kmap; const uint8_t* conj_flag; int acc;
};

template <typename C>
__device__ __forceinline__ void epilogue_store(const EpiDev& e, int b, int p, int64_t k, C v) {
  if (e.conj_flag && e.conj_flag[k]) v.y = -v.y;
  // transforms beyond the first four (batched basis pairs) keep the feed map inside their own group of four
  const int64_t idx = (int64_t)b * e.sb + (int64_t)(e.pmap[p & 3] + (p & ~3)) * e.sp + (e.kmap ? (int64_t)e.kmap[k] : k);
  C* o = (C*)e.out + idx;
  if (e.acc) { C t = *o; t.x += v.x; t.y += v.y; *o = t; } else { *o = v; }
}

template <typename T>
struct SpreadArgs {
  const T* x[3];
  const int32_t* n_dev;
  int64_t n_cap;
  int nf[3];
  int w;
  T beta, c, halfw;
  int ntr;
  int prephase;                 // type 3 with a non-zero target centre
  const cplx_t<T>* W;           // (nb, ntr, n_cap)
  cplx_t<T>* grid;              // (nb, ntr, nf3, nf2, nf1)
  const BatchParams* bp;
};

// ------------------------------------------------------------------------------------------------
// spread: one thread per (source, frequency); vector RED.ADD into the (L2-resident) fine grids.
// W == 0 selects the run-time width fallback.
// ------------------------------------------------------------------------------------------------
template <typename T, int DIM, int WT>
__global__ void __launch_bounds__(128)
spread_kernel(SpreadArgs<T> a) {
  using C = cplx_t<T>;
  const int n = *a.n_dev;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int b = blockIdx.y;
  const BatchParams bp = a.bp[b];
  const int w = WT > 0 ? WT : a.w;
  constexpr int WMAX = WT > 0 ? WT : kMaxW;
  T ker[DIM][WMAX];
  int i0[DIM];
  double phase = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const T xs = a.x[d][s];
    const T xm = xs * (T)bp.smul;                        // fl(topo * freq), reference :990-992
    const double xr = ((double)xm - bp.C[d]) * bp.invgam[d];
    if (a.prephase) phase += bp.D[d] * (double)xs;
    const double g = fold_grid(xr, a.nf[d]);
    const double gi = ceil(g - 0.5 * (double)w);
    i0[d] = (int)gi;
    const T z0 = (T)(gi - g);
#pragma unroll
    for (int j = 0; j < WMAX; ++j)
      if (j < w) ker[d][j] = es_kernel<T>(z0 + (T)j, a.beta, a.c, a.halfw);
  }
  C ph = make_c<T>(T(1), T(0));
  if (a.prephase) { double sn, cs; sincos(phase, &sn, &cs); ph = make_c<T>((T)cs, (T)sn); }
  const int64_t plane = (int64_t)a.nf[0] * a.nf[1] * (DIM == 3 ? a.nf[2] : 1);
  for (int p = 0; p < a.ntr; ++p) {
    C cw = a.W[((int64_t)b * a.ntr + p) * a.n_cap + s];
    if (a.prephase) cw = cmul(cw, ph);
    C* g = a.grid + ((int64_t)b * a.ntr + p) * plane;
    if (DIM == 2) {
#pragma unroll
      for (int j2 = 0; j2 < WMAX; ++j2) {
        if (j2 < w) {
          const int r = wrap_idx(i0[1] + j2, a.nf[1]);
          C* row = g + (int64_t)r * a.nf[0];
          const C c2 = make_c<T>(cw.x * ker[1][j2], cw.y * ker[1][j2]);
#pragma unroll
          for (int j1 = 0; j1 < WMAX; ++j1) {
            if (j1 < w) {
              const int cidx = wrap_idx(i0[0] + j1, a.nf[0]);
              atomic_add_c(row + cidx, make_c<T>(c2.x * ker[0][j1], c2.y * ker[0][j1]));
            }
          }
        }
      }
    } else {
      for (int j3 = 0; j3 < w; ++j3) {
        const int pz = wrap_idx(i0[DIM - 1] + j3, a.nf[DIM - 1]);
        const T k3 = ker[DIM - 1][j3];
#pragma unroll
        for (int j2 = 0; j2 < WMAX; ++j2) {
          if (j2 < w) {
            const int r = wrap_idx(i0[1] + j2, a.nf[1]);
            C* row = g + ((int64_t)pz * a.nf[1] + r) * a.nf[0];
            const T k23 = ker[1][j2] * k3;
            const C c2 = make_c<T>(cw.x * k23, cw.y * k23);
#pragma unroll
            for (int j1 = 0; j1 < WMAX; ++j1) {
              if (j1 < w) {
                const int cidx = wrap_idx(i0[0] + j1, a.nf[0]);
                atomic_add_c(row + cidx, make_c<T>(c2.x * ker[0][j1], c2.y * ker[0][j1]));
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// type 1: deconvolve + gather the requested integer modes straight into the visibility array
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_modes_kernel(const cplx_t<T>* __restrict__ ghat, int nf, int ntr, int half_modes,
                    const T* __restrict__ invphi, const int32_t* __restrict__ m1,
                    const int32_t* __restrict__ m2, int64_t nk, EpiDev e) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  const int bp_ = blockIdx.y;           // b * ntr + p
  const int b = bp_ / ntr, p = bp_ % ntr;
  const int a1 = m1[k], a2 = m2[k];
  // modes outside [-half, half] are not representable by this transform: emit NaN loudly
  const bool ok = abs(a1) <= half_modes && abs(a2) <= half_modes;
  const int i1 = a1 < 0 ? a1 + nf : a1, i2 = a2 < 0 ? a2 + nf : a2;
  cplx_t<T> v;
  if (ok) {
    v = ghat[((int64_t)bp_ * nf + i2) * nf + i1];
    const T s = invphi[abs(a1)] * invphi[abs(a2)];
    v.x *= s; v.y *= s;
  } else {
    v = make_c<T>((T)NAN, (T)NAN);
  }
  epilogue_store(e, b, p, k, v);
}

// ------------------------------------------------------------------------------------------------
// type 3, step 2a: deconvolve the spread grid (as Fourier coefficients, index i <-> mode i - nf/2)
// into the zero-padded FFT grid.  One thread per FFT-grid cell (coalesced full overwrite).
// ------------------------------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
deconv_pad_kernel(const cplx_t<T>* __restrict__ fw, cplx_t<T>* __restrict__ fw2, int nf1, int nf2,
                  int nf3, int ng1, int ng2, int ng3, const T* __restrict__ inv1,
                  const T* __restrict__ inv2, const T* __restrict__ inv3) {
  const int64_t cells = (int64_t)ng1 * ng2 * ng3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cells) return;
  const int bp_ = blockIdx.y;
  const int j1 = (int)(i % ng1);
  const int j2 = (int)((i / ng1) % ng2);
  const int j3 = DIM == 3 ? (int)(i / ((int64_t)ng1 * ng2)) : 0;
  // FFT-grid index j <-> mode m = j (j < ng/2) or j - ng; mode kept if -nf/2 <= m < nf/2
  const int m1 = j1 < ng1 / 2 ? j1 : j1 - ng1;
  const int m2 = j2 < ng2 / 2 ? j2 : j2 - ng2;
  const int m3 = DIM == 3 ? (j3 < ng3 / 2 ? j3 : j3 - ng3) : 0;
  bool in = m1 >= -nf1 / 2 && m1 < nf1 / 2 && m2 >= -nf2 / 2 && m2 < nf2 / 2;
  if (DIM == 3) in = in && m3 >= -nf3 / 2 && m3 < nf3 / 2;
  cplx_t<T> v = make_c<T>(T(0), T(0));
  if (in) {
    const int s1 = m1 + nf1 / 2, s2 = m2 + nf2 / 2, s3 = DIM == 3 ? m3 + nf3 / 2 : 0;
    const int64_t src = ((int64_t)s3 * nf2 + s2) * nf1 + s1;
    v = fw[(int64_t)bp_ * ((int64_t)nf1 * nf2 * nf3) + src];
    T sc = inv1[s1] * inv2[s2];
    if (DIM == 3) sc *= inv3[s3];
    v.x *= sc; v.y *= sc;
  }
  fw2[(int64_t)bp_ * cells + i] = v;
}

// ------------------------------------------------------------------------------------------------
// type 3, step 2b: interpolate the FFT grid at the rescaled targets, divide by the kernel's
// Fourier transform at the target frequency, apply the post-phase and store.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct InterpArgs {
  const T* u[3];
  int64_t nk;
  int ng[3];
  int w;
  T beta, c, halfw;
  int ntr;
  int postphase;
  const cplx_t<T>* fw2;        // (nb, ntr, ng3, ng2, ng1)
  const BatchParams* bp;
  Quad quad;
  EpiDev epi;
};

template <typename T, int DIM, int WT>
__global__ void __launch_bounds__(128)
interp_kernel(InterpArgs<T> a) {
  using C = cplx_t<T>;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.nk) return;
  const int b = blockIdx.y;
  const BatchParams bp = a.bp[b];
  const int w = WT > 0 ? WT : a.w;
  constexpr int WMAX = WT > 0 ? WT : kMaxW;
  T ker[DIM][WMAX];
  int i0[DIM];
  double phase = 0.0, phihat = 1.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const T um = a.u[d][k] * (T)bp.tmul;                  // uvw = bls * freq, reference :973
    const double sp = bp.hgam[d] * ((double)um - bp.D[d]);
    phase += ((double)um - bp.D[d]) * bp.C[d];
    double ft = 0.0;
    for (int n = 0; n < a.quad.q; ++n) ft += a.quad.f[n] * 2.0 * cos(sp * a.quad.z[n]);
    phihat *= ft;
    const double g = fold_grid(sp, a.ng[d]);
    const double gi = ceil(g - 0.5 * (double)w);
    i0[d] = (int)gi;
    const T z0 = (T)(gi - g);
#pragma unroll
    for (int j = 0; j < WMAX; ++j)
      if (j < w) ker[d][j] = es_kernel<T>(z0 + (T)j, a.beta, a.c, a.halfw);
  }
  double sn = 0.0, cs = 1.0;
  if (a.postphase) sincos(phase, &sn, &cs);
  const double inv = 1.0 / phihat;
  const C dec = make_c<T>((T)(cs * inv), (T)(sn * inv));
  const int64_t cells = (int64_t)a.ng[0] * a.ng[1] * (DIM == 3 ? a.ng[2] : 1);
  for (int p = 0; p < a.ntr; ++p) {
    const C* g = a.fw2 + ((int64_t)b * a.ntr + p) * cells;
    C acc = make_c<T>(T(0), T(0));
    const int n3 = DIM == 3 ? w : 1;
    for (int j3 = 0; j3 < n3; ++j3) {
      const int pz = DIM == 3 ? wrap_idx(i0[DIM - 1] + j3, a.ng[DIM - 1]) : 0;
      const T k3 = DIM == 3 ? ker[DIM - 1][j3] : T(1);
#pragma unroll
      for (int j2 = 0; j2 < WMAX; ++j2) {
        if (j2 < w) {
          const int r = wrap_idx(i0[1] + j2, a.ng[1]);
          const C* row = g + ((int64_t)pz * a.ng[1] + r) * a.ng[0];
          C racc = make_c<T>(T(0), T(0));
#pragma unroll
          for (int j1 = 0; j1 < WMAX; ++j1) {
            if (j1 < w) {
              const C v = row[wrap_idx(i0[0] + j1, a.ng[0])];
              racc.x += v.x * ker[0][j1];
              racc.y += v.y * ker[0][j1];
            }
          }
          const T k23 = ker[1][j2] * k3;
          acc.x += racc.x * k23;
          acc.y += racc.y * k23;
        }
      }
    }
    epilogue_store(a.epi, b, p, k, cmul(acc, dec));
  }
}

// min / max of the live part of an array (type 3 widths when the caller does not supply them)
template <typename T>
__global__ void minmax_kernel(const T* __restrict__ x, const int32_t* __restrict__ n_dev, int64_t n_fixed,
                              double* __restrict__ out /* {min,max}, pre-initialised */) {
  const int64_t n = n_dev ? (int64_t)*n_dev : n_fixed;
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    lo = fmin(lo, v); hi = fmax(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    // atomic min/max on doubles through their ordered integer image
    auto enc = [](double d) { long long i = __double_as_longlong(d); return i >= 0 ? i : i ^ 0x7fffffffffffffffLL; };
    atomicMin((long long*)out, enc(lo));
    atomicMax((long long*)out + 1, enc(hi));
  }
}

// direct sum on the GPU (fp64 phases and accumulation) -- validation aid / crossover baseline
template <typename T, int DIM>
__global__ void __launch_bounds__(128)
direct_sum_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                  const int32_t* __restrict__ n_dev, int64_t n_cap, const T* __restrict__ u,
                  const T* __restrict__ v, const T* __restrict__ wv, int64_t nk,
                  const BatchParams* __restrict__ bps, int ntr, const cplx_t<T>* __restrict__ W, EpiDev e) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  const int b = blockIdx.y;
  const int n = *n_dev;
  const T mul = (T)bps[b].tmul;
  const double uk = (double)(u[k] * mul), vk = (double)(v[k] * mul), wk = DIM == 3 ? (double)(wv[k] * mul) : 0.0;
  double ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
  for (int s = 0; s < n; ++s) {
    double ph = uk * (double)x[s] + vk * (double)y[s];
    if (DIM == 3) ph += wk * (double)z[s];
    double sn, cs;
    sincos(ph, &sn, &cs);
    for (int p = 0; p < ntr; ++p) {
      const cplx_t<T> c = W[((int64_t)b * ntr + p) * n_cap + s];
      ar[p] += (double)c.x * cs - (double)c.y * sn;
      ai[p] += (double)c.x * sn + (double)c.y * cs;
    }
  }
  for (int p = 0; p < ntr; ++p) epilogue_store(e, b, p, k, make_c<T>((T)ar[p], (T)ai[p]));
}

template <typename T>
__global__ void __launch_bounds__(256)
basis_contract_kernel(const cplx_t<T>* __restrict__ vkl, int64_t nk, const cplx_t<T>* __restrict__ coefs,
                      int K, int64_t nfreq_total, int64_t f0, int kk, int ll,
                      const int32_t* __restrict__ ant1, const int32_t* __restrict__ ant2, EpiDev e) {
  using C = cplx_t<T>;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  const int b = blockIdx.y;
  const int64_t f = f0 + b;
  const int a1 = ant1[k], a2 = ant2[k];
  const C c1k = coefs[((int64_t)a1 * K + kk) * nfreq_total + f], c1l = coefs[((int64_t)a1 * K + ll) * nfreq_total + f];
  const C c2k = coefs[((int64_t)a2 * K + kk) * nfreq_total + f], c2l = coefs[((int64_t)a2 * K + ll) * nfreq_total + f];
  const C wkl = cmulc(c1k, c2l);      // conj(c[a1,k]) c[a2,l]
  const C wlk = cmulc(c1l, c2k);      // conj(c[a1,l]) c[a2,k]
  C v[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) v[p] = vkl[((int64_t)b * 4 + p) * nk + k];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      C r = cmul(wkl, v[i * 2 + j]);
      if (kk != ll) r = cadd(r, cmul(wlk, v[j * 2 + i]));
      epilogue_store(e, b, i * 2 + j, k, r);
    }
}

// all K (K + 1) / 2 pairs in one pass: vkl (nb, npairs * 4, nk), pair order k <= l row-major
template <typename T>
__global__ void __launch_bounds__(256)
basis_contract_all_kernel(const cplx_t<T>* __restrict__ vkl, int64_t nk, const cplx_t<T>* __restrict__ coefs,
                          int K, int64_t nfreq_total, int64_t f0, const int32_t* __restrict__ ant1,
                          const int32_t* __restrict__ ant2, EpiDev e) {
  using C = cplx_t<T>;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  const int b = blockIdx.y;
  const int64_t f = f0 + b;
  const int a1 = ant1[k], a2 = ant2[k];
  const int npairs = K * (K + 1) / 2;
  C acc[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) acc[p] = make_c<T>(T(0), T(0));
  int q = 0;
  for (int kk = 0; kk < K; ++kk) {
    const C c1k = coefs[((int64_t)a1 * K + kk) * nfreq_total + f], c2k = coefs[((int64_t)a2 * K + kk) * nfreq_total + f];
    for (int ll = kk; ll < K; ++ll, ++q) {
      const C c1l = coefs[((int64_t)a1 * K + ll) * nfreq_total + f], c2l = coefs[((int64_t)a2 * K + ll) * nfreq_total + f];
      const C wkl = cmulc(c1k, c2l);      // conj(c[a1,k]) c[a2,l]
      const C wlk = cmulc(c1l, c2k);      // conj(c[a1,l]) c[a2,k]
      C v[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) v[p] = vkl[((int64_t)b * npairs * 4 + q * 4 + p) * nk + k];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // same order of operations as the per-pair kernel: r = wkl v (+ wlk v^T), then accumulate
          C r = cmul(wkl, v[i * 2 + j]);
          if (kk != ll) r = cadd(r, cmul(wlk, v[j * 2 + i]));
          acc[i * 2 + j] = cadd(acc[i * 2 + j], r);
        }
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) epilogue_store(e, b, p, k, acc[p]);
}

}  // namespace fv

#include "type1_fused.cuh"

// ================================================================================================
// plan object
// ================================================================================================
struct fv_plan {
  cudaStream_t stream = nullptr;
  std::map<std::tuple<int, int64_t, int64_t, int64_t, int64_t>, cufftHandle> ffts;  // (prec, n3, n2, n1, batch)
  std::map<std::tuple<int, int64_t, int64_t, int, double>, void*> invphi;            // (prec, nf, nfft, w, beta)
  void* grid = nullptr;   size_t grid_bytes = 0;
  void* grid2 = nullptr;  size_t grid2_bytes = 0;
  fv::BatchParams* bp_dev = nullptr; int bp_cap = 0;
  double* lim_dev = nullptr;
  size_t fft_work_bytes = 0;
  size_t table_bytes = 0;
  bool timing = false;                       // CUDA events around every stage launch
  int timing_mask = (1 << FV_STAGE_COUNT) - 1; // ... of the stages whose bit is set (option "timing_mask")
  double stage_ms[FV_STAGE_COUNT] = {0};
  int64_t stage_n[FV_STAGE_COUNT] = {0};
  struct Pending { int stage; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
  size_t max_grid_bytes = (size_t)96 << 30;   // refuse grids beyond this (B200 has 180 GB)
  // shared-memory FFT plans of the fused type-1 path, keyed by (prec, nf)
  struct SmemFft { fv::FftStages st; void* tw = nullptr; std::vector<int> pos; int32_t* pos_dev = nullptr;
                   void* wn = nullptr; };   // wn: exp(+2 pi i j / (2 n)), j < n (half-length type-3 passes)
  std::map<std::pair<int, int64_t>, SmemFft> smem_ffts;
  void* tbuf = nullptr; size_t tbuf_bytes = 0;   // half-transformed array T of the fused type-1 path
  void* prep = nullptr; size_t prep_bytes = 0;   // folded NU points (ix0, iy0, zx, zy) of the current batch
  void* bins = nullptr; size_t bins_bytes = 0;   // type-3 tile lists: counts, offsets, cursor, list
  void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
  int t3_tiles = 1;                              // 0 disables the tiled type-3 spreader
  void* grid3 = nullptr; size_t grid3_bytes = 0; // intermediate of the pruned type-3 FFT passes
  int t3_fft = 1;                                // 0: cuFFT on the padded grid; 1: own pruned shared-memory passes for 3-D
                                                 // (where they measure faster), cuFFT for 2-D; 2: own passes always
  int t3_v[3] = {0, 0, 0}, t3_thr[3] = {0, 0, 0}; // tuning overrides: vectors per CTA / threads of the x, y, z passes
  int t3_half = 6;                               // bit d: half-length pass along dimension d where ng = 2 nf (0: full-length forms); measured on cfg4: y + z
  int t3_minb[3] = {0, 0, 0};                    // CTAs per SM the half-length y / z passes are compiled for (0: automatic)
  int t1_rows = 0;                               // strip height override (0 = automatic)
  int t1_cols = 0;                               // columns per CTA override (0 = automatic)
  // small-grid type-1 path (type1_small.cuh): sort buffers, records, per-(nf, w) phase schedules
  void* small = nullptr; size_t small_bytes = 0;
  void* rec = nullptr; size_t rec_bytes = 0;
  int t1_xdirect = 1;                            // 0: never; 1: single precision, grid of several strips (type1_xdirect.cuh)
  int t1_small = 1;                              // 0: never; 1: automatic (whole grid in one CTA and >= 4096 source slots); 2: whenever possible
  struct SmallSched { int nphase = 0; int32_t* ph_off = nullptr; uint16_t* ph_bins = nullptr; };
  std::map<std::pair<int64_t, int>, SmallSched> small_scheds;
  // geometry of the last type-3 transform: dim, w, nf[3] (spread grid), ng[3] (FFT grid), tiled, own_fft, sub-batch
  int64_t last_geo[12] = {0};
};

// baselines of one beam pair as integer modes, bucketed by first mode number (fused type-1 path)
struct fv_modeset {
  std::vector<int32_t> m1, m2;
  int n_modes = 0;
  struct Tables {
    int ncols = 0;
    int32_t* col_pos = nullptr; int32_t* col_off = nullptr; int32_t* s_k = nullptr; int32_t* s_pos = nullptr;
    void* s_scale = nullptr;
    int32_t* col_k = nullptr;      // signed first mode number of every column (x-direct pass 1)
    void* s_scale_y = nullptr;     // 1 / phihat(m2) only: the x-direct pass 1 has no kernel to divide out along x
  };
  std::map<std::tuple<int, int64_t, int, double>, Tables> tables;   // (prec, nf, w, beta)
};

namespace fv {

// x-direct pass 1 (type1_xdirect.cu)
bool t1_xdirect_built(int w);
int t1_xdirect_rows();
int t1_xdirect_pass1_entry(fv_plan* P, const void* bx, const void* by, const int32_t* n_dev, int64_t n_cap, int nb, int ntr,
                           const void* W, int64_t nf, int w, double beta, const fv_modeset::Tables* tab);

// small-grid type-1 path (type1_small.cu)
bool t1s_width_built(int w);
int t1_small_pass1_entry(fv_plan* P, int prec, const int32_t* n_dev, int64_t n_cap, int nb, int ntr, const void* W,
                         int64_t nf, int w, double beta, const int32_t* ix0, const int32_t* iy0, const void* zx,
                         const void* zy, const fv_plan::SmemFft* F, const fv_modeset::Tables* tab);

static int ensure(void** p, size_t* have, size_t need) {
  if (*have >= need) return FV_OK;
  if (*p) { cudaError_t e = cudaFree(*p); if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; } *p = nullptr; *have = 0; }
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("device allocation of " + std::to_string(need) + " bytes failed: " + cudaGetErrorString(e));
    return FV_ERR_ALLOC;
  }
  *have = need;
  return FV_OK;
}

static int get_fft(fv_plan* P, int prec, int dim, int64_t n1, int64_t n2, int64_t n3, int64_t batch, cufftHandle* h) {
  auto key = std::make_tuple(prec, dim == 3 ? n3 : (int64_t)1, n2, n1, batch);
  auto it = P->ffts.find(key);
  if (it != P->ffts.end()) { *h = it->second; return FV_OK; }
  cufftHandle plan;
  if (cufftCreate(&plan) != CUFFT_SUCCESS) { set_error("cufftCreate failed"); return FV_ERR_CUFFT; }
  long long dims[3];
  int rank = dim;
  if (dim == 3) { dims[0] = n3; dims[1] = n2; dims[2] = n1; } else { dims[0] = n2; dims[1] = n1; }
  long long dist = n1 * n2 * (dim == 3 ? n3 : 1);
  size_t work = 0;
  cufftResult r = cufftMakePlanMany64(plan, rank, dims, nullptr, 1, dist, nullptr, 1, dist,
                                      prec == 1 ? CUFFT_C2C : CUFFT_Z2Z, batch, &work);
  if (r != CUFFT_SUCCESS) {
    cufftDestroy(plan);
    set_error("cufftMakePlanMany64 failed with code " + std::to_string((int)r) + " for grid " +
              std::to_string(n1) + "x" + std::to_string(n2) + "x" + std::to_string(n3) + " batch " + std::to_string(batch));
    return FV_ERR_CUFFT;
  }
  cufftSetStream(plan, P->stream);
  P->fft_work_bytes += work;
  P->ffts[key] = plan;
  *h = plan;
  return FV_OK;
}

// RAII stage timer: records an event pair on the plan's stream around the launches in its scope
struct StageScope {
  fv_plan* P; int stage; cudaEvent_t e0 = nullptr, e1 = nullptr;
  static cudaEvent_t get(fv_plan* P) {
    if (!P->event_pool.empty()) { cudaEvent_t e = P->event_pool.back(); P->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  StageScope(fv_plan* P_, int stage_) : P(P_), stage(stage_) {
    if (P->timing && ((P->timing_mask >> stage_) & 1)) { e0 = get(P); e1 = get(P); cudaEventRecord(e0, P->stream); }
  }
  ~StageScope() {
    if (e0) { cudaEventRecord(e1, P->stream); P->pending.push_back({stage, e0, e1}); }
  }
};

static int run_fft(fv_plan* P, cufftHandle h, int prec, void* data) {
  StageScope ts(P, FV_STAGE_FFT);
  cufftResult r = prec == 1 ? cufftExecC2C(h, (cufftComplex*)data, (cufftComplex*)data, CUFFT_INVERSE)
                            : cufftExecZ2Z(h, (cufftDoubleComplex*)data, (cufftDoubleComplex*)data, CUFFT_INVERSE);
  if (r != CUFFT_SUCCESS) { set_error("cufftExec failed with code " + std::to_string((int)r)); return FV_ERR_CUFFT; }
  return FV_OK;
}

// device table of 1/phihat in working precision; `centered` tables are indexed by grid index
// i <-> mode i - nf/2 (type 3 step 2a), plain ones by |mode| (type 1)
template <typename T>
static int get_invphi(fv_plan* P, int prec, int64_t nf_index, int64_t nfft, int w, double beta, bool centered, const T** out) {
  auto key = std::make_tuple(prec + (centered ? 10 : 0), nf_index, nfft, w, beta);
  auto it = P->invphi.find(key);
  if (it != P->invphi.end()) { *out = (const T*)it->second; return FV_OK; }
  Quad Q = make_quad(w, beta);
  std::vector<double> ph = kernel_ft_series(nfft, Q);
  std::vector<T> host;
  if (centered) {
    host.resize(nf_index);
    for (int64_t i = 0; i < nf_index; ++i) { int64_t m = i - nf_index / 2; host[i] = (T)(1.0 / ph[m < 0 ? -m : m]); }
  } else {
    host.resize(nf_index);     // nf_index = number of |mode| entries wanted
    for (int64_t k = 0; k < nf_index; ++k) host[k] = (T)(1.0 / ph[k]);
  }
  void* d = nullptr;
  FV_CUDA(cudaMalloc(&d, host.size() * sizeof(T)));
  FV_CUDA(cudaMemcpyAsync(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, P->stream));
  FV_CUDA(cudaStreamSynchronize(P->stream));   // host vector goes out of scope
  P->table_bytes += host.size() * sizeof(T);
  P->invphi[key] = d;
  *out = (const T*)d;
  return FV_OK;
}

static int upload_bp(fv_plan* P, const std::vector<BatchParams>& bp) {
  if ((int)bp.size() > P->bp_cap) {
    if (P->bp_dev) FV_CUDA(cudaFree(P->bp_dev));
    P->bp_cap = std::max<int>(256, (int)bp.size());
    FV_CUDA(cudaMalloc((void**)&P->bp_dev, sizeof(BatchParams) * P->bp_cap));
  }
  // pageable source: the driver stages it before returning, so `bp` may die afterwards
  FV_CUDA(cudaMemcpyAsync(P->bp_dev, bp.data(), sizeof(BatchParams) * bp.size(), cudaMemcpyHostToDevice, P->stream));
  return FV_OK;
}

static EpiDev make_epi(const fv_epilogue* e) {
  EpiDev d;
  d.out = e->out; d.sb = e->out_stride_b; d.sp = e->out_stride_p;
  for (int i = 0; i < 4; ++i) d.pmap[i] = e->pmap[i];
  d.kmap = e->kmap; d.conj_flag = e->conj_flag; d.acc = e->accumulate;
  return d;
}

#define FV_DISPATCH_W(WV, CALL)                         \
  switch (WV) {                                         \
    case 7: { constexpr int WT = 7; CALL; } break;      \
    case 9: { constexpr int WT = 9; CALL; } break;      \
    case 11: { constexpr int WT = 11; CALL; } break;    \
    case 13: { constexpr int WT = 13; CALL; } break;    \
    case 14: { constexpr int WT = 14; CALL; } break;    \
    default: { constexpr int WT = 0; CALL; } break;     \
  }

template <typename T>
static int launch_spread(fv_plan* P, int dim, SpreadArgs<T>& a, int nb) {
  if (a.n_cap == 0) return FV_OK;
  StageScope ts(P, FV_STAGE_SPREAD);
  dim3 grid(ceil_div(a.n_cap, 128), nb);
  if (dim == 2) { FV_DISPATCH_W(a.w, (spread_kernel<T, 2, WT><<<grid, 128, 0, P->stream>>>(a))); }
  else { FV_DISPATCH_W(a.w, (spread_kernel<T, 3, WT><<<grid, 128, 0, P->stream>>>(a))); }
  FV_LAUNCH_CHECK();
  return FV_OK;
}

template <typename T>
static int launch_interp(fv_plan* P, int dim, InterpArgs<T>& a, int nb) {
  StageScope ts(P, FV_STAGE_INTERP);
  dim3 grid(ceil_div(a.nk, 128), nb);
  if (dim == 2) { FV_DISPATCH_W(a.w, (interp_kernel<T, 2, WT><<<grid, 128, 0, P->stream>>>(a))); }
  else { FV_DISPATCH_W(a.w, (interp_kernel<T, 3, WT><<<grid, 128, 0, P->stream>>>(a))); }
  FV_LAUNCH_CHECK();
  return FV_OK;
}

static int get_smem_fft(fv_plan* P, int prec, int64_t nf, fv_plan::SmemFft** out) {
  auto key = std::make_pair(prec, nf);
  auto it = P->smem_ffts.find(key);
  if (it != P->smem_ffts.end()) { *out = &it->second; return FV_OK; }
  fv_plan::SmemFft f;
  // factor order: 8s, a 4, a 2, then 15s, 5s, 3s (odd radices last keep the late, short-stride
  // stages free of shared-memory bank conflicts)
  int64_t n = nf;
  std::vector<int> rad;
  while (n % 8 == 0) { rad.push_back(8); n /= 8; }
  while (n % 4 == 0) { rad.push_back(4); n /= 4; }
  while (n % 2 == 0) { rad.push_back(2); n /= 2; }
  while (n % 15 == 0) { rad.push_back(15); n /= 15; }
  while (n % 5 == 0) { rad.push_back(5); n /= 5; }
  while (n % 3 == 0) { rad.push_back(3); n /= 3; }
  if (n != 1 || (int)rad.size() > T1_MAX_STAGES || nf >= 65536) {
    set_error("fused type-1 path needs a 2-3-5-smooth grid size below 65536");
    return FV_ERR_UNSUPPORTED;
  }
  f.st.nstage = (int)rad.size();
  int64_t cur = nf;
  for (int i = 0; i < f.st.nstage; ++i) {
    f.st.radix[i] = rad[i];
    const int64_t m = cur / rad[i];
    f.st.inv_m[i] = m == 1 ? 0u : (unsigned)(((1ull << 32) / (unsigned long long)m) + 1ull);
    f.st.hw[i] = (m >= 9 && m <= 15 && (m & 1)) ? (int)m : 16;
    cur = m;
  }
  // digit-reversed output positions
  f.pos.resize(nf);
  for (int64_t k = 0; k < nf; ++k) {
    int64_t kk = k, wgt = nf, p = 0;
    for (int i = 0; i < f.st.nstage; ++i) { wgt /= rad[i]; p += (kk % rad[i]) * wgt; kk /= rad[i]; }
    f.pos[k] = (int)p;
  }
  // per-stage twiddle tables, laid out so that consecutive butterflies read consecutive words
  const size_t csz = prec == 1 ? sizeof(float2) : sizeof(double2);
  std::vector<double> twr, twi;
  cur = nf;
  for (int i = 0; i < f.st.nstage; ++i) {
    const int64_t r = rad[i], m = cur / r;
    f.st.tw_off[i] = (int)twr.size();
    if (m > 1)
      for (int64_t q = 1; q < r; ++q)
        for (int64_t j = 0; j < m; ++j) {
          const double ang = 2.0 * M_PI * (double)((j * q) % cur) / (double)cur;
          twr.push_back(cos(ang)); twi.push_back(sin(ang));
        }
    cur = m;
  }
  if (twr.empty()) { twr.push_back(1.0); twi.push_back(0.0); }
  f.st.tw_len = (int)twr.size();
  std::vector<unsigned char> host(csz * twr.size());
  for (size_t t = 0; t < twr.size(); ++t) {
    if (prec == 1) ((float2*)host.data())[t] = make_float2((float)twr[t], (float)twi[t]);
    else ((double2*)host.data())[t] = make_double2(twr[t], twi[t]);
  }
  FV_CUDA(cudaMalloc(&f.tw, host.size() + 16));      // + 16: bulk copies of the table round its size up to 16 bytes
  FV_CUDA(cudaMemsetAsync(f.tw, 0, host.size() + 16, P->stream));
  FV_CUDA(cudaMemcpyAsync(f.tw, host.data(), host.size(), cudaMemcpyHostToDevice, P->stream));
  FV_CUDA(cudaStreamSynchronize(P->stream));
  P->table_bytes += host.size();
  auto res = P->smem_ffts.emplace(key, std::move(f));
  *out = &res.first->second;
  return FV_OK;
}

}  // namespace fv
