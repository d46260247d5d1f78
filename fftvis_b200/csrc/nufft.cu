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
  const int64_t idx = (int64_t)b * e.sb + (int64_t)e.pmap[p] * e.sp + (e.kmap ? (int64_t)e.kmap[k] : k);
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

}  // namespace fv

#include "type1_fused.cuh"
#include "type3_tiles.cuh"
#include "type3_fft.cuh"

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
  double stage_ms[FV_STAGE_COUNT] = {0};
  int64_t stage_n[FV_STAGE_COUNT] = {0};
  struct Pending { int stage; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
  size_t max_grid_bytes = (size_t)96 << 30;   // refuse grids beyond this (B200 has 180 GB)
  // shared-memory FFT plans of the fused type-1 path, keyed by (prec, nf)
  struct SmemFft { fv::FftStages st; void* tw = nullptr; std::vector<int> pos; int32_t* pos_dev = nullptr; };
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
  int t1_np4 = 0;                                // 1: spread the 4 products of a small grid in one CTA (measured slower on cfg3: off)
  int t1_rows = 0;                               // strip height override (0 = automatic)
  int t1_cols = 0;                               // columns per CTA override (0 = automatic)
};

// baselines of one beam pair as integer modes, bucketed by first mode number (fused type-1 path)
struct fv_modeset {
  std::vector<int32_t> m1, m2;
  int n_modes = 0;
  struct Tables {
    int ncols = 0;
    int32_t* col_pos = nullptr; int32_t* col_off = nullptr; int32_t* s_k = nullptr; int32_t* s_pos = nullptr;
    void* s_scale = nullptr;
  };
  std::map<std::tuple<int, int64_t, int, double>, Tables> tables;   // (prec, nf, w, beta)
};

namespace fv {

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
    if (P->timing) { e0 = get(P); e1 = get(P); cudaEventRecord(e0, P->stream); }
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


// ---- type 1, fused shared-memory path ----------------------------------------------------------
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
  FV_CUDA(cudaMalloc(&f.tw, host.size()));
  FV_CUDA(cudaMemcpyAsync(f.tw, host.data(), host.size(), cudaMemcpyHostToDevice, P->stream));
  FV_CUDA(cudaStreamSynchronize(P->stream));
  P->table_bytes += host.size();
  auto res = P->smem_ffts.emplace(key, std::move(f));
  *out = &res.first->second;
  return FV_OK;
}

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
  std::vector<int32_t> col_pos, col_off, s_k(nk), s_pos(nk);
  std::vector<T> s_scale(nk);
  for (int64_t i = 0; i < nk; ++i) {
    const int32_t k = order[i];
    const int a1 = M->m1[k], a2 = M->m2[k];
    if (i == 0 || a1 != M->m1[order[i - 1]]) {
      col_off.push_back((int32_t)i);
      col_pos.push_back(F.pos[a1 < 0 ? a1 + nf : a1]);
    }
    s_k[i] = k;
    s_pos[i] = F.pos[a2 < 0 ? a2 + nf : a2];
    s_scale[i] = (T)(1.0 / (ph[abs(a1)] * ph[abs(a2)]));
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
  const size_t tneed = sizeof(C) * (size_t)nb * ntr * ncols * nf;
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
  const size_t row_bytes = sizeof(C) * pitch;
  const size_t smem_max = 227 * 1024 - 1024;
  // strip height R and CTA size: whole grid in one CTA when it fits; otherwise 16 rows x 512 threads
  // (one row per warp) if that fits, else 8 rows x 256 threads
  // strip height R: the whole grid when it fits one CTA, else as many rows as shared memory holds
  // (<= 32); one warp per strip row (256..768 threads): the row FFTs are warp tasks
  auto thr_for = [](int64_t rows) { return (int)std::min<int64_t>(t1_limits<T>::spread_threads, std::max<int64_t>(256, 32 * rows)); };
  int R, np = 1;
  const bool whole = t1_spread_fixed_smem<T>((int)nf, wmax, thr_for(nf)) + row_bytes * nf <= 200 * 1024;
  if (P->t1_rows > 0) R = (int)std::min<int64_t>(P->t1_rows, nf);
  else if (whole && ntr == 4 && P->t1_np4) {
    // small grid, four polarisation products: one CTA spreads all four (shared scan / kernel
    // evaluations / index arithmetic) on a quarter-height strip
    np = 4;
    R = (int)((nf + 3) / 4);
    while (R > 1 && t1_spread_fixed_smem<T>((int)nf, wmax, thr_for(R), 4) + 4 * row_bytes * R > smem_max) --R;
  } else if (whole) R = (int)nf;
  else {
    R = 32;
    while (R > 1 && t1_spread_fixed_smem<T>((int)nf, wmax, thr_for(R)) + row_bytes * R > smem_max) --R;
    if (R > 8) R -= R % 8;
  }
  int threads = thr_for(R);
  size_t fixed1 = t1_spread_fixed_smem<T>((int)nf, wmax, threads, np);
  while (R > 1 && fixed1 + np * row_bytes * R > smem_max) --R;
  if (fixed1 + np * row_bytes * R > smem_max) { set_error("fine-grid row does not fit shared memory: use the cuFFT type-1 path"); return FV_ERR_UNSUPPORTED; }
  // fold every (frequency, source) point once
  const size_t per = (size_t)nb * n_cap;
  rc = ensure(&P->prep, &P->prep_bytes, per * (2 * sizeof(int32_t) + 2 * sizeof(T)));
  if (rc) return rc;
  int32_t* ix0 = (int32_t*)P->prep;
  int32_t* iy0 = ix0 + per;
  T* zx = (T*)(iy0 + per);
  T* zy = zx + per;
  {
    StageScope ts(P, FV_STAGE_ZERO);
    dim3 grid(ceil_div(n_cap, 256), nb);
    t1_prep_kernel<T><<<grid, 256, 0, P->stream>>>((const T*)bx, (const T*)by, n_dev, n_cap, P->bp_dev, (int)nf, w, ix0, iy0, zx, zy);
    FV_LAUNCH_CHECK();
  }
  T1SpreadArgs<T> a{};
  a.n_dev = n_dev; a.n_cap = n_cap; a.ix0 = ix0; a.iy0 = iy0; a.zx = zx; a.zy = zy;
  a.nf = (int)nf; a.R = R; a.pitch = pitch; a.w = w;
  a.beta = (T)beta; a.c = (T)(4.0 / ((double)w * w)); a.halfw = (T)(w / 2.0);
  a.ntr = ntr; a.W = (const C*)W; a.tw = (const C*)F->tw; a.st = F->st;
  a.ncols = ncols; a.col_pos = tab->col_pos; a.Tbuf = (C*)P->tbuf;
  {
    StageScope ts(P, FV_STAGE_SPREAD);
    dim3 grid(ceil_div(nf, R), np == 4 ? nb : nb * ntr);
    const size_t smem = fixed1 + np * row_bytes * R;
    static const bool dbg = getenv("FV_DEBUG") != nullptr;
    static long long* dbg_dev = nullptr;
    if (dbg) {
      if (!dbg_dev) { cudaMalloc((void**)&dbg_dev, 96); }
      cudaMemsetAsync(dbg_dev, 0, 96, P->stream);
      a.dbg = dbg_dev;
    }
    if (dbg) fprintf(stderr, "[fv] t1 fused: nf=%lld w=%d ncols=%d R=%d threads=%d smem=%zu grid=(%u,%u)\n",
                     (long long)nf, w, ncols, R, threads, smem, grid.x, grid.y);
    if (np == 4) { FV_DISPATCH_W(w, (rc = launch_t1_spread<T, WT, 4>(P, a, grid, threads, smem))); }
    else { FV_DISPATCH_W(w, (rc = launch_t1_spread<T, WT, 1>(P, a, grid, threads, smem))); }
    if (rc) return rc;
    if (dbg) {
      long long hcyc[12];
      cudaMemcpyAsync(hcyc, dbg_dev, 96, cudaMemcpyDeviceToHost, P->stream);
      cudaStreamSynchronize(P->stream);
      const double nw = 8.0 * (threads / 32);
      fprintf(stderr, "[fv] t1 pass-1 cycles/warp: zero %.0f scan %.0f fill %.0f spread %.0f fft %.0f fftwait %.0f write %.0f hits/strip %.0f passA %.0f passB %.0f hits/warp %.1f passB-nonempty-frac %.3f passB-readL-cycles %.0f\n",
              hcyc[0] / nw, hcyc[1] / nw, hcyc[2] / nw, hcyc[3] / nw, hcyc[4] / nw, hcyc[5] / nw, hcyc[6] / nw, hcyc[7] / nw, hcyc[8] / nw, hcyc[9] / nw, hcyc[10] / nw, (double)(hcyc[11] >> 32) / nw, (double)(hcyc[11] & 0xffffffffll) / nw);
    }
  }
  // ---- pass 2: FFT along y + deconvolve + gather -----------------------------------------------
  const size_t fixed2 = sizeof(C) * nf;
  int cpc = P->t1_cols > 0 ? P->t1_cols : (int)std::max<size_t>(1, (100 * 1024 - std::min<size_t>(fixed2, 99 * 1024)) / row_bytes);
  cpc = std::min(cpc, 16);
  if (cpc >= 8) cpc -= cpc % 8;
  cpc = std::min(cpc, ncols);
  while (cpc > 1 && fixed2 + row_bytes * cpc > smem_max) --cpc;
  T1GatherArgs<T> g{};
  g.Tbuf = (const C*)P->tbuf; g.nf = (int)nf; g.pitch = pitch; g.ncols = ncols; g.cols_per_cta = cpc; g.ntr = ntr;
  g.tw = (const C*)F->tw; g.st = F->st; g.col_off = tab->col_off; g.s_k = tab->s_k; g.s_pos = tab->s_pos;
  g.s_scale = (const T*)tab->s_scale; g.epi = make_epi(epi);
  {
    StageScope ts(P, FV_STAGE_GATHER);
    auto kern = t1_ffty_gather_kernel<T>;
    const size_t smem = fixed2 + row_bytes * cpc;
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(ncols, cpc), nb * ntr);
    const int gthreads = std::min(T1_THREADS, std::max(64, 32 * cpc));     // one warp per column
    kern<<<grid, gthreads, smem, P->stream>>>(g);
    FV_LAUNCH_CHECK();
  }
  return FV_OK;
}

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
  int ntiles = 0;
  if (tiled) {
    geo.x = xs[0]; geo.y = xs[1]; geo.z = xs[2]; geo.n_dev = n_dev; geo.w = w;
    for (int d = 0; d < 3; ++d) { geo.C[d] = Cc[d]; geo.invgam[d] = 1.0 / gam[d]; geo.nf[d] = (int)nf[d]; }
    geo.ntx = ceil_div(nf[0], T3_TILE); geo.nty = ceil_div(nf[1], T3_TILE);
    ntiles = geo.ntx * geo.nty;
    const size_t nt1 = (size_t)ntiles + 1;
    const size_t need = sizeof(int32_t) * (3 * nt1 + 16 * (size_t)n_cap);
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
  bool own_fft = (P->t3_fft == 2 || (P->t3_fft == 1 && dim == 3)) && vx >= 1 && vy >= 1 && vz >= 1;
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
  const size_t cells3 = dim == 3 ? (size_t)nf[2] * ng[1] * ng[0] : (size_t)nf[1] * ng[0];

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
      ta.offsets = bin_offsets; ta.list = bin_list; ta.grid = (C*)P->grid;
      StageScope ts(P, FV_STAGE_SPREAD);
      dim3 tg(ntiles, sub * ntr);
      FV_DISPATCH_W(w, (t3_col_spread_kernel<T, WT><<<tg, T3_TILE * T3_TILE, 0, P->stream>>>(ta)));
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
        const size_t smem = sizeof(C) * ((size_t)vx * (ng[0] + 1) + ng[0]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_contig_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g(ceil_div(fa.nvec, vx), q);
        t3_fft_contig_kernel<T><<<g, thx, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
      }
      {
        T3FftArgs<T> fa{};
        fa.in = A1; fa.out = dim == 3 ? (C*)P->grid3 : (C*)P->grid2; fa.nin = (int)nf[1]; fa.n = (int)ng[1]; fa.nvec_cta = vy;
        fa.ninner = (int)ng[0]; fa.nouter = (int)nf[2];
        fa.in_q = dim == 3 ? (int64_t)cells2 : nf[1] * ng[0]; fa.in_a = nf[1] * ng[0]; fa.in_k = ng[0];
        fa.out_q = dim == 3 ? nf[2] * ng[1] * ng[0] : (int64_t)cells2; fa.out_a = ng[1] * ng[0]; fa.out_k = ng[0];
        fa.tw = (const C*)F[1]->tw; fa.st = F[1]->st; fa.pos = F[1]->pos_dev;
        const size_t smem = sizeof(C) * ((size_t)vy * (ng[1] + 1) + ng[1]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g((unsigned)(nf[2] * ceil_div(ng[0], vy)), q);
        t3_fft_strided_kernel<T><<<g, thy, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
      }
      if (dim == 3) {
        T3FftArgs<T> fa{};
        fa.in = (const C*)P->grid3; fa.out = (C*)P->grid2; fa.nin = (int)nf[2]; fa.n = (int)ng[2]; fa.nvec_cta = vz;
        fa.ninner = (int)(ng[1] * ng[0]); fa.nouter = 1;
        fa.in_q = nf[2] * ng[1] * ng[0]; fa.in_a = 0; fa.in_k = ng[1] * ng[0];
        fa.out_q = (int64_t)cells2; fa.out_a = 0; fa.out_k = ng[1] * ng[0];
        fa.tw = (const C*)F[2]->tw; fa.st = F[2]->st; fa.pos = F[2]->pos_dev;
        const size_t smem = sizeof(C) * ((size_t)vz * (ng[2] + 1) + ng[2]);
        FV_CUDA(cudaFuncSetAttribute(t3_fft_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g((unsigned)ceil_div(ng[1] * ng[0], vz), q);
        t3_fft_strided_kernel<T><<<g, thz, smem, P->stream>>>(fa);
        FV_LAUNCH_CHECK();
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
  for (auto& kv : P->smem_ffts) { cudaFree(kv.second.tw); if (kv.second.pos_dev) cudaFree(kv.second.pos_dev); }
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
  return (int64_t)(P->grid_bytes + P->grid2_bytes + P->tbuf_bytes + P->prep_bytes + P->bins_bytes + P->grid3_bytes + P->scan_tmp_bytes + P->fft_work_bytes + P->table_bytes);
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
  }
  delete M;
  return FV_OK;
}

extern "C" int fv_plan_set_option(fv_plan* P, const char* name, int64_t value) {
  FV_REQUIRE(P && name, "null pointer");
  const std::string n(name);
  if (n == "t1_rows") P->t1_rows = (int)value;
  else if (n == "t1_cols") P->t1_cols = (int)value;
  else if (n == "max_grid_bytes") P->max_grid_bytes = (size_t)value;
  else if (n == "t3_tiles") P->t3_tiles = (int)value;
  else if (n == "t1_np4") P->t1_np4 = (int)value;
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
  FV_REQUIRE(ntr >= 1 && ntr <= 4, "ntr must be 1..4");
  FV_REQUIRE(nb >= 0 && (int64_t)nb * ntr <= 65535, "batch too large");
  FV_REQUIRE(upsampfac > 1.0 && eps > 0, "bad eps / upsampfac");
  if (nb == 0 || modes->m1.empty()) return FV_OK;
  if (prec == 1) return fv::nufft2d1_fused_impl<float>(plan, prec, bx, by, n_dev, n_cap, scale_host, nb, ntr, W, modes, eps, upsampfac, epi_host);
  return fv::nufft2d1_fused_impl<double>(plan, prec, bx, by, n_dev, n_cap, scale_host, nb, ntr, W, modes, eps, upsampfac, epi_host);
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
  if (prec == 1) return fv::nufft3_impl<float>(plan, prec, dim, x, y, z, n_dev, n_cap, xlim_host, u, v, w, nk, ulim_host, scale_host, nb, ntr, W, eps, upsampfac, epi_host);
  return fv::nufft3_impl<double>(plan, prec, dim, x, y, z, n_dev, n_cap, xlim_host, u, v, w, nk, ulim_host, scale_host, nb, ntr, W, eps, upsampfac, epi_host);
}

extern "C" int fv_minmax(fv_plan* plan, int prec, int dim, const void* x, const void* y, const void* z,
                         const int32_t* n_dev, int64_t n_fixed, double* lim_host) {
  FV_REQUIRE(plan && x && lim_host, "null pointer");
  FV_REQUIRE(dim >= 1 && dim <= 3 && (dim < 2 || y) && (dim < 3 || z), "bad dim / arrays");
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  if (prec == 1) {
    const float* a[3] = {(const float*)x, (const float*)y, (const float*)z};
    return fv::device_limits<float>(plan, a, dim, n_dev, n_fixed, lim_host);
  }
  const double* a[3] = {(const double*)x, (const double*)y, (const double*)z};
  return fv::device_limits<double>(plan, a, dim, n_dev, n_fixed, lim_host);
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
