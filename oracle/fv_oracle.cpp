// TEST INFRASTRUCTURE ONLY -- CPU oracle for fftvis_b200.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library. The product path (fftvis_b200/) never does.
//
// What it restates
//   * The visibility sum the reference's hot loop evaluates,
//       V_k = sum_s W_s exp(+i (u_k x_s + v_k y_s + w_k z_s))
//     (reference call sites: /root/reference/src/fftvis/cpu/nufft.py:48,105,162; sign +1,
//     SURVEY.md Appendix A.1) as a direct fp64 sum  -> fvo_direct_sum.
//   * The spreading / interpolation halves of the NUFFT the reference delegates to the
//     third-party `finufft` package (dependency of /root/reference/pyproject.toml:35, version
//     UNPINNED, source absent from /root/reference). The published algorithm (Barnett, Magland,
//     af Klinteberg, SIAM J. Sci. Comput. 41(5), 2019; arXiv:1808.06736) is restated here:
//     "exponential of semicircle" kernel phi(z)=exp(beta(sqrt(1-(2z/w)^2)-1)), |z|<=w/2,
//     fold of NU points to the periodic fine grid, w^d spreading (type 1 / type 3 step 1) and
//     w^d interpolation (type 2 / type 3 step 2).  The FFT between them and the
//     deconvolution are driven from oracle/nufft_cpu.py (scipy pocketfft).
//   * PARITY UNPINNED for this third-party boundary: the reference holds no golden vectors
//     for finufft and finufft cannot be installed offline, so this file is validated against
//     the direct sum only (tests/test_oracle.py).
//
// Build: oracle/build.py  (g++ -O3 -fopenmp -shared -fPIC) -> oracle/_build/libfv_oracle.so
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <typename T>
inline T es_eval(T z, T beta, T c, T halfw) {
  // phi(z), z in grid units.  finufft's evaluate_kernel (kerevalmeth=0 form).
  if (std::fabs(z) >= halfw) return T(0);
  return std::exp(beta * (std::sqrt(T(1) - c * z * z) - T(1)));
}

template <typename T>
inline T fold_to_grid(T x, int64_t nf) {
  // finufft fold_rescale: x (radians, any real) -> [0, nf), x=-pi -> 0, x=0 -> nf/2.
  const T inv2pi = T(0.159154943091895345554011992339482617);
  T r = x * inv2pi + T(0.5);
  r -= std::floor(r);
  T g = r * T(nf);
  if (g >= T(nf)) g -= T(nf);  // guard r*nf rounding up to nf
  return g;
}

inline int64_t wrap(int64_t i, int64_t n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// Spread ntr strength vectors (each n long) onto ntr periodic grids of nf1*nf2*nf3
// (x fastest).  dim is 2 or 3 (nf3 = 1 for 2-D).  Threads own slabs of y rows:
// each thread walks the points whose footprint touches its rows (points pre-binned by
// y row), so there are no atomics and the result is deterministic.
template <typename T>
void spread_impl(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const T* x,
                 const T* y, const T* z, int ntr, const std::complex<T>* c,
                 std::complex<T>* fw, int ns, double beta_d, int nthreads) {
  const T beta = T(beta_d), cc = T(4.0 / (double(ns) * ns)), halfw = T(ns) / T(2);
  const int64_t ngrid = nf1 * nf2 * nf3;
  std::memset(static_cast<void*>(fw), 0, sizeof(std::complex<T>) * ngrid * ntr);
  if (n == 0) return;
  // bin by first row touched in y
  std::vector<T> gx(n), gy(n), gz(dim == 3 ? n : 0);
  std::vector<int64_t> row0(n);
  std::vector<int64_t> cnt(nf2 + 1, 0);
  for (int64_t j = 0; j < n; ++j) {
    gx[j] = fold_to_grid(x[j], nf1);
    gy[j] = fold_to_grid(y[j], nf2);
    if (dim == 3) gz[j] = fold_to_grid(z[j], nf3);
    int64_t i2 = (int64_t)std::ceil(gy[j] - halfw);
    row0[j] = wrap(i2, nf2);
    cnt[row0[j] + 1]++;
  }
  for (int64_t r = 0; r < nf2; ++r) cnt[r + 1] += cnt[r];
  std::vector<int64_t> order(n), fill(cnt.begin(), cnt.end() - 1);
  for (int64_t j = 0; j < n; ++j) order[fill[row0[j]]++] = j;

  int nt = std::max(1, nthreads);
  if (nf2 < (int64_t)nt * 2 * ns) nt = (int)std::max<int64_t>(1, nf2 / (2 * ns));
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num(), tn = omp_get_num_threads();
#else
    const int tid = 0, tn = 1;
#endif
    const int64_t r_lo = nf2 * tid / tn, r_hi = nf2 * (tid + 1) / tn;  // owned rows
    std::vector<T> k1(ns), k2(ns), k3(ns);
    // points whose first row lies in [r_lo - ns + 1, r_hi) (periodic) can touch owned rows
    // (a single thread owns every row: walk each bin exactly once)
    for (int64_t rr = (tn == 1 ? 0 : r_lo - ns + 1); rr < r_hi; ++rr) {
      const int64_t r = wrap(rr, nf2);
      for (int64_t q = cnt[r]; q < cnt[r + 1]; ++q) {
        const int64_t j = order[q];
        const int64_t i1 = (int64_t)std::ceil(gx[j] - halfw);
        const int64_t i2 = (int64_t)std::ceil(gy[j] - halfw);
        const T x1 = T(i1) - gx[j], x2 = T(i2) - gy[j];
        for (int a = 0; a < ns; ++a) {
          k1[a] = es_eval(x1 + T(a), beta, cc, halfw);
          k2[a] = es_eval(x2 + T(a), beta, cc, halfw);
        }
        int64_t i3 = 0;
        int n3 = 1;
        if (dim == 3) {
          i3 = (int64_t)std::ceil(gz[j] - halfw);
          const T x3 = T(i3) - gz[j];
          for (int a = 0; a < ns; ++a) k3[a] = es_eval(x3 + T(a), beta, cc, halfw);
          n3 = ns;
        }
        for (int t = 0; t < ntr; ++t) {
          const std::complex<T> cj = c[(int64_t)t * n + j];
          std::complex<T>* g = fw + (int64_t)t * ngrid;
          for (int a3 = 0; a3 < n3; ++a3) {
            const int64_t p3 = dim == 3 ? wrap(i3 + a3, nf3) : 0;
            const T w3 = dim == 3 ? k3[a3] : T(1);
            for (int a2 = 0; a2 < ns; ++a2) {
              // rows are walked un-wrapped relative to rr so that ownership is unambiguous
              const int64_t row = wrap(i2 + a2, nf2);
              if (!(row >= r_lo && row < r_hi)) continue;
              const std::complex<T> cw = cj * (k2[a2] * w3);
              std::complex<T>* grow = g + (p3 * nf2 + row) * nf1;
              for (int a1 = 0; a1 < ns; ++a1) {
                grow[wrap(i1 + a1, nf1)] += cw * k1[a1];
              }
            }
          }
        }
      }
    }
  }
}

// Interpolate ntr periodic grids at n NU points (type-2 step).
template <typename T>
void interp_impl(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const T* x,
                 const T* y, const T* z, int ntr, const std::complex<T>* fw,
                 std::complex<T>* c, int ns, double beta_d, int nthreads) {
  const T beta = T(beta_d), cc = T(4.0 / (double(ns) * ns)), halfw = T(ns) / T(2);
  const int64_t ngrid = nf1 * nf2 * nf3;
#pragma omp parallel for schedule(static) num_threads(std::max(1, nthreads))
  for (int64_t j = 0; j < n; ++j) {
    T k1[16], k2[16], k3[16];
    const T g1 = fold_to_grid(x[j], nf1), g2 = fold_to_grid(y[j], nf2);
    const int64_t i1 = (int64_t)std::ceil(g1 - halfw), i2 = (int64_t)std::ceil(g2 - halfw);
    const T x1 = T(i1) - g1, x2 = T(i2) - g2;
    for (int a = 0; a < ns; ++a) {
      k1[a] = es_eval(x1 + T(a), beta, cc, halfw);
      k2[a] = es_eval(x2 + T(a), beta, cc, halfw);
    }
    int64_t i3 = 0;
    int n3 = 1;
    if (dim == 3) {
      const T g3 = fold_to_grid(z[j], nf3);
      i3 = (int64_t)std::ceil(g3 - halfw);
      const T x3 = T(i3) - g3;
      for (int a = 0; a < ns; ++a) k3[a] = es_eval(x3 + T(a), beta, cc, halfw);
      n3 = ns;
    }
    for (int t = 0; t < ntr; ++t) {
      const std::complex<T>* g = fw + (int64_t)t * ngrid;
      std::complex<T> acc(0, 0);
      for (int a3 = 0; a3 < n3; ++a3) {
        const int64_t p3 = dim == 3 ? wrap(i3 + a3, nf3) : 0;
        const T w3 = dim == 3 ? k3[a3] : T(1);
        for (int a2 = 0; a2 < ns; ++a2) {
          const std::complex<T>* grow = g + (p3 * nf2 + wrap(i2 + a2, nf2)) * nf1;
          std::complex<T> racc(0, 0);
          for (int a1 = 0; a1 < ns; ++a1) racc += grow[wrap(i1 + a1, nf1)] * k1[a1];
          acc += racc * (k2[a2] * w3);
        }
      }
      c[(int64_t)t * n + j] = acc;
    }
  }
}

}  // namespace

extern "C" {

// V[t][k] = sum_s W[t][s] exp(i*isign*(u_k x_s + v_k y_s + w_k z_s)); all fp64.
// z / w may be NULL (2-D).  Restates SURVEY.md A.1; ground truth for every parity test.
void fvo_direct_sum(int64_t n, const double* x, const double* y, const double* z, int ntr,
                    const double* w_re_im /* (ntr, n, 2) */, int64_t nk, const double* u,
                    const double* v, const double* w, int isign,
                    double* out_re_im /* (ntr, nk, 2) */, int nthreads) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(std::max(1, nthreads))
  for (int64_t k = 0; k < nk; ++k) {
    const double uk = u[k], vk = v[k], wk = (w && z) ? w[k] : 0.0;
    std::vector<double> accr(ntr, 0.0), acci(ntr, 0.0);
    for (int64_t s = 0; s < n; ++s) {
      double ph = uk * x[s] + vk * y[s];
      if (w && z) ph += wk * z[s];
      ph *= isign;
      const double cs = std::cos(ph), sn = std::sin(ph);
      for (int t = 0; t < ntr; ++t) {
        const double wr = w_re_im[((int64_t)t * n + s) * 2], wi = w_re_im[((int64_t)t * n + s) * 2 + 1];
        accr[t] += wr * cs - wi * sn;
        acci[t] += wr * sn + wi * cs;
      }
    }
    for (int t = 0; t < ntr; ++t) {
      out_re_im[((int64_t)t * nk + k) * 2] = accr[t];
      out_re_im[((int64_t)t * nk + k) * 2 + 1] = acci[t];
    }
  }
}

void fvo_spread_f64(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const double* x,
                    const double* y, const double* z, int ntr, const void* c, void* fw, int ns,
                    double beta, int nthreads) {
  spread_impl<double>(dim, nf1, nf2, nf3, n, x, y, z, ntr,
                      static_cast<const std::complex<double>*>(c),
                      static_cast<std::complex<double>*>(fw), ns, beta, nthreads);
}
void fvo_spread_f32(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const float* x,
                    const float* y, const float* z, int ntr, const void* c, void* fw, int ns,
                    double beta, int nthreads) {
  spread_impl<float>(dim, nf1, nf2, nf3, n, x, y, z, ntr,
                     static_cast<const std::complex<float>*>(c),
                     static_cast<std::complex<float>*>(fw), ns, beta, nthreads);
}
void fvo_interp_f64(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const double* x,
                    const double* y, const double* z, int ntr, const void* fw, void* c, int ns,
                    double beta, int nthreads) {
  interp_impl<double>(dim, nf1, nf2, nf3, n, x, y, z, ntr,
                      static_cast<const std::complex<double>*>(fw),
                      static_cast<std::complex<double>*>(c), ns, beta, nthreads);
}
void fvo_interp_f32(int dim, int64_t nf1, int64_t nf2, int64_t nf3, int64_t n, const float* x,
                    const float* y, const float* z, int ntr, const void* fw, void* c, int ns,
                    double beta, int nthreads) {
  interp_impl<float>(dim, nf1, nf2, nf3, n, x, y, z, ntr,
                     static_cast<const std::complex<float>*>(fw),
                     static_cast<std::complex<float>*>(c), ns, beta, nthreads);
}

int fvo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
