// fftvis_b200 -- shared device/host helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fftvis_b200.h"

namespace fv {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
extern int64_t g_launches;   // kernels launched by this library (host-side counter)

#define FV_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      fv::set_error(std::string(#call) + " -> " + cudaGetErrorString(e_) + " (" +      \
                    __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
      return (int)e_;                                                                  \
    }                                                                                  \
  } while (0)

#define FV_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      fv::set_error(std::string("invalid argument: ") + (msg)); \
      return FV_ERR_INVALID;                                    \
    }                                                           \
  } while (0)

#define FV_LAUNCH_CHECK()           \
  do {                              \
    ++fv::g_launches;               \
    FV_CUDA(cudaGetLastError());    \
  } while (0)

constexpr int kNumSMs = 148;       // B200
constexpr int kMaxW = 16;          // widest exponential-of-semicircle kernel (finufft clamp)

// ---- complex types ----------------------------------------------------------------------------
template <typename T> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };
template <typename T> using cplx_t = typename cplx_of<T>::type;

template <typename T> __host__ __device__ inline cplx_t<T> make_c(T re, T im) {
  cplx_t<T> r; r.x = re; r.y = im; return r;
}
template <typename C> __host__ __device__ inline C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename C> __host__ __device__ inline C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __host__ __device__ inline C cconj(C a) { a.y = -a.y; return a; }
// conj(a) * b
template <typename C> __host__ __device__ inline C cmulc(C a, C b) {
  C r; r.x = a.x * b.x + a.y * b.y; r.y = a.x * b.y - a.y * b.x; return r;
}

// vector atomic add of one complex value (float2: single RED.ADD.F32x2 on sm_90+)
__device__ inline void atomic_add_c(float2* p, float2 v) { atomicAdd(p, v); }
__device__ inline void atomic_add_c(double2* p, double2 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}

// ---- exponential-of-semicircle kernel ------------------------------------------------------------
// phi(z) = exp(beta (sqrt(1 - (2z/w)^2) - 1)), |z| < w/2, z in grid cells.
template <typename T> __device__ inline T es_kernel(T z, T beta, T c, T halfw);
template <> __device__ inline float es_kernel<float>(float z, float beta, float c, float halfw) {
  float a = 1.0f - c * z * z;
  return (fabsf(z) < halfw && a > 0.f) ? __expf(beta * (sqrtf(a) - 1.0f)) : 0.0f;
}
template <> __device__ inline double es_kernel<double>(double z, double beta, double c, double halfw) {
  double a = 1.0 - c * z * z;
  return (fabs(z) < halfw && a > 0.0) ? exp(beta * (sqrt(a) - 1.0)) : 0.0;
}

// fold an angle (radians, any real) onto the periodic fine grid [0, nf): -pi -> 0, 0 -> nf/2.
// Always evaluated in fp64 (a handful of flops per point) so that the fold never costs accuracy
// on top of the caller's own rounding of the coordinate.
__device__ inline double fold_grid(double x, int nf) {
  double r = x * 0.15915494309189533577 + 0.5;
  r -= floor(r);
  double g = r * (double)nf;
  return g >= (double)nf ? g - (double)nf : g;
}

__device__ __forceinline__ int wrap_idx(int i, int n) {
  // i in [-n, 2n)
  return i < 0 ? i + n : (i >= n ? i - n : i);
}

// ---- mbarrier + 1-D bulk copy (TMA, cp.async.bulk) helpers; addresses are 32-bit shared-window addresses
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace fv
