// Stage a1-a3 of the hot path as one fused pass per time step:
//   equatorial -> topocentric rotation, horizon cut, ORDER-PRESERVING stream compaction with warp
//   ballots, az/za of the kept sources (from the un-rotated ENU vector), array-plane rotation and
//   the 2*pi scale.
// Replaces: coord_mgr.rotate / select_chunk (reference cpu_simulate.py:937-946; matvis),
//   enu_to_az_za (:957-959; matvis), inplace_rot (cpu/utils.py:5-24; calls :961-965), topo *= 2pi (:967).
//
// HBM-bound streaming kernel: reads 24 B/source (fp64 unit vector), writes (5 reals + 1 int32) per
// kept source.  Three launches: per-block counts -> one-block exclusive scan -> ordered write.
// The count stays on the device (n_dev) so no host synchronisation is needed per time step.
#include "common.cuh"

namespace fv {

constexpr int RC_THREADS = 256;
constexpr int RC_ITEMS = 4;
constexpr int RC_TILE = RC_THREADS * RC_ITEMS;

// Per-time rotation + the star-independent astrometry parameters (fv_astrom block of the host,
// fftvis_b200/core/astrometry.py): Sun -> observer unit vector and distance (light deflection),
// observer barycentric velocity / c and sqrt(1 - v^2) (annual + diurnal aberration).
struct Mat3 { double m[9]; double eh[3], em, v[3], bm1, dlim; int on; };

constexpr double kSRS = 1.97412574336e-8;   // Schwarzschild radius of the Sun (au)

// ICRS unit vector -> local East-North-Up: light deflection by the Sun and aberration (the published
// SOFA ldsun / ab arithmetic, per source), then the 3x3 of the time step.  One out-of-line copy: the
// count and the write kernels must agree bit for bit on which sources are above the horizon.
__device__ __noinline__ void enu_fp64(const double* __restrict__ eq, int64_t nsrc, int64_t s, const Mat3& M,
                                      double* out) {
  double x = eq[s], y = eq[nsrc + s], z = eq[2 * nsrc + s];
  if (M.on) {
    // deflection: p1 = p + w * (p x (e x p)),  w = SRS / em / max(p . (p + e), dlim)
    const double qx = x + M.eh[0], qy = y + M.eh[1], qz = z + M.eh[2];
    const double w = kSRS / M.em / fmax(x * qx + y * qy + z * qz, M.dlim);
    const double cx = M.eh[1] * z - M.eh[2] * y, cy = M.eh[2] * x - M.eh[0] * z, cz = M.eh[0] * y - M.eh[1] * x;
    const double x1 = x + w * (y * cz - z * cy), y1 = y + w * (z * cx - x * cz), z1 = z + w * (x * cy - y * cx);
    // aberration
    const double pdv = x1 * M.v[0] + y1 * M.v[1] + z1 * M.v[2];
    const double w1 = 1.0 + pdv / (1.0 + M.bm1), w2 = kSRS / M.em;
    const double ax = x1 * M.bm1 + w1 * M.v[0] + w2 * (M.v[0] - pdv * x1);
    const double ay = y1 * M.bm1 + w1 * M.v[1] + w2 * (M.v[1] - pdv * y1);
    const double az = z1 * M.bm1 + w1 * M.v[2] + w2 * (M.v[2] - pdv * z1);
    const double r = 1.0 / sqrt(ax * ax + ay * ay + az * az);
    x = ax * r; y = ay * r; z = az * r;
  }
  out[0] = M.m[0] * x + M.m[1] * y + M.m[2] * z;
  out[1] = M.m[3] * x + M.m[4] * y + M.m[5] * z;
  out[2] = M.m[6] * x + M.m[7] * y + M.m[8] * z;
}

template <typename T>
__device__ inline void enu_of(const double* __restrict__ eq, int64_t nsrc, int64_t s, const Mat3& M,
                              T& e, T& n, T& u) {
  double o[3];
  enu_fp64(eq, nsrc, s, M, o);
  e = (T)o[0]; n = (T)o[1]; u = (T)o[2];
}

template <typename T>
__global__ void __launch_bounds__(RC_THREADS)
rc_count_kernel(const double* __restrict__ eq, int64_t nsrc, int64_t lo, int64_t hi, Mat3 M,
                int32_t* __restrict__ block_counts) {
  __shared__ int warp_cnt[RC_THREADS / 32];
  const int64_t base = lo + (int64_t)blockIdx.x * RC_TILE;
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < RC_ITEMS; ++it) {
    const int64_t s = base + it * RC_THREADS + threadIdx.x;
    bool up = false;
    if (s < hi) {
      T e, n, u;
      enu_of<T>(eq, nsrc, s, M, e, n, u);
      up = u > T(0);
    }
    cnt += __popc(__ballot_sync(0xffffffffu, up));
  }
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < RC_THREADS / 32; ++w) tot += warp_cnt[w];
    block_counts[blockIdx.x] = tot;
  }
}

// single-block exclusive scan over the per-block counts (<= a few thousand entries)
__global__ void __launch_bounds__(1024)
rc_scan_kernel(int32_t* __restrict__ block_counts, int nblocks, int64_t n_cap,
               int32_t* __restrict__ n_dev) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int i = start + threadIdx.x;
    const int v = i < nblocks ? block_counts[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
      int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    if (i < nblocks) block_counts[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_dev = (int64_t)carry > n_cap ? -carry : carry;
}

template <typename T> __device__ inline T t_atan2(T a, T b);
template <> __device__ inline float t_atan2(float a, float b) { return atan2f(a, b); }
template <> __device__ inline double t_atan2(double a, double b) { return atan2(a, b); }
template <typename T> __device__ inline T t_asin(T a);
template <> __device__ inline float t_asin(float a) { return asinf(a); }
template <> __device__ inline double t_asin(double a) { return asin(a); }
template <typename T> __device__ inline T t_sqrt(T a);
template <> __device__ inline float t_sqrt(float a) { return sqrtf(a); }
template <> __device__ inline double t_sqrt(double a) { return sqrt(a); }
__device__ inline float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ inline double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ inline float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ inline double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T> struct Mat3T { T m[9]; };

template <typename T>
__global__ void __launch_bounds__(RC_THREADS)
rc_write_kernel(const double* __restrict__ eq, int64_t nsrc, int64_t lo, int64_t hi, Mat3 M,
                Mat3T<T> P, const int32_t* __restrict__ block_offsets,
                const int32_t* __restrict__ n_dev, int64_t n_cap, T* __restrict__ xyz,
                T* __restrict__ az, T* __restrict__ za, int32_t* __restrict__ src_idx) {
  if (*n_dev < 0) return;   // overflow: outputs undefined, caller reports the error
  __shared__ int warp_cnt[RC_ITEMS][RC_THREADS / 32];
  const int64_t base = lo + (int64_t)blockIdx.x * RC_TILE;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T e[RC_ITEMS], n[RC_ITEMS], u[RC_ITEMS];
  unsigned ball[RC_ITEMS];
#pragma unroll
  for (int it = 0; it < RC_ITEMS; ++it) {
    const int64_t s = base + it * RC_THREADS + threadIdx.x;
    bool up = false;
    if (s < hi) {
      enu_of<T>(eq, nsrc, s, M, e[it], n[it], u[it]);
      up = u[it] > T(0);
    }
    ball[it] = __ballot_sync(0xffffffffu, up);
    if (lane == 0) warp_cnt[it][wid] = __popc(ball[it]);
  }
  __syncthreads();
  // order: item-major then thread (matches s ascending inside the tile)
  int running = block_offsets[blockIdx.x];
#pragma unroll
  for (int it = 0; it < RC_ITEMS; ++it) {
    int before = 0;
    for (int w = 0; w < RC_THREADS / 32; ++w) {
      const int c = warp_cnt[it][w];
      before += (w < wid) ? c : 0;
    }
    int tot = 0;
    for (int w = 0; w < RC_THREADS / 32; ++w) tot += warp_cnt[it][w];
    if ((ball[it] >> lane) & 1u) {
      const int64_t o = running + before + __popc(ball[it] & ((1u << lane) - 1u));
      const int64_t s = base + it * RC_THREADS + threadIdx.x;
      const T ee = e[it], nn = n[it], uu = u[it];
      // az/za from the un-rotated ENU direction cosines, "uvbeam" convention
      T r2 = T(1) - ee * ee - nn * nn;
      T zeta = r2 > T(0) ? t_sqrt(r2) : T(0);
      T zz = T(1.5707963267948966) - t_asin(zeta);
      T aa = T(1.5707963267948966) - t_atan2(ee, nn);
      const T twopi = T(6.283185307179586);
      if (aa < T(0)) aa += twopi;
      if (aa >= twopi) aa -= twopi;
      az[o] = aa;
      za[o] = zz;
      // array-plane rotation in working precision, products and sums rounded separately in the
      // reference's order (cpu/utils.py:19-22), then * 2 pi (cpu_simulate.py:967)
      const T x0 = add_rn(add_rn(mul_rn(P.m[0], ee), mul_rn(P.m[1], nn)), mul_rn(P.m[2], uu));
      const T x1 = add_rn(add_rn(mul_rn(P.m[3], ee), mul_rn(P.m[4], nn)), mul_rn(P.m[5], uu));
      const T x2 = add_rn(add_rn(mul_rn(P.m[6], ee), mul_rn(P.m[7], nn)), mul_rn(P.m[8], uu));
      xyz[o] = mul_rn(x0, twopi);
      xyz[n_cap + o] = mul_rn(x1, twopi);
      xyz[2 * n_cap + o] = mul_rn(x2, twopi);
      src_idx[o] = (int32_t)s;
    }
    running += tot;
  }
}

template <typename T>
__global__ void inplace_rot_kernel(Mat3T<T> R, T* __restrict__ b, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const T b0 = b[s], b1 = b[n + s], b2 = b[2 * n + s];
  b[s] = add_rn(add_rn(mul_rn(R.m[0], b0), mul_rn(R.m[1], b1)), mul_rn(R.m[2], b2));
  b[n + s] = add_rn(add_rn(mul_rn(R.m[3], b0), mul_rn(R.m[4], b1)), mul_rn(R.m[5], b2));
  b[2 * n + s] = add_rn(add_rn(mul_rn(R.m[6], b0), mul_rn(R.m[7], b1)), mul_rn(R.m[8], b2));
}

template <typename T>
static int rotate_cut_impl(const double* eq, int64_t nsrc, int64_t lo, int64_t hi, const double* enu,
                           const double* astrom, const double* plane, void* xyz, void* az, void* za, int32_t* src_idx,
                           int64_t n_cap, int32_t* n_dev, void* scratch, cudaStream_t st) {
  Mat3 M;
  Mat3T<T> P;
  for (int i = 0; i < 9; ++i) { M.m[i] = enu[i]; P.m[i] = (T)plane[i]; }
  M.on = astrom != nullptr && astrom[9] != 0.0;
  for (int i = 0; i < 3; ++i) { M.eh[i] = M.on ? astrom[i] : 0.0; M.v[i] = M.on ? astrom[4 + i] : 0.0; }
  M.em = M.on ? astrom[3] : 1.0; M.bm1 = M.on ? astrom[7] : 1.0; M.dlim = M.on ? astrom[8] : 1e-6;
  const int64_t cnt = hi - lo;
  const int nblocks = (int)((cnt + RC_TILE - 1) / RC_TILE);
  int32_t* counts = (int32_t*)scratch;
  if (nblocks == 0) {
    FV_CUDA(cudaMemsetAsync(n_dev, 0, sizeof(int32_t), st));
    return FV_OK;
  }
  rc_count_kernel<T><<<nblocks, RC_THREADS, 0, st>>>(eq, nsrc, lo, hi, M, counts);
  FV_LAUNCH_CHECK();
  rc_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, n_cap, n_dev);
  FV_LAUNCH_CHECK();
  rc_write_kernel<T><<<nblocks, RC_THREADS, 0, st>>>(eq, nsrc, lo, hi, M, P, counts, n_dev, n_cap,
                                                     (T*)xyz, (T*)az, (T*)za, src_idx);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

extern "C" int64_t fv_rotate_cut_scratch_bytes(int64_t nsrc) {
  return ((nsrc + fv::RC_TILE - 1) / fv::RC_TILE + 1) * (int64_t)sizeof(int32_t);
}

extern "C" int fv_rotate_cut(int prec, const double* eq_xyz, int64_t nsrc, int64_t src_lo,
                             int64_t src_hi, const double* enu_mat_host, const double* astrom_host,
                             const double* plane_mat_host, void* xyz, void* az, void* za,
                             int32_t* src_idx, int64_t n_cap, int32_t* n_dev, void* scratch,
                             void* stream) {
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(src_lo >= 0 && src_hi <= nsrc && src_lo <= src_hi, "bad source slice");
  FV_REQUIRE(nsrc < (int64_t)INT32_MAX, "catalogue too large for int32 source indices");
  FV_REQUIRE(eq_xyz && enu_mat_host && plane_mat_host && xyz && az && za && src_idx && n_dev && scratch,
             "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (prec == 1)
    return fv::rotate_cut_impl<float>(eq_xyz, nsrc, src_lo, src_hi, enu_mat_host, astrom_host, plane_mat_host, xyz,
                                      az, za, src_idx, n_cap, n_dev, scratch, st);
  return fv::rotate_cut_impl<double>(eq_xyz, nsrc, src_lo, src_hi, enu_mat_host, astrom_host, plane_mat_host, xyz,
                                     az, za, src_idx, n_cap, n_dev, scratch, st);
}

extern "C" int fv_inplace_rot(int prec, const double* rot_host, void* b, int64_t n, void* stream) {
  FV_REQUIRE(prec == 1 || prec == 2, "prec must be 1 or 2");
  FV_REQUIRE(rot_host && (b || n == 0), "null pointer");
  if (n == 0) return FV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = fv::ceil_div(n, 256);
  if (prec == 1) {
    fv::Mat3T<float> R;
    for (int i = 0; i < 9; ++i) R.m[i] = (float)rot_host[i];
    fv::inplace_rot_kernel<float><<<blocks, 256, 0, st>>>(R, (float*)b, n);
  } else {
    fv::Mat3T<double> R;
    for (int i = 0; i < 9; ++i) R.m[i] = rot_host[i];
    fv::inplace_rot_kernel<double><<<blocks, 256, 0, st>>>(R, (double*)b, n);
  }
  FV_LAUNCH_CHECK();
  return FV_OK;
}
