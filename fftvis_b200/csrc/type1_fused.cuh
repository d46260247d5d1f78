// Type-1 transform for gridded arrays, B200 form: the fine grid never exists in global memory.
//
//   pass 1  t1_spread_fftx_kernel   one CTA per (strip of R grid rows, frequency, transform):
//           scan the live sources, warp-ballot-compact those whose w-row footprint touches the
//           strip, spread them into the shared-memory strip (rows owned by warps: no atomics), FFT every
//           row along x in shared memory, and write ONLY the columns some baseline needs
//           (n_cols <= n_modes of the nf columns) to the half-transformed array T[col][row].
//   pass 2  t1_ffty_gather_kernel   one CTA per (group of needed columns, frequency, transform):
//           load the columns (contiguous rows), FFT along y in shared memory, then for every
//           baseline whose first mode number is this column: deconvolve, conjugate if flipped,
//           and store / accumulate straight into the visibility array (the epilogue).
//
// Replaces the spread -> cuFFT -> gather chain of fv_nufft2d1 (finufft.nufft2d1 + mode gather,
// reference cpu/nufft.py:120-175) for the same kernel, grid size and deconvolution, so the error
// model is unchanged.  Global traffic per transform drops from ~5 passes over nf^2 cells (memset,
// atomics, two cuFFT passes) to one write + one read of nf x n_cols cells.
//
// The shared-memory FFT is an in-place decimation-in-frequency mixed-radix (4, 2, 5, 3) transform:
// outputs land in digit-reversed positions, which costs nothing here because both passes read
// their outputs through a position table.
#pragma once
#include <limits.h>
#include <cuda_pipeline.h>

namespace fv {

constexpr int T1_THREADS = 256;
constexpr int T1_MAX_STAGES = 16;

struct FftStages {
  int nstage;
  int radix[T1_MAX_STAGES];
  unsigned inv_m[T1_MAX_STAGES];   // floor(2^32 / m) + 1 for the stage's sub-length m (exact division for t < 2^16)
  int tw_off[T1_MAX_STAGES];       // start of the stage's twiddles in the table: entry (q - 1) * m + j holds
                                   // exp(+2 pi i j q / n): consecutive lanes (j) read consecutive words
  int tw_len;                      // total table length (<= N)
  int hw[T1_MAX_STAGES];           // butterflies per half-warp of the wide-radix stages: 16, or m when 9 <= m <= 15
                                   // is odd (a half-warp then stays inside one block: no bank conflicts)
};

template <typename C> __device__ __forceinline__ C c_add(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C c_sub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C c_muli(C a) { C r; r.x = -a.y; r.y = a.x; return r; }   // i * a

// Shared-memory accesses by 32-bit shared address, as volatile PTX: the compiler keeps their relative
// order (loads of the next hit's record stay ahead of the current hit's stores) and emits a single
// LDS / STS each instead of re-deriving the shared window per access.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_i32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u8(unsigned a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_real(unsigned a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_real(unsigned a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_cplx(unsigned a, float) {
  float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ double2 lds_cplx(unsigned a, double) {
  double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_cplx(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_cplx(unsigned a, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v.x), "d"(v.y) : "memory");
}

// r-point DFTs on registers, sign +1:  y_k = sum_q x_q exp(+2 pi i q k / r)
template <typename T> __device__ __forceinline__ void dft2(cplx_t<T>* x) {
  const cplx_t<T> t = c_sub(x[0], x[1]);
  x[0] = c_add(x[0], x[1]); x[1] = t;
}
template <typename T> __device__ __forceinline__ void dft4(cplx_t<T>* x) {
  const cplx_t<T> t0 = c_add(x[0], x[2]), t1 = c_sub(x[0], x[2]);
  const cplx_t<T> t2 = c_add(x[1], x[3]), t3 = c_muli(c_sub(x[1], x[3]));
  x[0] = c_add(t0, t2); x[2] = c_sub(t0, t2); x[1] = c_add(t1, t3); x[3] = c_sub(t1, t3);
}
template <typename T> __device__ __forceinline__ void dft3(cplx_t<T>* x) {
  const T h = T(0.86602540378443864676);
  const cplx_t<T> s = c_add(x[1], x[2]), d = c_sub(x[1], x[2]);
  cplx_t<T> m; m.x = x[0].x - T(0.5) * s.x; m.y = x[0].y - T(0.5) * s.y;
  cplx_t<T> id; id.x = -h * d.y; id.y = h * d.x;
  x[0] = c_add(x[0], s); x[1] = c_add(m, id); x[2] = c_sub(m, id);
}
template <typename T> __device__ __forceinline__ void dft5(cplx_t<T>* x) {
  const T c1 = T(0.30901699437494742410), c2 = T(-0.80901699437494742410);
  const T s1 = T(0.95105651629515357212), s2 = T(0.58778525229247312917);
  const cplx_t<T> a1 = c_add(x[1], x[4]), a2 = c_add(x[2], x[3]);
  const cplx_t<T> b1 = c_sub(x[1], x[4]), b2 = c_sub(x[2], x[3]);
  cplx_t<T> r1, r2, i1, i2;
  r1.x = x[0].x + c1 * a1.x + c2 * a2.x; r1.y = x[0].y + c1 * a1.y + c2 * a2.y;
  r2.x = x[0].x + c2 * a1.x + c1 * a2.x; r2.y = x[0].y + c2 * a1.y + c1 * a2.y;
  i1.x = -(s1 * b1.y + s2 * b2.y); i1.y = s1 * b1.x + s2 * b2.x;      // i * (s1 b1 + s2 b2)
  i2.x = -(s2 * b1.y - s1 * b2.y); i2.y = s2 * b1.x - s1 * b2.x;      // i * (s2 b1 - s1 b2)
  x[0].x += a1.x + a2.x; x[0].y += a1.y + a2.y;
  x[1] = c_add(r1, i1); x[4] = c_sub(r1, i1); x[2] = c_add(r2, i2); x[3] = c_sub(r2, i2);
}

// One DIF stage of radix R over the vectors owned by this warp (loop specialised per radix; indices
// are plain ints relative to the vector base so the compiler emits immediate-offset LDS/STS).
template <typename T> __device__ __forceinline__ void dft8(cplx_t<T>* x) {
  using C = cplx_t<T>;
  const T h = T(0.70710678118654752440);
  // three radix-2 levels, decimation in frequency, then the bit-reversed outputs are put in order
  C a[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) { a[q] = c_add(x[q], x[q + 4]); a[q + 4] = c_sub(x[q], x[q + 4]); }
  // twiddles w8^q on the lower half, w8 = exp(+i pi / 4)
  { C t = a[5]; a[5].x = h * (t.x - t.y); a[5].y = h * (t.x + t.y); }
  a[6] = c_muli(a[6]);
  { C t = a[7]; a[7].x = -h * (t.x + t.y); a[7].y = h * (t.x - t.y); }
  C b[8];
#pragma unroll
  for (int g = 0; g < 8; g += 4) {
    b[g] = c_add(a[g], a[g + 2]); b[g + 2] = c_sub(a[g], a[g + 2]);
    b[g + 1] = c_add(a[g + 1], a[g + 3]); b[g + 3] = c_muli(c_sub(a[g + 1], a[g + 3]));
  }
  // last level: pairs (0,1), (2,3), (4,5), (6,7) -> outputs k = 0,4 | 2,6 | 1,5 | 3,7
  x[0] = c_add(b[0], b[1]); x[4] = c_sub(b[0], b[1]);
  x[2] = c_add(b[2], b[3]); x[6] = c_sub(b[2], b[3]);
  x[1] = c_add(b[4], b[5]); x[5] = c_sub(b[4], b[5]);
  x[3] = c_add(b[6], b[7]); x[7] = c_sub(b[6], b[7]);
}

// 15-point DFT by the Good-Thomas prime-factor map (3 x 5, no twiddles): input n = (5 n1 + 3 n2)
// mod 15, output k = (10 k1 + 6 k2) mod 15; all index shuffling is register renaming.
template <typename T> __device__ __forceinline__ void dft15(cplx_t<T>* x) {
  using C = cplx_t<T>;
  C a[3][5];
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) {
    C c[3] = {x[(3 * n2) % 15], x[(5 + 3 * n2) % 15], x[(10 + 3 * n2) % 15]};
    dft3<T>(c);
    a[0][n2] = c[0]; a[1][n2] = c[1]; a[2][n2] = c[2];
  }
#pragma unroll
  for (int k1 = 0; k1 < 3; ++k1) {
    dft5<T>(a[k1]);
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) x[(10 * k1 + 6 * k2) % 15] = a[k1][k2];
  }
}

template <typename T, int R>
__device__ __forceinline__ void dft_r(cplx_t<T>* x) {
  if (R == 15) dft15<T>(x);
  if (R == 8) dft8<T>(x);
  if (R == 2) dft2<T>(x);
  if (R == 3) dft3<T>(x);
  if (R == 4) dft4<T>(x);
  if (R == 5) dft5<T>(x);
}

// One DIF stage of radix R over the vectors owned by this warp.  Two butterflies per iteration with
// every load issued before the first store (the butterflies are disjoint, which the compiler cannot
// prove for shared memory): the second butterfly's loads overlap the first one's arithmetic.
template <typename T, int R>
__device__ __forceinline__ void fft_stage(cplx_t<T>* data, int nvec, int pitch, int N, int n, unsigned inv, int hw,
                                          const cplx_t<T>* __restrict__ tws, int lane, int warp, int nwarps) {
  using C = cplx_t<T>;
  const int m = n / R, per_vec = N / R;
  for (int v = warp; v < nvec; v += nwarps) {
    C* vec = data + v * pitch;
    if (R >= 8) {
      // wide butterfly: one per iteration (register budget), loads before stores.  Lanes take
      // butterflies half-warp by half-warp, `hw` each (16, or the sub-length m when that keeps the 16
      // lanes of a shared-memory wavefront inside one block of the stage).
      for (int t0 = (lane >> 4) * hw; t0 < per_vec; t0 += 2 * hw) {
        const int t = t0 + (lane & 15);
        if ((lane & 15) >= hw || t >= per_vec) continue;
        const int blk = inv ? (int)__umulhi((unsigned)t, inv) : t;
        const int j = t - blk * m;
        const int base = blk * n + j;
        C x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = vec[base + q * m];
        dft_r<T, R>(x);
        vec[base] = x[0];
        if (m > 1) {
#pragma unroll
          for (int q = 1; q < R; ++q) vec[base + q * m] = cmul(x[q], tws[(q - 1) * m + j]);
        } else {
#pragma unroll
          for (int q = 1; q < R; ++q) vec[base + q] = x[q];
        }
      }
      continue;
    }
    for (int t0 = lane; t0 < per_vec; t0 += 64) {
      const int t1 = t0 + 32;
      const bool two = t1 < per_vec;
      const int blk0 = inv ? (int)__umulhi((unsigned)t0, inv) : t0;
      const int blk1 = inv ? (int)__umulhi((unsigned)t1, inv) : t1;
      const int j0 = t0 - blk0 * m, j1 = t1 - blk1 * m;
      const int base0 = blk0 * n + j0, base1 = two ? blk1 * n + j1 : base0;
      C x[R], y[R], wx[R], wy[R];
#pragma unroll
      for (int q = 0; q < R; ++q) x[q] = vec[base0 + q * m];
#pragma unroll
      for (int q = 0; q < R; ++q) y[q] = vec[base1 + q * m];
      if (m > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) { wx[q] = tws[(q - 1) * m + j0]; wy[q] = tws[(q - 1) * m + (two ? j1 : j0)]; }
      }
      dft_r<T, R>(x);
      dft_r<T, R>(y);
      if (m > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) { x[q] = cmul(x[q], wx[q]); y[q] = cmul(y[q], wy[q]); }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) vec[base0 + q * m] = x[q];
      if (two) {
#pragma unroll
        for (int q = 0; q < R; ++q) vec[base1 + q * m] = y[q];
      }
    }
  }
}

// In-place FFT of `nvec` vectors of length N (element stride 1, vector stride `pitch`) held in shared
// memory; every warp owns whole vectors.  Output X[k] sits at the digit-reversed position pos[k].
template <typename T>
__device__ void smem_fft(cplx_t<T>* data, int nvec, int pitch, int N, const cplx_t<T>* __restrict__ tw,
                         const FftStages& st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int n = N;
  for (int s = 0; s < st.nstage; ++s) {
    const int r = st.radix[s];
    const unsigned inv = st.inv_m[s];            // 0 when m == 1
    const int hw = st.hw[s];
    const cplx_t<T>* tws = tw + st.tw_off[s];
    switch (r) {                                 // the next sub-length by a constant division in every case
      case 15: fft_stage<T, 15>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 15; break;
      case 8: fft_stage<T, 8>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 8; break;
      case 4: fft_stage<T, 4>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 4; break;
      case 2: fft_stage<T, 2>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 2; break;
      case 5: fft_stage<T, 5>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 5; break;
      default: fft_stage<T, 3>(data, nvec, pitch, N, n, inv, hw, tws, lane, warp, nwarps); n /= 3; break;
    }
    __syncwarp();                                // a vector's stages only depend on that vector (one warp)
  }
}

template <typename T>
struct T1SpreadArgs {
  const int32_t* n_dev;
  int64_t n_cap;
  int nf, R, pitch, w;
  int nseg, seg;                 // column segments of the segment spreader (host: min(warps, T1_MAXSEG, nf / 4w))
  T beta, c, halfw;
  int ntr;
  unsigned inv_ntr;              // floor(2^32 / ntr) + 1 (0 when ntr == 1)
  const cplx_t<T>* W;            // (nb, ntr, n_cap)
  // per (frequency, source) fold results from t1_prep_kernel, each (nb, n_cap)
  const int32_t* ix0; const int32_t* iy0;   // first grid column / row of the footprint (may be < 0)
  const T* zx; const T* zy;                 // kernel argument of that first cell
  const uint32_t* hm0; const uint32_t* hm1; // strip masks of every (frequency, source) (null: more than 64 strips)
  const cplx_t<T>* tw;           // per-stage twiddle tables (FftStages::tw_off), <= nf entries
  FftStages st;
  int ncols;
  const int32_t* col_pos;        // smem position of each needed column's FFT output (ncols)
  cplx_t<T>* Tbuf;               // (nb, ntr, ncols, nf)
  long long* dbg;                // optional per-phase cycle counters (development aid), 8 per CTA
};

// Fold every (frequency, source) NU point onto the fine grid ONCE per batch (fp64 fold of the
// working-precision product fl(topo * freq), reference cpu_simulate.py:990-992), so that the strip
// CTAs of pass 1 only compare integers when they look for their sources.
template <typename T>
__global__ void __launch_bounds__(256)
t1_prep_kernel(const T* __restrict__ bx, const T* __restrict__ by, const int32_t* __restrict__ n_dev,
               int64_t n_cap, const BatchParams* __restrict__ bp, int nf, int w, int32_t* __restrict__ ix0,
               int32_t* __restrict__ iy0, T* __restrict__ zx, T* __restrict__ zy, int R, int nstrips,
               uint32_t* __restrict__ hm0, uint32_t* __restrict__ hm1) {
  const int n = *n_dev;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int b = blockIdx.y;
  const T smul = (T)bp[b].smul;
  const int64_t o = (int64_t)b * n_cap + s;
  const double hw = 0.5 * (double)w;
  const double gx = fold_grid((double)(bx[s] * smul), nf), gix = ceil(gx - hw);
  const double gy = fold_grid((double)(by[s] * smul), nf), giy = ceil(gy - hw);
  ix0[o] = (int)gix; zx[o] = (T)(gix - gx);
  iy0[o] = (int)giy; zy[o] = (T)(giy - gy);
  if (hm0) {
    // which strips (of R rows; at most 64) the w footprint rows touch, as two 32-bit masks: the strip
    // CTAs of pass 1 then test one bit per source instead of redoing the row arithmetic
    const int y = (int)giy < 0 ? (int)giy + nf : (int)giy;
    const int first = y / R, last_row = y + w - 1;
    unsigned long long m = 0ull;
    if (last_row < nf) {
      for (int st = first, e = last_row / R; st <= e; ++st) m |= 1ull << st;
    } else {
      for (int st = first; st < nstrips; ++st) m |= 1ull << st;
      for (int st = 0, e = (last_row - nf) / R; st <= e; ++st) m |= 1ull << st;
    }
    hm0[o] = (uint32_t)m; hm1[o] = (uint32_t)(m >> 32);
  }
}

constexpr int T1_RC = 192;        // hit records evaluated and spread per flush chunk
constexpr int T1_MAXSEG = 24;     // column segments (= warps) of the thread-per-row spreader
constexpr int T1_SCAN = 8;        // sources per thread per scan iteration (all loads in flight together)
constexpr int T1_SPT = 2;         // sources scanned per thread per tile (hit list holds one tile's worst case)

// shared-memory bytes of pass 1 besides the strip itself
template <typename T>
inline size_t t1_spread_fixed_smem(int nf, int wmax, int threads, int np = 1) {
  return sizeof(cplx_t<T>) * nf                                    // twiddles
         + (size_t)2 * T1_SPT * threads * sizeof(unsigned short)    // hit list (source index relative to `sbase`)
         + sizeof(int) * nf                                         // needed-column positions
         + (size_t)2 * T1_MAXSEG * (sizeof(int) + T1_RC)            // per-segment hit counters + lists
         + (size_t)T1_RC * (np * sizeof(cplx_t<T>) + 2 * wmax * sizeof(T) + 2 * sizeof(int));
}

template <typename T> struct t1_limits;
template <> struct t1_limits<float> { static constexpr int spread_threads = 768, gather_blocks = 3; };
template <> struct t1_limits<double> { static constexpr int spread_threads = 512, gather_blocks = 1; };

// NP = products spread by one CTA: 1 (grid.y = frequencies x products) or 4 (grid.y = frequencies; the
// four polarisation products share the source scan, the kernel evaluations and all index arithmetic,
// and each keeps its own strip in shared memory).
template <typename T, int WT, int NP>
__global__ void __launch_bounds__(t1_limits<T>::spread_threads)
t1_spread_fftx_kernel(T1SpreadArgs<T> a) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t1_smem[];
  const int w = WT > 0 ? WT : a.w;
  constexpr int WMAX = WT > 0 ? WT : kMaxW;
  const int nthr = blockDim.x, lcap = 2 * T1_SPT * nthr;
  C* strip = (C*)t1_smem;                                  // NP * R * pitch (product-major)
  const int pstride = a.R * a.pitch;                       // elements between the strips of two products
  C* tw = strip + (size_t)NP * pstride;                    // nf
  C* rec_w = tw + a.nf;                                    // T1_RC * NP
  T* rec_kx = (T*)(rec_w + T1_RC * NP);                    // T1_RC * WMAX
  T* rec_ky = rec_kx + T1_RC * WMAX;                       // T1_RC * WMAX
  int* rec_i0x = (int*)(rec_ky + T1_RC * WMAX);            // T1_RC
  int* rec_d = rec_i0x + T1_RC;                            // T1_RC
  int* colp = rec_d + T1_RC;                               // ncols (<= nf)
  int* seg_cnt = colp + a.nf;                              // 2 * T1_MAXSEG: [2 * segment + straddles-its-edge]
  unsigned char* seg_list = (unsigned char*)(seg_cnt + 2 * T1_MAXSEG);   // 2 * T1_MAXSEG * T1_RC
  unsigned short* lst_s = (unsigned short*)(seg_list + 2 * T1_MAXSEG * T1_RC);   // lcap, tile-relative
  __shared__ int wcnt[32];

  const int nf = a.nf, pitch = a.pitch;
  const int bpi = NP == 1 ? blockIdx.y : blockIdx.y * a.ntr;   // index of the (first) product's transform
  const int b = NP == 1 ? (a.inv_ntr ? (int)__umulhi(blockIdx.y, a.inv_ntr) : (int)blockIdx.y) : (int)blockIdx.y;
  const int r0 = blockIdx.x * a.R;
  const int rows = min(a.R, nf - r0);
  const int n = *a.n_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int G = 32 / w;                                    // footprint rows per warp instruction (multi-row path)
  // column segments of the thread-per-row path: at least 4 w columns each, one warp per segment
  const int nseg = a.nseg, seg = a.seg;
  const bool use_seg = nseg >= min(nwarps, 8);             // narrow grids: row-block ownership instead
  const int jj = lane / w, jx = lane - jj * w;              // multi-row path: (row in group, column)
  const int jrow = lane / G, tcol = lane - jrow * G;       // segment path: (footprint row, column group)
  const int32_t* iy0 = a.iy0 + (int64_t)b * a.n_cap;
  const int32_t* ix0 = a.ix0 + (int64_t)b * a.n_cap;
  const T* zxp = a.zx + (int64_t)b * a.n_cap;
  const T* zyp = a.zy + (int64_t)b * a.n_cap;
  const C* Wp = a.W + (int64_t)bpi * a.n_cap;

  // optional per-phase cycle counters (development aid): lane 0 of every warp of the first strips adds
  // its phase time straight to global memory, so that only the last time stamp stays live
  long long tc = a.dbg ? clock64() : 0;
  const bool dbg_me = a.dbg && lane == 0 && blockIdx.y == 0 && blockIdx.x < 8;
#define T1_PHASE(i) do { if (a.dbg) { const long long t_ = clock64(); \
    if (dbg_me) atomicAdd((unsigned long long*)&a.dbg[i], (unsigned long long)(t_ - tc)); tc = t_; } } while (0)
  // the prologue's global loads (twiddles, needed-column positions) are issued before the strip is
  // cleared, so that their latency hides under the clear
  C tw_pre[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) { const int i = tid + u * nthr; if (i < a.st.tw_len) tw_pre[u] = a.tw[i]; }
  const int cp_pre = tid < a.ncols ? a.col_pos[tid] : 0;
  {
    // clear the strip with 16-byte stores (32-bit shared addresses: one STS.128 + one add per store)
    const int n16 = (int)(((size_t)NP * pstride * sizeof(C)) / 16);
    const unsigned a0 = smem_addr(strip);
    const unsigned step = 16u * (unsigned)nthr;
    unsigned ap = a0 + 16u * (unsigned)tid;
    const unsigned aend = a0 + 16u * (unsigned)n16;
#pragma unroll 4
    for (; ap < aend; ap += step) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" :: "r"(ap), "r"(0) : "memory");
    for (int i = n16 * (16 / (int)sizeof(C)) + tid; i < NP * pstride; i += nthr) strip[i] = make_c<T>(T(0), T(0));
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) { const int i = tid + u * nthr; if (i < a.st.tw_len) tw[i] = tw_pre[u]; }
  for (int i = tid + 2 * nthr; i < a.st.tw_len; i += nthr) tw[i] = a.tw[i];
  if (tid < a.ncols) colp[tid] = cp_pre;
  for (int i = tid + nthr; i < a.ncols; i += nthr) colp[i] = a.col_pos[i];
  __syncthreads();
  T1_PHASE(0);

  // ---- which sources' w-row footprints touch this strip (integer compares on the pre-folded rows).
  // Two passes over a range of sources: count per warp, then store at deterministic offsets (warp-major
  // order); one barrier each instead of two per tile.  The range is everything left (up to the 16-bit
  // offset limit) when its hits fit the list, else the worst-case-safe `lcap` sources.
  const bool masked = a.hm0 != nullptr;
  const int32_t* scan_src = masked ? reinterpret_cast<const int32_t*>((blockIdx.x < 32 ? a.hm0 : a.hm1) + (int64_t)b * a.n_cap) : iy0;
  const int scan_bit = 1 << (blockIdx.x & 31);
  const int scan_none = masked ? 0 : INT_MIN;              // a word that is never a hit
  auto is_hit = [&](int v) {                                // v: strip mask word, or the first footprint row
    if (masked) return (v & scan_bit) != 0;
    int d = v - r0;
    if (d < 0) d += nf;
    if (d < 0) d += nf;
    return d < rows || d + w > nf;
  };
  int sbase = 0;                                           // the hit list stores 16-bit offsets from here
  while (sbase < n) {
    int shi = min(n, sbase + 65536);
    int nh = 0, wbase = 0;
    // the common case: the whole range is one scan iteration, whose rows stay in registers between
    // the count pass and the store pass (no second trip to global memory)
    const bool single = shi - sbase <= T1_SCAN * nthr;
    unsigned hits_u = 0;                                   // single: bit u = my source u is a hit
    for (int attempt = 0; attempt < 2; ++attempt) {
      int cnt = 0;
      for (int k = sbase + tid; k < shi + lane; k += T1_SCAN * nthr) {     // warp-uniform trip count
        int yv[T1_SCAN];
#pragma unroll
        for (int u = 0; u < T1_SCAN; ++u) { const int s = k + u * nthr; yv[u] = s < shi ? scan_src[s] : scan_none; }
        hits_u = 0;
#pragma unroll
        for (int u = 0; u < T1_SCAN; ++u) {
          const bool hit = yv[u] != scan_none && is_hit(yv[u]);
          hits_u |= hit ? (1u << u) : 0u;
          cnt += __popc(__ballot_sync(0xffffffffu, hit));
        }
      }
      if (lane == 0) wcnt[warp] = cnt;
      __syncthreads();
      {
        // totals and this warp's base by a shuffle scan over the per-warp counts
        int c = lane < nwarps ? wcnt[lane] : 0, inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        nh = __shfl_sync(0xffffffffu, inc, 31);
        wbase = __shfl_sync(0xffffffffu, inc - c, warp);
      }
      __syncthreads();                                     // wcnt may be rewritten by the next attempt
      if (nh <= lcap) break;
      shi = min(n, sbase + lcap);                          // dense strip: take a worst-case-safe range
    }
    if (single && shi - sbase <= T1_SCAN * nthr && nh <= lcap && (shi == min(n, sbase + 65536))) {
      int run = wbase;
#pragma unroll
      for (int u = 0; u < T1_SCAN; ++u) {
        const bool hit = (hits_u >> u) & 1u;
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (hit) lst_s[run + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)(tid + u * nthr);
        run += __popc(ball);
      }
    } else {
      int run = wbase;
      for (int k = sbase + tid; k < shi + lane; k += T1_SCAN * nthr) {
        int yv[T1_SCAN];
#pragma unroll
        for (int u = 0; u < T1_SCAN; ++u) { const int s = k + u * nthr; yv[u] = s < shi ? scan_src[s] : scan_none; }
#pragma unroll
        for (int u = 0; u < T1_SCAN; ++u) {
          const bool hit = yv[u] != scan_none && is_hit(yv[u]);
          const unsigned ball = __ballot_sync(0xffffffffu, hit);
          if (hit) lst_s[run + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)(k + u * nthr - sbase);
          run += __popc(ball);
        }
      }
    }
    __syncthreads();
    T1_PHASE(1);
    const int next = shi;
    // ---- flush: evaluate kernels densely, then spread by row ownership -------------------------
    for (int c0 = 0; c0 < nh; c0 += T1_RC) {
      const int cn = min(T1_RC, nh - c0);
      // Hit records.  Single precision: one thread per (hit, dimension, half of the kernel samples), so
      // that the sqrt + exp evaluations of a chunk use most of the CTA's warps; double precision (fewer
      // threads per CTA, longer evaluations): one thread per (hit, dimension).
      constexpr bool kHalves = sizeof(T) == 4;
      constexpr int WH = kHalves ? (WMAX + 1) / 2 : WMAX;      // samples per thread
      for (int t = tid; t < (kHalves ? 4 : 2) * cn; t += nthr) {
        const int hd = kHalves ? t >> 1 : t, half = kHalves ? (t & 1) : 0;
        const int h = hd >> 1, dim = hd & 1, src = sbase + (int)lst_s[c0 + h];
        T z0;
        if (dim == 0) {
          z0 = zxp[src];
          if (half == 0) {
            rec_i0x[h] = wrap_idx(ix0[src], nf);
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) rec_w[h * NP + pp] = Wp[(int64_t)pp * a.n_cap + src];
          }
        } else {
          z0 = zyp[src];
          if (half == 0) {
            int d = iy0[src] - r0;
            if (d < 0) d += nf;
            if (d < 0) d += nf;
            rec_d[h] = d;
          }
        }
        const int jb = half * WH;
        T* kk = (dim == 0 ? rec_kx : rec_ky) + h * WMAX + jb;
        z0 += (T)jb;
#pragma unroll
        for (int j = 0; j < WH; ++j)
          if (jb + j < w) kk[j] = es_kernel<T>(z0 + (T)j, a.beta, a.c, a.halfw);
      }
      __syncthreads();
      T1_PHASE(2);
      if (use_seg) {
        // Spreading without atomics: warp = column segment.  A warp walks, in ascending hit order, the
        // hits whose footprint has columns in its segment (a hit that crosses a segment edge is taken by
        // both neighbours, each adding only its own columns), one hit at a time with the hit's w x w
        // cells spread over the lanes: lane = (footprint row, column group), ceil(w / G) cells each.
        // The cells of one hit are distinct, warps own disjoint columns and a warp's hits are
        // program-ordered: no races, a deterministic sum, and no CTA barrier inside the phase.
        const int s0 = warp * seg, s1 = min(nf, s0 + seg);
        if (warp < nseg && s0 < nf) {
          unsigned char* sl = seg_list + warp * T1_RC;
          // my hits, in order: ballot compaction over the chunk's first columns
          int L = 0;
          for (int hb = 0; hb < cn; hb += 32) {
            const int hl = hb + lane;
            bool mine = false;
            if (hl < cn) {
              const int c0 = rec_i0x[hl], c1 = c0 + w;       // columns [c0, c1), wrapped at nf
              mine = (c0 < s1 && c1 > s0) || (c1 - nf > s0);
            }
            const unsigned ball = __ballot_sync(0xffffffffu, mine);
            if (mine) sl[L + __popc(ball & ((1u << lane) - 1u))] = (unsigned char)hl;
            L += __popc(ball);
          }
          __syncwarp();
          constexpr int CPL = WT > 0 ? (WT + (32 / WT) - 1) / (32 / WT) : kMaxW / 2;   // cells per lane (w <= 16: G >= 2)
          const int jr = min(jrow, w - 1);
          const bool lane_on = jrow < w;
          const unsigned a_sl = smem_addr(sl), a_i0x = smem_addr(rec_i0x), a_d = smem_addr(rec_d);
          const unsigned a_w = smem_addr(rec_w), a_strip = smem_addr(strip);
          const unsigned a_ky = smem_addr(rec_ky) + jr * (unsigned)sizeof(T);
          const unsigned a_kx = smem_addr(rec_kx) + tcol * (unsigned)sizeof(T);
          const unsigned seglen = (unsigned)(s1 - s0);
          // record of the next hit (its loads are issued before the current hit's stores)
          int h = 0, c0 = 0, d = 0;
          T ky = T(0), kx[CPL];
          C cw = make_c<T>(T(0), T(0));
          auto load_rec = [&](int r) {
            h = lds_u8(a_sl + r);
            c0 = lds_i32(a_i0x + 4u * h);
            d = lds_i32(a_d + 4u * h);
            cw = lds_cplx(a_w + (unsigned)(NP * sizeof(C)) * h, T(0));
            ky = lds_real(a_ky + (unsigned)(WMAX * sizeof(T)) * h, T(0));
#pragma unroll
            for (int i = 0; i < CPL; ++i)
              kx[i] = (tcol + i * G < w) ? lds_real(a_kx + (unsigned)(WMAX * sizeof(T)) * h + (unsigned)(i * G * sizeof(T)), T(0)) : T(0);
          };
          if (L > 0) load_rec(0);
          for (int r = 0; r < L; ++r) {
            // this hit's cells
            // this lane's footprint row relative to the strip, wrapped at nf (d is in [0, nf)): a footprint
            // that starts below the strip's end and runs past the grid edge re-enters the strip at row 0
            // (whole-grid strips, or a short last strip), so the wrap is applied per lane, not per hit
            int rr = d + jrow;
            if (rr >= nf) rr -= nf;
            const bool row_ok = lane_on && rr < rows;
            const unsigned a_row = a_strip + (unsigned)(rr * pitch) * (unsigned)sizeof(C);
            unsigned ca[CPL];
            bool ok[CPL];
            T kr[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
              int col = c0 + tcol + i * G;
              if (col >= nf) col -= nf;
              ok[i] = row_ok && (tcol + i * G < w) && (unsigned)(col - s0) < seglen;
              ca[i] = a_row + (unsigned)col * (unsigned)sizeof(C);
              kr[i] = kx[i] * ky;
            }
            const C cwc = cw;
            C v[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) if (ok[i]) v[i] = lds_cplx(ca[i], T(0));
            if (r + 1 < L) load_rec(r + 1);
#pragma unroll
            for (int i = 0; i < CPL; ++i)
              if (ok[i]) { v[i].x += cwc.x * kr[i]; v[i].y += cwc.y * kr[i]; sts_cplx(ca[i], v[i]); }
            __syncwarp();                                  // the next hit may touch the same cells from other lanes
          }
        }
      } else if (warp * ((rows + nwarps - 1) / nwarps) < rows) {
        const int rpw = (rows + nwarps - 1) / nwarps;        // strip rows owned by each warp
        const int rb0 = warp * rpw, rb1 = min(rows, rb0 + rpw);
        // several rows per warp (small grids held whole in one CTA): warp q owns strip rows
        // [q rpw, (q + 1) rpw); one hit at a time, G of its footprint rows per instruction (lane =
        // row in group * w + column).  Lanes of one instruction touch distinct cells and no other
        // warp touches these rows: race-free, deterministic order.
        const int blk = rb1 - rb0;
        for (int hb = 0; hb < cn; hb += 32) {
          const int hl = hb + lane;
          int ja_l = 0, jb_l = 0, dh_l = 0;
          if (hl < cn) {
            dh_l = rec_d[hl];
            int aoff = dh_l - rb0;
            if (aoff < 0) aoff += nf;
            if (aoff < blk) { ja_l = 0; jb_l = min(w, blk - aoff); }
            else { ja_l = nf - aoff; jb_l = min(w, ja_l + blk); }
          }
          unsigned rel = __ballot_sync(0xffffffffu, ja_l < jb_l);
          while (rel) {
            const int src_lane = __ffs(rel) - 1;
            rel &= rel - 1;
            const int h = hb + src_lane;
            const int ja = __shfl_sync(0xffffffffu, ja_l, src_lane);
            const int jb = __shfl_sync(0xffffffffu, jb_l, src_lane);
            const int dh = __shfl_sync(0xffffffffu, dh_l, src_lane);
            const T kxv = (jj < G) ? rec_kx[h * WMAX + jx] : T(0);
            const int col = wrap_idx(rec_i0x[h] + jx, nf);
            C cw[NP];
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) cw[pp] = rec_w[h * NP + pp];
            // The cells one hit touches in this warp's rows are all distinct, so up to IT row groups
            // are loaded before the first store (the loads pipeline) and the warp only synchronises
            // between hits.
            constexpr int IT = NP == 1 ? 4 : 1;
            for (int j0 = ja; j0 < jb; j0 += G * IT) {
              C v[IT][NP];
              C* cell[IT];
              T k2[IT];
              bool on[IT];
#pragma unroll
              for (int i = 0; i < IT; ++i) {
                const int j = j0 + i * G + jj;
                on[i] = jj < G && j < jb;
                cell[i] = strip;
                k2[i] = T(0);
                if (on[i]) {
                  int rr = dh + j;
                  if (rr >= nf) rr -= nf;
                  k2[i] = rec_ky[h * WMAX + j] * kxv;
                  cell[i] = strip + rr * pitch + col;
#pragma unroll
                  for (int pp = 0; pp < NP; ++pp) v[i][pp] = cell[i][pp * pstride];
                }
              }
#pragma unroll
              for (int i = 0; i < IT; ++i) {
                if (on[i]) {
#pragma unroll
                  for (int pp = 0; pp < NP; ++pp) {
                    v[i][pp].x += cw[pp].x * k2[i]; v[i][pp].y += cw[pp].y * k2[i];
                    cell[i][pp * pstride] = v[i][pp];
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
      T1_PHASE(3);
    }
    sbase = next;
  }

  // rows of all products are `pitch` apart (product strips are contiguous), so one call transforms them
  smem_fft<T>(strip, NP == 1 ? rows : NP * a.R, pitch, nf, tw, a.st);
  T1_PHASE(4);
  __syncthreads();
  T1_PHASE(5);

  // needed columns of this strip -> T[col][row] (rows contiguous)
  const int total = a.ncols * rows;
  const unsigned inv_rows = rows > 1 ? 0xFFFFFFFFu / (unsigned)rows + 1u : 0u;   // floor(2^32 / rows) + 1 without a 64-bit division
#pragma unroll
  for (int pp = 0; pp < NP; ++pp) {
    C* Tb = a.Tbuf + (int64_t)(bpi + pp) * a.ncols * nf + r0;
    const C* sp = strip + pp * pstride;
    if (sizeof(C) == 8 && ((rows | r0) & 1) == 0) {
      // two consecutive rows per thread: 16-byte stores (the store-instruction rate bounds this phase)
      const int np2 = rows >> 1, total2 = a.ncols * np2;
      const unsigned inv_np2 = np2 > 1 ? 0xFFFFFFFFu / (unsigned)np2 + 1u : 0u;
      const unsigned a_sp = smem_addr(sp);
#pragma unroll 2
      for (int i = tid; i < total2; i += nthr) {
        const int ci = inv_np2 ? (int)__umulhi((unsigned)i, inv_np2) : i;
        const int pr = i - ci * np2;
        const unsigned ac = a_sp + (unsigned)(2 * pr * pitch + (int)colp[ci]) * 8u;
        const float2 v0 = lds_cplx(ac, 0.f), v1 = lds_cplx(ac + (unsigned)pitch * 8u, 0.f);
        *reinterpret_cast<float4*>(reinterpret_cast<float2*>(Tb) + (int64_t)ci * nf + 2 * pr) = make_float4(v0.x, v0.y, v1.x, v1.y);
      }
    } else {
#pragma unroll 4
      for (int i = tid; i < total; i += nthr) {
        const int ci = inv_rows ? (int)__umulhi((unsigned)i, inv_rows) : i;
        const int rr = i - ci * rows;
        Tb[(int64_t)ci * nf + rr] = sp[rr * pitch + colp[ci]];
      }
    }
  }
  T1_PHASE(6);
#undef T1_PHASE
}

template <typename T>
struct T1GatherArgs {
  const cplx_t<T>* Tbuf;         // (nb, ntr, ncols, nf); (nb, ntr, ceil(ncols / 8), nf, 8) when t_blocked (x-direct pass 1)
  int nf, pitch, ncols, cols_per_cta, ntr, t_blocked;
  const cplx_t<T>* tw;
  FftStages st;
  const int32_t* col_off;        // (ncols + 1) ranges into the column-sorted baseline tables
  const int32_t* s_k;            // baseline index (position in the caller's m1/m2 arrays)
  const int32_t* s_pos;          // smem position of the baseline's second mode number
  const T* s_scale;              // 1 / (phihat(m1) phihat(m2))
  EpiDev epi;
};

template <typename T>
__global__ void __launch_bounds__(T1_THREADS, t1_limits<T>::gather_blocks)
t1_ffty_gather_kernel(T1GatherArgs<T> a) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t1_smem[];
  C* cols = (C*)t1_smem;                                    // cols_per_cta * pitch
  C* tw = cols + (size_t)a.cols_per_cta * a.pitch;          // nf
  const int nf = a.nf, pitch = a.pitch;
  const int bpi = blockIdx.y, b = bpi / a.ntr, p = bpi - b * a.ntr;
  const int c0 = blockIdx.x * a.cols_per_cta;
  const int nc = min(a.cols_per_cta, a.ncols - c0);
  const int tid = threadIdx.x;
  const C* Tb = a.Tbuf + ((int64_t)bpi * a.ncols + c0) * nf;
  // columns -> shared memory with asynchronous copies (LDGSTS): every element's load is in flight
  // at once instead of a register round trip per element
  if (a.t_blocked) {
    // x-direct pass 1 writes T in blocks of 8 columns: [column group][row][8 columns], so a CTA's columns are whole
    // contiguous blocks of nf x 8 entries.  16-byte loads (two columns of one row), five in flight per thread,
    // then one 8-byte shared-memory store per column: the transpose to column vectors happens on the way in.
    const int ncg = (a.ncols + 7) >> 3;
    if (sizeof(C) == 8 && (c0 & 7) == 0) {
      for (int g = 0; g * 8 < nc; ++g) {
        const float4* blk = reinterpret_cast<const float4*>(a.Tbuf) + ((int64_t)bpi * ncg + (c0 >> 3) + g) * nf * 4;
        float2* cg = reinterpret_cast<float2*>(cols) + (size_t)g * 8 * pitch;
        const int ncg_cols = min(8, nc - g * 8);
        const int tot = nf * 4;
        for (int i0 = tid; i0 < tot; i0 += 5 * blockDim.x) {
          float4 v[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) { const int i = i0 + u * blockDim.x; if (i < tot) v[u] = blk[i]; }
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < tot) {
              const int rr = i >> 2, cc = (i & 3) * 2;
              if (cc < ncg_cols) cg[cc * pitch + rr] = make_float2(v[u].x, v[u].y);
              if (cc + 1 < ncg_cols) cg[(cc + 1) * pitch + rr] = make_float2(v[u].z, v[u].w);
            }
          }
        }
      }
    } else {
      for (int i = tid; i < nc * nf; i += blockDim.x) {
        const int ci = i / nf, rr = i - ci * nf, cgl = c0 + ci;
        __pipeline_memcpy_async(cols + ci * pitch + rr,
                                a.Tbuf + (((int64_t)bpi * ncg + (cgl >> 3)) * nf + rr) * 8 + (cgl & 7), sizeof(C));
      }
    }
  } else {
    for (int ci = tid >> 5; ci < nc; ci += blockDim.x >> 5) {
      const C* src = Tb + (int64_t)ci * nf;
      C* dst = cols + ci * pitch;
      for (int rr = tid & 31; rr < nf; rr += 32) __pipeline_memcpy_async(dst + rr, src + rr, sizeof(C));
    }
  }
  __pipeline_commit();
  for (int i = tid; i < a.st.tw_len; i += blockDim.x) tw[i] = a.tw[i];
  __pipeline_wait_prior(0);
  __syncthreads();
  smem_fft<T>(cols, nc, pitch, nf, tw, a.st);
  __syncthreads();
  const int k0 = a.col_off[c0], k1 = a.col_off[c0 + nc];
  // walk the column-sorted baselines of this CTA's columns
  for (int kk = k0 + tid; kk < k1; kk += blockDim.x) {
    // column of kk: binary search in col_off[c0 .. c0 + nc]
    int lo = 0, hi = nc;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.col_off[c0 + mid] <= kk) lo = mid; else hi = mid; }
    C v = cols[lo * pitch + a.s_pos[kk]];
    const T sc = a.s_scale[kk];
    v.x *= sc; v.y *= sc;
    epilogue_store(a.epi, b, p, (int64_t)a.s_k[kk], v);
  }
}

// ---- pass 2 on the blocked T of the x-direct pass 1 (single precision) --------------------------------------------
// One CTA per block of 8 columns.  The block ([row][8 columns], nf x 64 bytes, contiguous) comes in by bulk copies
// (TMA, cp.async.bulk + mbarrier) and is transformed IN that layout: a lane is (column, butterfly), so the 32 lanes of
// every shared-memory access cover 4 rows x 8 columns = 256 contiguous bytes, and no transpose or padding is needed.
// All warps share the 8 columns, so the stages are separated by CTA barriers.
template <int R>
__device__ __forceinline__ void fft_stage_il(float2* d, int N, int n, unsigned inv, const float2* __restrict__ tws,
                                             int tid, int nthr) {
  const int m = n / R, per_vec = N / R;
  const int col = tid & 7;
  for (int t = tid >> 3; t < per_vec; t += nthr >> 3) {
    const int blk = inv ? (int)__umulhi((unsigned)t, inv) : t;
    const int j = t - blk * m;
    float2* p = d + (blk * n + j) * 8 + col;
    float2 x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = p[q * m * 8];
    dft_r<float, R>(x);
    p[0] = x[0];
    if (m > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) p[q * m * 8] = cmul(x[q], tws[(q - 1) * m + j]);
    } else {
#pragma unroll
      for (int q = 1; q < R; ++q) p[q * 8] = x[q];
    }
  }
}

template <typename T>                                     // T = float only (a template so that every translation unit may include it)
__global__ void __launch_bounds__(T1_THREADS, 3)
t1_ffty8_gather_kernel(T1GatherArgs<T> a) {
  static_assert(sizeof(T) == 4, "single precision");
  using C = float2;
  extern __shared__ __align__(128) unsigned char t1_smem[];
  __shared__ __align__(8) unsigned long long bar;
  C* blk = (C*)t1_smem;                                     // [nf][8]
  C* tw = blk + (size_t)a.nf * 8;                           // twiddles
  const int nf = a.nf;
  const int bpi = blockIdx.y, b = bpi / a.ntr, p = bpi - b * a.ntr;
  const int c0 = blockIdx.x * 8;
  const int nc = min(8, a.ncols - c0);
  const int tid = threadIdx.x;
  const int ncg = (a.ncols + 7) >> 3;
  const unsigned bar_a = smem_u32(&bar);
  if (tid == 0) mbar_init(bar_a, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    // the block in runs of at most 16 KB, one per lane
    const unsigned total = (unsigned)nf * 64u;
    const unsigned tw_bytes = ((unsigned)a.st.tw_len * 8u + 15u) & ~15u;     // the table is allocated with 16 bytes to spare
    const char* src = reinterpret_cast<const char*>(a.Tbuf) + ((int64_t)bpi * ncg + blockIdx.x) * (int64_t)total;
    if (tid == 0) mbar_expect_tx(bar_a, total + tw_bytes);
    __syncwarp();
    for (unsigned off = (unsigned)tid * 16384u; off < total; off += 32u * 16384u)
      bulk_g2s(smem_u32(blk) + off, src + off, min(16384u, total - off), bar_a);
    if (tid == 31) bulk_g2s(smem_u32(tw), a.tw, tw_bytes, bar_a);
  }
  const int k0 = a.col_off[c0], k1 = a.col_off[c0 + nc];
  mbar_wait(bar_a, 0u);
  __syncthreads();
  int n = nf;
  for (int s = 0; s < a.st.nstage; ++s) {
    const unsigned inv = a.st.inv_m[s];
    const C* tws = tw + a.st.tw_off[s];
    switch (a.st.radix[s]) {
      case 15: fft_stage_il<15>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 15; break;
      case 8: fft_stage_il<8>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 8; break;
      case 4: fft_stage_il<4>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 4; break;
      case 2: fft_stage_il<2>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 2; break;
      case 5: fft_stage_il<5>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 5; break;
      default: fft_stage_il<3>(blk, nf, n, inv, tws, tid, blockDim.x); n /= 3; break;
    }
    __syncthreads();
  }
  // walk the column-sorted baselines of this CTA's columns
  for (int kk = k0 + tid; kk < k1; kk += blockDim.x) {
    int lo = 0, hi = nc;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.col_off[c0 + mid] <= kk) lo = mid; else hi = mid; }
    C v = blk[a.s_pos[kk] * 8 + lo];
    const float sc = a.s_scale[kk];
    v.x *= sc; v.y *= sc;
    epilogue_store(a.epi, b, p, (int64_t)a.s_k[kk], v);
  }
}

}  // namespace fv
