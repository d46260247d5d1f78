// Pass 1 of the fused type-1 path without an x grid ("x-direct"), single precision.
//
// The half-transformed array the y pass needs is
//     T[col][row] = sum_s  W_s * phi_y(row - y_s) * exp(+i k_col x_s)          k_col = first mode number of column col
// i.e. gridded in y (exponential-of-semicircle kernel, as in finufft.nufft2d1, reference cpu/nufft.py:120-175) but an
// exact Fourier sum in x: only n_cols <= n_modes of the nf columns are ever read by a baseline, so spreading w cells
// in x, transforming nf-point rows and discarding most outputs costs more than evaluating the n_cols phases
// directly.  One CTA owns a strip of R rows for a group of <= 256 columns; a thread owns ONE column and keeps its R
// accumulators in registers.  Per (hit, column): one phase by 32-bit fixed-point turn arithmetic (k * x wraps mod one
// turn exactly), one sin/cos pair, one complex product with the source strength, and 2 w fused multiply-adds selected
// by a warp-uniform switch on the hit's first row: no shared-memory grid, no atomics, no FFT, no barrier inside the
// hit loop.  The y pass (t1_ffty_gather_kernel) then divides by the kernel transform in y only.
//
// Error model: x is exact up to the phase rounding (2^-32 turn * |k| + the sin/cos approximation, ~4e-7 absolute),
// y is the finufft kernel at the requested width: never worse than the two-dimensional spreader it replaces.
#pragma once

namespace fv {

constexpr int XD_THREADS = 256;   // columns per CTA
constexpr int XD_RC = 256;        // hit records per chunk (one thread fills one record)
constexpr int XD_LCAP = 2048;     // hit-list entries (16-bit offsets from the range start)
constexpr int XD_SCAN = 8;        // sources per thread per scan iteration (all loads in flight together)

template <int W> struct XdRec { static constexpr int LEN = 4 * ((W + 4 + 3) / 4); };   // floats per record

struct T1XdArgs {
  const int32_t* n_dev;
  int64_t n_cap;
  int nf, R, w;
  float beta, c, halfw;
  int ntr;
  const float2* W;               // (nb, ntr, n_cap)
  const int32_t* iy0;            // first footprint row (may be < 0), (nb, n_cap)
  const float* zy;               // kernel argument of that row
  const uint32_t* xt;            // x as a fraction of a turn, 32-bit fixed point
  const uint32_t* hm0; const uint32_t* hm1;   // strip masks (null: more than 64 strips)
  int ncols;
  const int32_t* col_k;          // signed first mode number of every needed column
  float2* Tbuf;                  // (nb, ntr, ncols, nf)
};

// acc[c - (W - 1) + j] += ky[j] * p for the rows of the footprint that fall inside the strip
template <int W, int R, int CASE>
__device__ __forceinline__ void xd_update(float2 (&acc)[R], const float (&ky)[W], float px, float py) {
  if constexpr (CASE < R + W - 1) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      constexpr int r0 = CASE - (W - 1);
      const int r = r0 + j;
      if (r >= 0 && r < R) { acc[r].x = fmaf(ky[j], px, acc[r].x); acc[r].y = fmaf(ky[j], py, acc[r].y); }
    }
  }
}

// All hits of one switch case (= first footprint row relative to the strip), then the next case: the hit
// records of a chunk are sorted by case, so each case is a branch-free loop the compiler can pipeline.
template <int W, int R, int CASE>
__device__ __forceinline__ void xd_case_loop(float2 (&acc)[R], const float* __restrict__ rec, const int* __restrict__ cbase,
                                             unsigned k_me) {
  constexpr int LEN = XdRec<W>::LEN;
  const int h0 = cbase[CASE], h1 = cbase[CASE + 1];
#pragma unroll 2
  for (int h = h0; h < h1; ++h) {
    const float4* rp = reinterpret_cast<const float4*>(rec + h * LEN);
    const float4 q0 = rp[0];
    float ky[W];
#pragma unroll
    for (int v = 0; v < (LEN - 4) / 4; ++v) {
      const float4 q = rp[1 + v];
      if (4 * v + 0 < W) ky[4 * v + 0] = q.x;
      if (4 * v + 1 < W) ky[4 * v + 1] = q.y;
      if (4 * v + 2 < W) ky[4 * v + 2] = q.z;
      if (4 * v + 3 < W) ky[4 * v + 3] = q.w;
    }
    // k * x mod one turn, exactly, as a signed fraction of a turn
    const int ph = (int)(k_me * __float_as_uint(q0.x));
    const float ang = (float)ph * 1.4629180792671596e-9f;          // 2 pi / 2^32
    const float sn = __sinf(ang), cs = __cosf(ang);
    const float px = q0.y * cs - q0.z * sn, py = q0.y * sn + q0.z * cs;
    xd_update<W, R, CASE>(acc, ky, px, py);
  }
  if constexpr (CASE + 1 < R + W - 1) xd_case_loop<W, R, CASE + 1>(acc, rec, cbase, k_me);
}

template <int W, int R>
__global__ void __launch_bounds__(XD_THREADS, 3)
t1_xdirect_kernel(T1XdArgs a) {
  static_assert(R % 2 == 0 && R + W - 1 <= 62, "even strips; case 63 marks an idle fill lane");
  constexpr int LEN = XdRec<W>::LEN, NCASE = R + W - 1;
  __shared__ __align__(16) float rec[XD_RC * LEN];
  __shared__ unsigned short lst_s[XD_LCAP];
  __shared__ int wcnt[XD_THREADS / 32];
  __shared__ int cbase[NCASE + 1];                          // start of every case's records in the sorted chunk
  __shared__ unsigned short wc[XD_THREADS / 32][NCASE];     // per-warp count (then offset) of every case

  const int nf = a.nf;
  const int bpi = blockIdx.y, b = bpi / a.ntr;
  const int r0 = blockIdx.x * R;
  const int rows = min(R, nf - r0);
  const int n = *a.n_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nthr = XD_THREADS, nwarps = XD_THREADS / 32;
  const int col = blockIdx.z * XD_THREADS + tid;
  const unsigned k_me = col < a.ncols ? (unsigned)a.col_k[col] : 0u;
  const int32_t* iy0 = a.iy0 + (int64_t)b * a.n_cap;
  const float* zyp = a.zy + (int64_t)b * a.n_cap;
  const uint32_t* xtp = a.xt + (int64_t)b * a.n_cap;
  const float2* Wp = a.W + (int64_t)bpi * a.n_cap;

  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);

  const bool masked = a.hm0 != nullptr;
  const int32_t* scan_src = masked ? reinterpret_cast<const int32_t*>((blockIdx.x < 32 ? a.hm0 : a.hm1) + (int64_t)b * a.n_cap) : iy0;
  const int scan_bit = 1 << (blockIdx.x & 31);
  const int scan_none = masked ? 0 : INT_MIN;
  auto is_hit = [&](int v) {
    if (masked) return (v & scan_bit) != 0;
    int d = v - r0;
    if (d < 0) d += nf;
    if (d < 0) d += nf;
    return d < rows || d + W > nf;
  };

  int sbase = 0;
  while (sbase < n) {
    // hit list of a range of sources: count per warp, then store at deterministic offsets (warp-major order)
    int shi = min(n, sbase + 65536);
    int nh = 0, wbase = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      int cnt = 0;
      for (int k = sbase + tid; k < shi + lane; k += XD_SCAN * nthr) {     // warp-uniform trip count
        int yv[XD_SCAN];
#pragma unroll
        for (int u = 0; u < XD_SCAN; ++u) { const int s = k + u * nthr; yv[u] = s < shi ? scan_src[s] : scan_none; }
#pragma unroll
        for (int u = 0; u < XD_SCAN; ++u) {
          const bool hit = yv[u] != scan_none && is_hit(yv[u]);
          cnt += __popc(__ballot_sync(0xffffffffu, hit));
        }
      }
      if (lane == 0) wcnt[warp] = cnt;
      __syncthreads();
      {
        int c = lane < nwarps ? wcnt[lane] : 0, inc = c;
#pragma unroll
        for (int o = 1; o < nwarps; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        nh = __shfl_sync(0xffffffffu, inc, nwarps - 1);
        wbase = __shfl_sync(0xffffffffu, inc - c, warp);
      }
      __syncthreads();                                     // wcnt may be rewritten by the next attempt
      if (nh <= XD_LCAP) break;
      shi = min(n, sbase + XD_LCAP);                       // dense strip: a range whose hits fit in any case
    }
    {
      int run = wbase;
      for (int k = sbase + tid; k < shi + lane; k += XD_SCAN * nthr) {
        int yv[XD_SCAN];
#pragma unroll
        for (int u = 0; u < XD_SCAN; ++u) { const int s = k + u * nthr; yv[u] = s < shi ? scan_src[s] : scan_none; }
#pragma unroll
        for (int u = 0; u < XD_SCAN; ++u) {
          const bool hit = yv[u] != scan_none && is_hit(yv[u]);
          const unsigned ball = __ballot_sync(0xffffffffu, hit);
          if (hit) lst_s[run + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)(k + u * nthr - sbase);
          run += __popc(ball);
        }
      }
    }
    __syncthreads();

    for (int c0 = 0; c0 < nh; c0 += XD_RC) {
      const int cn = min(XD_RC, nh - c0);
      // one hit record per thread: phase word, strength, the w kernel samples along y; written at its place
      // in the order (case, hit index) found by a counting sort over the chunk
      for (int i = tid; i < nwarps * NCASE; i += nthr) (&wc[0][0])[i] = 0;
      int cs = 63, src = 0;
      if (tid < cn) {
        src = sbase + (int)lst_s[c0 + tid];
        int y = iy0[src];
        if (y < 0) y += nf;
        int d = y - r0;                                     // first footprint row relative to the strip
        if (d >= R) d -= nf;                                // the footprint wraps around the grid edge into this strip
        cs = min(max(d + (W - 1), 0), NCASE - 1);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, cs);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      __syncthreads();                                      // wc cleared
      if (tid < cn && rank == 0) wc[warp][cs] = (unsigned short)__popc(peers);
      __syncthreads();
      if (warp == 0) {
        // per case: totals over the warps -> start of the case; per-warp counts -> offsets inside the case
        int tot[2] = {0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = lane + 32 * u;
          if (c < NCASE) {
#pragma unroll
            for (int wv = 0; wv < nwarps; ++wv) { const int t = wc[wv][c]; wc[wv][c] = (unsigned short)tot[u]; tot[u] += t; }
          }
        }
        int run = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int inc = tot[u];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
          const int c = lane + 32 * u;
          if (c < NCASE) cbase[c] = run + inc - tot[u];
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cbase[NCASE] = run;
      }
      __syncthreads();
      if (tid < cn) {
        float* rp = rec + (cbase[cs] + (int)wc[warp][cs] + rank) * LEN;
        const float2 wv = Wp[src];
        const float z0 = zyp[src];
        *reinterpret_cast<float4*>(rp) = make_float4(__uint_as_float(xtp[src]), wv.x, wv.y, 0.f);
#pragma unroll
        for (int v = 0; v < (LEN - 4) / 4; ++v) {
          float kk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) kk[j] = 4 * v + j < W ? es_kernel<float>(z0 + (float)(4 * v + j), a.beta, a.c, a.halfw) : 0.f;
          *reinterpret_cast<float4*>(rp + 4 + 4 * v) = make_float4(kk[0], kk[1], kk[2], kk[3]);
        }
      }
      __syncthreads();
      xd_case_loop<W, R, 0>(acc, rec, cbase, k_me);
      __syncthreads();
    }
    sbase = shi;
  }

  if (col < a.ncols) {
    float2* Tb = a.Tbuf + ((int64_t)bpi * a.ncols + col) * nf + r0;
    if (rows == R && ((r0 | nf) & 1) == 0) {
#pragma unroll
      for (int r = 0; r < R; r += 2)
        *reinterpret_cast<float4*>(Tb + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) if (r < rows) Tb[r] = acc[r];
    }
  }
}

}  // namespace fv
