// Pass 1 of the fused type-1 path without an x grid ("x-direct"), single precision.
//
// The half-transformed array the y pass needs is
//     T[col][row] = sum_s  W_s * phi_y(row - y_s) * exp(+i k_col x_s)          k_col = first mode number of column col
// i.e. gridded in y (exponential-of-semicircle kernel, as in finufft.nufft2d1, reference cpu/nufft.py:120-175) but an
// exact Fourier sum in x: only n_cols <= n_modes of the nf columns are ever read by a baseline, so spreading w cells
// in x, transforming nf-point rows and discarding most outputs costs more than evaluating the n_cols phases
// directly.
//
//   t1_xd_prep_kernel   one CTA per frequency: fold every source once (first footprint row, kernel argument, x as a
//                       32-bit fraction of a turn) and build, by a stable counting sort, the list of sources whose
//                       w-row footprint touches each strip of R rows.
//   t1_xdirect_kernel   one CTA per (strip, frequency x product, group of <= 256 columns); a thread owns ONE column and
//                       keeps its R accumulators in registers.  Per (hit, column): the phase k * x + arg(W) by 32-bit
//                       fixed-point turn arithmetic (wraps mod one turn exactly), one sin/cos pair, and 2 w fused
//                       multiply-adds with the |W|-scaled kernel samples into the registers of the footprint's rows
//                       (the hit records of a chunk are sorted by first row, so each first row is its own branch-free
//                       loop): no shared-memory grid, no atomics, no FFT, no barrier inside the hit loops.
//
// The y pass (t1_ffty_gather_kernel) then divides by the kernel transform in y only.  T is stored in blocks of 8
// columns, [column group][row][8]: a warp stores four 64-byte runs per row and the y pass reads whole blocks.
//
// Error model: x is exact up to the phase rounding (2^-32 turn * |k| + the sin/cos approximation, ~4e-7 absolute),
// y is the finufft kernel at the requested width: never worse than the two-dimensional spreader it replaces.
// Summation order is fixed (list order, then first row): results are bitwise reproducible.
#pragma once

namespace fv {

constexpr int XD_ROWS = 24;            // strip height: 2 * 24 accumulator registers per thread, three CTAs per SM
constexpr int XD_THREADS = 256;        // columns per CTA
constexpr int XD_RC = 256;             // hit records per chunk (one thread fills one record)
constexpr int XD_PREP_THREADS = 1024;
constexpr int XD_PREP_K = 2;           // sources per thread per block of the prep pass
constexpr int XD_PREP_SPLIT = 4;       // CTAs per frequency in the prep pass, each with its own strip lists over a quarter of the sources
constexpr int XD_LIST_PER_SRC = 3;     // a footprint of w < R rows touches at most 3 strips (short last strip + wrap)

// record: ky[0..W) scaled by |W_s|, padding, x (turn fraction), arg W_s (turn fraction)
template <int W> struct XdRec { static constexpr int LEN = 4 * ((W + 2 + 3) / 4); };

struct T1XdPrepArgs {
  const float* bx; const float* by;
  const int32_t* n_dev;
  int64_t n_cap;
  const BatchParams* bp;
  int nf, w, nstrips;            // strips of XD_ROWS rows
  int32_t* yrow;                 // first footprint row wrapped into [0, nf), (nb, n_cap)
  float* zy;                     // kernel argument of that row
  uint32_t* xt;                  // x as a fraction of a turn, 32-bit fixed point (x = 0 -> 0)
  int32_t* strip_off;            // (nb, XD_PREP_SPLIT, nstrips + 1) ranges into the part's region of `list`
  int32_t* list;                 // (nb, XD_PREP_SPLIT, lcap) source indices, strip-major, ascending inside a strip
  int64_t lcap;                  // XD_LIST_PER_SRC * ceil(n_cap / XD_PREP_SPLIT)
};

// strips a footprint starting at (wrapped) row y touches: the one holding y, the next one, strip 0 after the wrap
__device__ __forceinline__ void xd_strips(int y, int w, int nf, int (&sid)[XD_LIST_PER_SRC]) {
  const int last = y + w - 1;
  sid[0] = y / XD_ROWS;
  const int s1 = min(last, nf - 1) / XD_ROWS;
  sid[1] = s1 != sid[0] ? s1 : -1;
  sid[2] = last >= nf ? 0 : -1;
}

__global__ void __launch_bounds__(XD_PREP_THREADS)
t1_xd_prep_kernel(T1XdPrepArgs a) {
  extern __shared__ int xd_sm[];
  constexpr int NW = XD_PREP_THREADS / 32;
  const int ns = a.nstrips;
  int* wc = xd_sm;                       // [NW][ns] per-warp count, then running offset, of every strip
  int* tot = wc + NW * ns;               // [ns + 1]
  int* wsum = tot + ns + 1;              // [NW] scan scratch
  const int b = blockIdx.x, part = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this CTA's share of the live sources: [lo, n)
  const int n_all = *a.n_dev, per_part = (n_all + XD_PREP_SPLIT - 1) / XD_PREP_SPLIT;
  const int lo = min(n_all, part * per_part), n = min(n_all, lo + per_part);
  const int64_t o0 = (int64_t)b * a.n_cap;
  const float smul = (float)a.bp[b].smul;
  const int nf = a.nf, w = a.w;
  const double hw = 0.5 * (double)w;
  for (int i = tid; i < NW * ns; i += XD_PREP_THREADS) wc[i] = 0;
  __syncthreads();
  // pass 1: fold (fp64 fold of the working-precision product fl(coordinate * freq), reference
  // cpu_simulate.py:990-992) and count; XD_PREP_K sources per thread with all of their loads in flight together
  for (int s0 = lo; s0 < n; s0 += XD_PREP_K * XD_PREP_THREADS) {
    int sid[XD_PREP_K][XD_LIST_PER_SRC];
    float cx[XD_PREP_K], cy[XD_PREP_K];
#pragma unroll
    for (int k = 0; k < XD_PREP_K; ++k) {
      const int s = s0 + k * XD_PREP_THREADS + tid;
      cx[k] = s < n ? a.bx[s] : 0.f;
      cy[k] = s < n ? a.by[s] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < XD_PREP_K; ++k) {
      const int s = s0 + k * XD_PREP_THREADS + tid;
      sid[k][0] = sid[k][1] = sid[k][2] = -1;
      if (s < n) {
        const double gy = fold_grid((double)(cy[k] * smul), nf), giy = ceil(gy - hw);
        int y = (int)giy;
        if (y < 0) y += nf;
        a.yrow[o0 + s] = y;
        a.zy[o0 + s] = (float)(giy - gy);
        double r = (double)(cx[k] * smul) * 0.15915494309189533577;
        r -= floor(r);
        a.xt[o0 + s] = (uint32_t)(unsigned long long)(r * 4294967296.0 + 0.5);
        xd_strips(y, w, nf, sid[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < XD_PREP_K; ++k) {
#pragma unroll
      for (int q = 0; q < XD_LIST_PER_SRC; ++q) {
        if (q > 0 && !__any_sync(0xffffffffu, sid[k][q] >= 0)) continue;       // warp-uniform
        const unsigned peers = __match_any_sync(0xffffffffu, sid[k][q]);
        if (sid[k][q] >= 0 && (peers & ((1u << lane) - 1u)) == 0) wc[warp * ns + sid[k][q]] += __popc(peers);
        __syncwarp();
      }
    }
  }
  __syncthreads();
  // per strip: totals over the warps; per-warp counts -> offsets inside the strip's range
  for (int t = tid; t < ns; t += XD_PREP_THREADS) {
    int acc = 0;
    for (int wv = 0; wv < NW; ++wv) { const int c = wc[wv * ns + t]; wc[wv * ns + t] = acc; acc += c; }
    tot[t] = acc;
  }
  __syncthreads();
  // exclusive scan of the strip totals (ns may exceed the CTA: chunks of 1024 with a running base)
  {
    int base = 0;
    for (int t0 = 0; t0 < ns; t0 += XD_PREP_THREADS) {
      const int t = t0 + tid;
      const int v = t < ns ? tot[t] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      int wb, all;
      {
        const int u = wsum[lane];                           // NW == 32: one warp total per lane
        int winc = u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += x; }
        all = __shfl_sync(0xffffffffu, winc, 31);
        wb = __shfl_sync(0xffffffffu, winc - u, warp);
      }
      __syncthreads();
      if (t < ns) tot[t] = base + wb + inc - v;
      base += all;
    }
    if (tid == 0) tot[ns] = base;
  }
  __syncthreads();
  for (int t = tid; t <= ns; t += XD_PREP_THREADS) a.strip_off[((int64_t)b * XD_PREP_SPLIT + part) * (ns + 1) + t] = tot[t];
  // pass 2: place (same order of visits as pass 1: ascending source index inside every strip)
  int32_t* lst = a.list + ((int64_t)b * XD_PREP_SPLIT + part) * a.lcap;
  for (int s0 = lo; s0 < n; s0 += XD_PREP_K * XD_PREP_THREADS) {
    int yv[XD_PREP_K];
#pragma unroll
    for (int k = 0; k < XD_PREP_K; ++k) {
      const int s = s0 + k * XD_PREP_THREADS + tid;
      yv[k] = s < n ? a.yrow[o0 + s] : -1;
    }
#pragma unroll
    for (int k = 0; k < XD_PREP_K; ++k) {
      const int s = s0 + k * XD_PREP_THREADS + tid;
      int sid[XD_LIST_PER_SRC] = {-1, -1, -1};
      if (yv[k] >= 0) xd_strips(yv[k], w, nf, sid);
#pragma unroll
      for (int q = 0; q < XD_LIST_PER_SRC; ++q) {
        if (q > 0 && !__any_sync(0xffffffffu, sid[q] >= 0)) continue;          // warp-uniform
        const unsigned peers = __match_any_sync(0xffffffffu, sid[q]);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int cur = 0;
        if (sid[q] >= 0) cur = wc[warp * ns + sid[q]];
        __syncwarp();
        if (sid[q] >= 0) {
          lst[tot[sid[q]] + cur + rank] = s;
          if (rank == 0) wc[warp * ns + sid[q]] = cur + __popc(peers);
        }
        __syncwarp();
      }
    }
  }
}

struct T1XdArgs {
  int64_t n_cap;
  int nf, nstrips;
  float beta, c, halfw;
  int ntr;
  const float2* W;               // (nb, ntr, n_cap)
  const int32_t* yrow; const float* zy; const uint32_t* xt;   // from t1_xd_prep_kernel
  const int32_t* strip_off; const int32_t* list; int64_t lcap;
  int ncols;
  const int32_t* col_k;          // signed first mode number of every needed column
  float2* Tbuf;                  // (nb, ntr, ceil(ncols / 8), nf, 8)
};

// acc[c - (W - 1) + j] += ky[j] * p for the rows of the footprint that fall inside the strip
template <int W, int R, int CASE>
__device__ __forceinline__ void xd_update(float2 (&acc)[R], const float (&ky)[W], float px, float py) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    constexpr int r0 = CASE - (W - 1);
    const int r = r0 + j;
    if (r >= 0 && r < R) { acc[r].x = fmaf(ky[j], px, acc[r].x); acc[r].y = fmaf(ky[j], py, acc[r].y); }
  }
}

// All hits of one case (= first footprint row relative to the strip), then the next case: the hit records of a
// chunk are sorted by case, so each case is a branch-free loop the compiler can pipeline.
template <int W, int R, int CASE>
__device__ __forceinline__ void xd_case_loop(float2 (&acc)[R], const float* __restrict__ rec, const int* __restrict__ cbase,
                                             unsigned k_me) {
  constexpr int LEN = XdRec<W>::LEN;
  const int h0 = cbase[CASE], h1 = cbase[CASE + 1];
#pragma unroll 2
  for (int h = h0; h < h1; ++h) {
    const float4* rp = reinterpret_cast<const float4*>(rec + h * LEN);
    float q[LEN];
#pragma unroll
    for (int v = 0; v < LEN / 4; ++v) {
      const float4 t = rp[v];
      q[4 * v] = t.x; q[4 * v + 1] = t.y; q[4 * v + 2] = t.z; q[4 * v + 3] = t.w;
    }
    float ky[W];
#pragma unroll
    for (int j = 0; j < W; ++j) ky[j] = q[j];
    // k * x + arg(W) mod one turn, exactly, as a signed fraction of a turn
    const int ph = (int)(k_me * __float_as_uint(q[LEN - 2]) + __float_as_uint(q[LEN - 1]));
    const float ang = (float)ph * 1.4629180792671596e-9f;          // 2 pi / 2^32
    xd_update<W, R, CASE>(acc, ky, __cosf(ang), __sinf(ang));
  }
  if constexpr (CASE + 1 < R + W - 1) xd_case_loop<W, R, CASE + 1>(acc, rec, cbase, k_me);
}

template <int W, int R>
__global__ void __launch_bounds__(XD_THREADS, 3)
t1_xdirect_kernel(T1XdArgs a) {
  static_assert(R == XD_ROWS, "the prep pass bins by XD_ROWS");
  static_assert(XD_PREP_THREADS == 1024, "block scan of the prep pass assumes 32 warps");
  static_assert(R + W - 1 <= 62, "case 63 marks an idle fill lane");
  constexpr int LEN = XdRec<W>::LEN, NCASE = R + W - 1;
  __shared__ __align__(16) float rec[XD_RC * LEN];
  __shared__ int cbase[NCASE + 1];                          // start of every case's records in the sorted chunk
  __shared__ unsigned short wc[XD_THREADS / 32][NCASE];     // per-warp count (then offset) of every case

  const int nf = a.nf;
  const int bpi = blockIdx.y, b = bpi / a.ntr;
  const int r0 = blockIdx.x * R;
  const int rows = min(R, nf - r0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nthr = XD_THREADS, nwarps = XD_THREADS / 32;
  const int col = blockIdx.z * XD_THREADS + tid;
  const unsigned k_me = col < a.ncols ? (unsigned)a.col_k[col] : 0u;
  const int32_t* yrow = a.yrow + (int64_t)b * a.n_cap;
  const float* zyp = a.zy + (int64_t)b * a.n_cap;
  const uint32_t* xtp = a.xt + (int64_t)b * a.n_cap;
  const float2* Wp = a.W + (int64_t)bpi * a.n_cap;
  // this strip's hits: the XD_PREP_SPLIT parts' lists one after the other (ascending source index overall)
  int p_lo[XD_PREP_SPLIT], p_end[XD_PREP_SPLIT], nh = 0;    // start in the part's list; end in the concatenated numbering
#pragma unroll
  for (int g = 0; g < XD_PREP_SPLIT; ++g) {
    const int32_t* off = a.strip_off + ((int64_t)b * XD_PREP_SPLIT + g) * (a.nstrips + 1) + blockIdx.x;
    p_lo[g] = off[0];
    nh += off[1] - off[0];
    p_end[g] = nh;
  }
  const int32_t* lst = a.list + (int64_t)b * XD_PREP_SPLIT * a.lcap;

  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);

  for (int c0 = 0; c0 < nh; c0 += XD_RC) {
    const int cn = min(XD_RC, nh - c0);
    // one hit record per thread, written at its place in the order (case, list position) found by a counting
    // sort over the chunk
    for (int i = tid; i < nwarps * NCASE; i += nthr) (&wc[0][0])[i] = 0;
    int cs = 63, src = 0;
    if (tid < cn) {
      int g = 0, i = c0 + tid;
#pragma unroll
      for (int u = 0; u < XD_PREP_SPLIT - 1; ++u) g += i >= p_end[u] ? 1 : 0;
      int idx = p_lo[0] + i;
#pragma unroll
      for (int u = 1; u < XD_PREP_SPLIT; ++u) if (g == u) idx = p_lo[u] + i - p_end[u - 1];
      src = lst[g * a.lcap + idx];
      int d = yrow[src] - r0;                               // first footprint row relative to the strip
      if (d >= R) d -= nf;                                  // the footprint wraps around the grid edge into this strip
      cs = min(max(d + (W - 1), 0), NCASE - 1);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, cs);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    __syncthreads();                                        // wc cleared; the previous chunk's records are consumed
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
      const float mag = sqrtf(wv.x * wv.x + wv.y * wv.y);
      // arg W as a fraction of a turn (fp64: one evaluation per hit and strip)
      const double turn = atan2((double)wv.y, (double)wv.x) * 0.15915494309189533577;
      float q[LEN];
#pragma unroll
      for (int j = 0; j < LEN; ++j) q[j] = j < W ? mag * es_kernel<float>(z0 + (float)j, a.beta, a.c, a.halfw) : 0.f;
      q[LEN - 2] = __uint_as_float(xtp[src]);
      q[LEN - 1] = __uint_as_float((uint32_t)(long long)__double2ll_rn(turn * 4294967296.0));
#pragma unroll
      for (int v = 0; v < LEN / 4; ++v)
        *reinterpret_cast<float4*>(rp + 4 * v) = make_float4(q[4 * v], q[4 * v + 1], q[4 * v + 2], q[4 * v + 3]);
    }
    __syncthreads();
    xd_case_loop<W, R, 0>(acc, rec, cbase, k_me);
  }

  if (col < a.ncols) {
    // T in blocks of 8 columns, [column group][row][8]: a warp stores four 64-byte runs per row, and pass 2 reads
    // its 8 columns as one contiguous block
    const int ncg = (a.ncols + 7) >> 3;
    float2* Tb = a.Tbuf + (((int64_t)bpi * ncg + (col >> 3)) * nf + r0) * 8 + (col & 7);
#pragma unroll
    for (int r = 0; r < R; ++r) if (r < rows) Tb[r * 8] = acc[r];
  }
}

}  // namespace fv
