// Type-3 spreading (3-D, thin in z) with bin-sorted tiles and NO atomics on the grid.
//
// finufft's type-3 step 1 spreads every rescaled source onto w^3 cells of the fine grid
// (finufft.nufft3d3, reference cpu/nufft.py:105-118).  A global-atomics spreader issues 2 w^3 REDs per
// source (5500 at w = 14) and is bound by L2 atomic throughput.  Interferometer arrays are nearly
// flat, so the z extent of the grid is the minimum 2w..32 cells: here a thread OWNS one (x, y) column
// of the grid and keeps all of its z cells in registers.
//
//   t3_bin_kernel (count / fill) counting sort of the sources into the 16 x 16-column tiles their
//                                footprint touches (a source lands in up to 4 lists, 9 on a grid it wraps around);
//                                every tile's list is then sorted by source index (cub segmented sort), so the
//                                summation order -- and the result, bit for bit -- does not depend on the order in
//                                which the fill pass's atomic cursors were served
//   t3_records_kernel            the w kernel samples per dimension and the first cells of every live source, once
//                                per batch geometry
//   t3_col_spread_kernel         one CTA per (tile, frequency, product), one thread per column:
//                                walk the tile's list; a thread whose column is inside the source's
//                                (x, y) footprint adds W k_x k_y k_z[.] to its register column; at the
//                                end every column is stored once (plain stores: no memset, no atomics)
#pragma once
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <type_traits>
#include <cuda_pipeline.h>

namespace fv {

constexpr int T3_TILE = 16;     // columns per tile side (256 threads = one column each)
constexpr int T3_NZMAX = 32;    // z cells a thread can hold in registers
constexpr int T3_RS = 64;       // sources staged per round (each round has one phase of dependent loads: list -> source -> record rows)

template <typename T>
struct T3Geom {
  const T* x; const T* y; const T* z;
  const int32_t* n_dev;
  double C[3], invgam[3];
  int nf[3];
  int w;
  int ntx, nty;
};

template <typename T>
__device__ __forceinline__ void t3_fold(const T3Geom<T>& g, int d, T v, int* i0w, T* z0) {
  const double xr = ((double)v - g.C[d]) * g.invgam[d];
  const double gg = fold_grid(xr, g.nf[d]);
  const double gi = ceil(gg - 0.5 * (double)g.w);
  int i0 = (int)gi;
  if (i0 < 0) i0 += g.nf[d];
  *i0w = i0;
  *z0 = (T)(gi - gg);
}

// distinct tiles a w-cell footprint starting at i0w (wrapped) touches along one dimension
__device__ __forceinline__ int t3_tiles_1d(int i0w, int w, int nf, int* out) {
  int cnt = 0;
  for (int j = 0; j < w; ++j) {
    int c = i0w + j;
    if (c >= nf) c -= nf;
    const int t = c / T3_TILE;
    // a footprint that wraps around a grid of fewer than two tiles re-enters its first tile: compare with
    // every tile emitted so far, not only the last one
    bool seen = false;
    for (int q = 0; q < cnt; ++q) seen |= out[q] == t;
    if (!seen) out[cnt++] = t;
  }
  return cnt;
}

// MODE 0: count list lengths; MODE 1: fill the lists (offsets from the exclusive scan of the counts)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
t3_bin_kernel(T3Geom<T> g, int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
              int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  const int n = *g.n_dev;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int ix, iy;
  T zx, zy;
  t3_fold<T>(g, 0, g.x[s], &ix, &zx);
  t3_fold<T>(g, 1, g.y[s], &iy, &zy);
  int tx[4], ty[4];
  const int nx = t3_tiles_1d(ix, g.w, g.nf[0], tx), ny = t3_tiles_1d(iy, g.w, g.nf[1], ty);
  for (int b = 0; b < ny; ++b)
    for (int a = 0; a < nx; ++a) {
      const int tile = ty[b] * g.ntx + tx[a];
      if (MODE == 0) atomicAdd(&counts[tile], 1);
      else list[offsets[tile] + atomicAdd(&cursor[tile], 1)] = s;
    }
}

// Kernel rows of every live source, evaluated ONCE per batch geometry (a source sits in ~3 tile lists and the
// lists are walked again for every frequency and product): record s = three rows of WMAX samples (x, y, z) and
// the three first cells.  One thread per (source, dimension).
template <typename T, int WT>
__global__ void __launch_bounds__(256)
t3_records_kernel(T3Geom<T> g, T beta, T c, T halfw, T* __restrict__ rows, int32_t* __restrict__ cells) {
  const int w = WT > 0 ? WT : g.w;
  constexpr int WMAX = WT > 0 ? WT : kMaxW;
  const int n = *g.n_dev;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = t / 3;
  if (s >= n) return;
  const int d = (int)(t - 3 * s);
  const T v = d == 0 ? g.x[s] : (d == 1 ? g.y[s] : g.z[s]);
  int i0; T z0;
  t3_fold<T>(g, d, v, &i0, &z0);
  cells[3 * s + d] = i0;
  T* r = rows + (3 * s + d) * WMAX;
#pragma unroll
  for (int j = 0; j < WMAX; ++j) r[j] = j < w ? es_kernel<T>(z0 + (T)j, beta, c, halfw) : T(0);
}

// dynamic shared memory of t3_col_spread_kernel
template <typename T>
inline size_t t3_spread_smem(int wmax) {
  return sizeof(T) * (2 * 3 * T3_RS * (size_t)wmax + (size_t)T3_RS * T3_NZMAX) + sizeof(cplx_t<T>) * 2 * T3_RS +
         sizeof(int) * 2 * 3 * T3_RS;
}

template <typename T>
struct T3SpreadArgs {
  T3Geom<T> g;
  int64_t n_cap;
  T beta, c, halfw;
  int ntr, prephase;
  const cplx_t<T>* W;           // (nb, ntr, n_cap)
  const BatchParams* bp;        // D per frequency (pre-phase)
  const int32_t* offsets;       // (ntiles + 1)
  const int32_t* list;
  const T* rows;                // (n_cap, 3, WMAX) kernel rows from t3_records_kernel
  const int32_t* cells;         // (n_cap, 3) first cells
  cplx_t<T>* grid;              // (nb, ntr, nf2, nf1, nf0)
};

template <typename T, int WT>
__global__ void __launch_bounds__(T3_TILE * T3_TILE)
t3_col_spread_kernel(T3SpreadArgs<T> a) {
  using C = cplx_t<T>;
  const int w = WT > 0 ? WT : a.g.w;
  constexpr int WMAX = WT > 0 ? WT : kMaxW;
  // two slots of staged sources: kernel rows (x, y, z), first cells, strengths; slot (i + 1) & 1 fills -- the rows by
  // asynchronous copies (LDGSTS) -- while the z loops of round i run on slot i & 1
  extern __shared__ __align__(16) unsigned char t3s_smem[];
  T (*s_k)[3][T3_RS][WMAX] = reinterpret_cast<T (*)[3][T3_RS][WMAX]>(t3s_smem);                 // [2]
  T (*s_kzr)[T3_NZMAX] = reinterpret_cast<T (*)[T3_NZMAX]>(t3s_smem + sizeof(T) * 2 * 3 * T3_RS * WMAX);   // [T3_RS]: z row
                                                      // rotated onto the grid: value for cell z, 0 outside
  C (*s_w)[T3_RS] = reinterpret_cast<C (*)[T3_RS]>(reinterpret_cast<unsigned char*>(s_kzr) + sizeof(T) * T3_RS * T3_NZMAX);
  int (*s_i)[3][T3_RS] = reinterpret_cast<int (*)[3][T3_RS]>(reinterpret_cast<unsigned char*>(s_w) + sizeof(C) * 2 * T3_RS);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, ty = tile / a.g.ntx, tx = tile - ty * a.g.ntx;
  const int bpi = blockIdx.y, b = bpi / a.ntr;
  const int nf0 = a.g.nf[0], nf1 = a.g.nf[1], nf2 = a.g.nf[2];
  // a warp owns a 4 x 8 patch of columns (not two 16-column rows): a w x w footprint then covers whole warps or
  // none of them more often, so fewer warps run the z loop with most of their lanes switched off
  const int wrp = tid >> 5, ln = tid & 31;
  const int gx = tx * T3_TILE + 8 * (wrp & 1) + (ln & 7), gy = ty * T3_TILE + 4 * (wrp >> 1) + (ln >> 3);
  const bool owner = gx < nf0 && gy < nf1;
  C acc[T3_NZMAX];
#pragma unroll
  for (int z = 0; z < T3_NZMAX; ++z) acc[z] = make_c<T>(T(0), T(0));
  const int l0 = a.offsets[tile], l1 = a.offsets[tile + 1];
  const C* Wp = a.W + (int64_t)bpi * a.n_cap;
  const BatchParams bpar = a.bp[b];
  // The per-source scalars of a round (list entry -> first cells, strength, coordinates) are fetched into registers
  // one round before they are staged, and staged one round before they are used: the chain of dependent global
  // loads list -> source -> {cells, strength, rows} runs under the z loops of the two rounds before.
  int p_s = -1, p_c[3] = {0, 0, 0};
  C p_w = make_c<T>(T(0), T(0));
  T p_x[3] = {T(0), T(0), T(0)};
  auto prefetch = [&](int r0) {
    p_s = -1;
    if (tid < T3_RS && r0 + tid < l1) {
      p_s = a.list[r0 + tid];
      p_c[0] = a.cells[3 * p_s]; p_c[1] = a.cells[3 * p_s + 1]; p_c[2] = a.cells[3 * p_s + 2];
      p_w = Wp[p_s];
      if (a.prephase) { p_x[0] = a.g.x[p_s]; p_x[1] = a.g.y[p_s]; p_x[2] = a.g.z[p_s]; }
    }
  };
  // stage the prefetched round into `slot`: scalars by plain stores, this thread's source's rows by async copies
  auto stage = [&](int slot) {
    if (p_s >= 0) {
      s_i[slot][0][tid] = p_c[0]; s_i[slot][1][tid] = p_c[1]; s_i[slot][2][tid] = p_c[2];
      C cw = p_w;
      if (a.prephase) {
        double sn, cs;
        sincos(bpar.D[0] * (double)p_x[0] + bpar.D[1] * (double)p_x[1] + bpar.D[2] * (double)p_x[2], &sn, &cs);
        cw = cmul(cw, make_c<T>((T)cs, (T)sn));
      }
      s_w[slot][tid] = cw;
      const T* src = a.rows + (int64_t)p_s * 3 * WMAX;
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int j = 0; j < WMAX; ++j) __pipeline_memcpy_async(&s_k[slot][d][tid][j], src + d * WMAX + j, sizeof(T));
    }
    __pipeline_commit();
  };
  prefetch(l0);
  stage(0);
  prefetch(l0 + T3_RS);
  int slot = 0;
  for (int r0 = l0; r0 < l1; r0 += T3_RS, slot ^= 1) {
    const int rn = min(T3_RS, l1 - r0);
    __pipeline_wait_prior(0);
    __syncthreads();                                       // slot's rows and scalars are in; the other slot is free
    // rotate the z rows onto the grid cells so that the inner loop is unconditional
    for (int e = tid; e < rn * T3_NZMAX; e += T3_TILE * T3_TILE) {
      const int rr = e / T3_NZMAX, z = e - rr * T3_NZMAX;
      int jz = z - s_i[slot][2][rr]; if (jz < 0) jz += nf2;
      s_kzr[rr][z] = (z < nf2 && jz < w) ? s_k[slot][2][rr][jz] : T(0);
    }
    stage(slot ^ 1);                                       // the round after this one (in registers since the last round)
    prefetch(r0 + 2 * T3_RS);                              // and the one after that
    __syncthreads();
    for (int r = 0; r < rn; ++r) {
      int jx = gx - s_i[slot][0][r]; if (jx < 0) jx += nf0;
      int jy = gy - s_i[slot][1][r]; if (jy < 0) jy += nf1;
      if (owner && jx < w && jy < w) {
        const T kxy = s_k[slot][0][r][jx] * s_k[slot][1][r][jy];
        const C cw = s_w[slot][r];
        const T cr = cw.x * kxy, ci = cw.y * kxy;
        // two z cells per shared-memory load (the row is 16-byte aligned and read as a broadcast)
        using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
        const T2* kz2 = reinterpret_cast<const T2*>(&s_kzr[r][0]);
#pragma unroll
        for (int z = 0; z < T3_NZMAX; z += 2) {
          const T2 k = kz2[z >> 1];
          acc[z].x += cr * k.x; acc[z].y += ci * k.x;
          acc[z + 1].x += cr * k.y; acc[z + 1].y += ci * k.y;
        }
      }
    }
    __syncthreads();                                       // s_kzr is rewritten by the next round
  }
  if (owner) {
    C* gp = a.grid + (int64_t)bpi * nf2 * nf1 * nf0 + (int64_t)gy * nf0 + gx;
#pragma unroll
    for (int z = 0; z < T3_NZMAX; ++z)
      if (z < nf2) gp[(int64_t)z * nf1 * nf0] = acc[z];
  }
}

}  // namespace fv
