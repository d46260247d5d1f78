// Beam-table interpolation with the table staged in shared memory (stages a4 + a5 for az/za table beams).
//
// weights_kernel gathers the 4 taps x 4 Jones entries of every (source, frequency) straight from global
// memory: ~16 sectors per evaluation, and the same table cells are fetched again by every source near them.
// Here the live sources of a time step are sorted ONCE by the 16 x 16-cell tile of the beam grid their
// direction falls in (the directions do not depend on frequency), and one CTA per (tile, group of
// frequencies) stages the tile's 17 x 17-point patch of every Jones entry in shared memory -- 1-D bulk copies
// (TMA, cp.async.bulk + mbarrier) of the patch rows, double-buffered over the frequencies -- for all of the
// tile's sources: the table is read once per (frequency, tile) instead of once per source.
//
//   tile_key_kernel      tile of every live slot (same index arithmetic as eval_beam), sentinel for dead slots
//   cub radix sort       stable: slots of a tile stay in ascending order
//   tile_permute_kernel  the live set (xyz, az, za, src_idx) re-ordered by tile, so that slot == sorted position and
//                        every later access by slot (strengths, NUFFT) stays coalesced
//   weights_tiled_kernel the strengths: pair form (fv_weights) and basis form (fv_weights_basis)
#pragma once

namespace fv {

constexpr int WT_TILE = 16;                  // grid cells per tile side
constexpr int WT_PTS = WT_TILE + 1;          // patch points per side (bilinear taps reach one cell further)
constexpr int WT_THREADS = 256;
constexpr int WT_MAXSTAGE = 4;               // ring of patch buffers (as many as shared memory holds)
constexpr int WT_PASSES = 2;                 // sources per thread and staging (tiles of up to 512 sources: one staging per frequency)

struct TileGrid { int ntz, nta; };
__host__ __device__ inline TileGrid tile_grid(const fv_beam& b) {
  TileGrid g;
  g.ntz = (b.nza + WT_TILE - 1) / WT_TILE; g.nta = (b.naz + WT_TILE - 1) / WT_TILE;
  return g;
}

template <typename T>
__global__ void __launch_bounds__(256)
tile_key_kernel(fv_beam b, const T* __restrict__ az, const T* __restrict__ za, const int32_t* __restrict__ n_dev,
                int64_t n_cap, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cap) return;
  const TileGrid g = tile_grid(b);
  uint32_t key = (uint32_t)(g.ntz * g.nta);              // dead slots sort to the end
  if (s < *n_dev) {
    double zi, ai, wz, wa;
    int z0, z1, a0, a1;
    table_index(b, (double)az[s], (double)za[s], zi, ai);
    table_taps(b, zi, ai, z0, z1, a0, a1, wz, wa);
    // the tile of the lower tap; the upper tap is at most one cell further: inside the 17-point patch
    key = (uint32_t)((min(z0, z1) / WT_TILE) * g.nta + min(a0, a1) / WT_TILE);
  }
  keys[s] = key; vals[s] = (int32_t)s;
}

template <typename T>
__global__ void __launch_bounds__(256)
tile_permute_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ n_dev, int64_t n_cap,
                    const T* __restrict__ xyz, const T* __restrict__ az, const T* __restrict__ za,
                    const int32_t* __restrict__ src_idx, T* __restrict__ o_xyz, T* __restrict__ o_az,
                    T* __restrict__ o_za, int32_t* __restrict__ o_src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  const int64_t s = perm[i];
#pragma unroll
  for (int d = 0; d < 3; ++d) o_xyz[d * n_cap + i] = xyz[d * n_cap + s];
  o_az[i] = az[s]; o_za[i] = za[s]; o_src[i] = src_idx[s];
}

template <typename T>
struct WtArgs {
  int mode;                      // 0 / 1 / 2 as fv_weights
  int K;                         // beams staged: pair form 1 (same beam) or 2; basis form K
  int basis;                     // 1: all pairs k <= l (fv_weights_basis); 0: the pair (0, K - 1)
  fv_beam b[8];                  // same grid geometry and order for all (checked by the host)
  const int32_t* tile_off;       // (ntiles + 1) slot ranges
  const T* az; const T* za; const int32_t* src_idx;
  int64_t n_cap;
  const double* freqs; int64_t f0; int nf, freqs_per_cta;
  const cplx_t<T>* flux; int64_t nsrc_total;
  cplx_t<T>* out;
  int bulk;                      // 1: patch rows by cp.async.bulk (16-byte table entries); 0: plain loads
  int nstage;                    // ring slots
};

// off[k] = first sorted position whose key is >= k, k = 0 .. nkeys (nkeys = the sentinel of dead slots)
__global__ void __launch_bounds__(256)
tile_bounds_kernel(const uint32_t* __restrict__ skeys, int64_t n, uint32_t nkeys, int32_t* __restrict__ off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint32_t cur = i < n ? min(skeys[i], nkeys) : nkeys;
  if (i == 0) { for (uint32_t k = 0; k <= cur; ++k) off[k] = 0; return; }
  const uint32_t prev = min(skeys[i - 1], nkeys);
  for (uint32_t k = prev + 1; k <= cur; ++k) off[k] = (int32_t)i;
}

// E = table entry type (cplx_t<T> for E-field tables, T for power tables), NC = entries per grid point,
// KMAX = most beams staged together (register budget of the Jones matrices)
template <typename T, typename E, int NC, int KMAX>
__global__ void __launch_bounds__(WT_THREADS)
weights_tiled_kernel(WtArgs<T> a) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char wt_smem[];
  __shared__ __align__(8) unsigned long long bars[WT_MAXSTAGE];
  const fv_beam& b0 = a.b[0];
  const TileGrid g = tile_grid(b0);
  const int tile = blockIdx.x, tz = tile / g.nta, ta = tile - tz * g.nta;
  const int s_lo = a.tile_off[tile], s_hi = a.tile_off[tile + 1];
  if (s_lo >= s_hi) return;
  const int fb0 = blockIdx.y * a.freqs_per_cta, fb1 = min(a.nf, fb0 + a.freqs_per_cta);
  if (fb0 >= fb1) return;
  const int z_lo = tz * WT_TILE, a_lo = ta * WT_TILE;
  const int nrows = min(WT_PTS, b0.nza - z_lo), ncols = min(WT_PTS, b0.naz - a_lo);
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)b0.nza * b0.naz;
  const int patch = NC * WT_PTS * WT_PTS;                  // entries per beam and stage
  const int nstage = a.nstage;
  E* buf = reinterpret_cast<E*>(wt_smem);                  // [nstage][K][NC][WT_PTS][WT_PTS]
  const unsigned bar0 = smem_u32(&bars[0]);
  if (a.bulk && tid == 0) { for (int i = 0; i < nstage; ++i) mbar_init(bar0 + 8 * i, 1); }
  if (a.bulk) { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();

  // stage the patches of frequency fb into ring slot `st`
  auto stage = [&](int fb, int st) {
    E* dst = buf + (size_t)st * a.K * patch;
    const int nrow_all = a.K * NC * nrows;
    if (a.bulk) {
      if (tid < 32) {
        if (tid == 0) mbar_expect_tx(bar0 + 8 * st, (unsigned)(nrow_all * ncols * (int)sizeof(E)));
        __syncwarp();
        for (int r = tid; r < nrow_all; r += 32) {
          const int k = r / (NC * nrows), rem = r - k * NC * nrows, c = rem / nrows, row = rem - c * nrows;
          const E* src = reinterpret_cast<const E*>(a.b[k].table) +
                         ((int64_t)(a.b[k].freq_offset + fb) * NC + c) * plane + (int64_t)(z_lo + row) * b0.naz + a_lo;
          bulk_g2s(smem_u32(dst + (size_t)k * patch + (c * WT_PTS + row) * WT_PTS), src, (unsigned)(ncols * sizeof(E)),
                   bar0 + 8 * st);
        }
      }
    } else {
      for (int e = tid; e < nrow_all * ncols; e += WT_THREADS) {
        const int r = e / ncols, col = e - r * ncols;
        const int k = r / (NC * nrows), rem = r - k * NC * nrows, c = rem / nrows, row = rem - c * nrows;
        const E* src = reinterpret_cast<const E*>(a.b[k].table) +
                       ((int64_t)(a.b[k].freq_offset + fb) * NC + c) * plane + (int64_t)(z_lo + row) * b0.naz + a_lo;
        dst[(size_t)k * patch + (c * WT_PTS + row) * WT_PTS + col] = src[col];
      }
    }
  };

  const int P = a.mode == 0 ? 1 : 4;
  const int npairs = a.basis ? a.K * (a.K + 1) / 2 : 1;
  int uses = 0;                                            // stage() calls so far: slot = uses % nstage
  // the tile's sources in groups of WT_PASSES x 256: a patch is staged once per frequency for the whole group
  for (int p0 = s_lo; p0 < s_hi; p0 += WT_PASSES * WT_THREADS) {
    int off2[WT_PASSES][2];                                // tap offsets inside the patch, two 16-bit halves each
    double wgt[WT_PASSES][4];
    int srcs[WT_PASSES];
#pragma unroll
    for (int q = 0; q < WT_PASSES; ++q) {
      const int s = p0 + q * WT_THREADS + tid;
      off2[q][0] = off2[q][1] = 0; srcs[q] = -1;
      wgt[q][0] = wgt[q][1] = wgt[q][2] = wgt[q][3] = 0.0;
      if (s < s_hi) {
        double zi, ai, wz, wa;
        int z0, z1, a0, a1;
        table_index(b0, (double)a.az[s], (double)a.za[s], zi, ai);
        table_taps(b0, zi, ai, z0, z1, a0, a1, wz, wa);
        off2[q][0] = ((z0 - z_lo) * WT_PTS + (a0 - a_lo)) | (((z0 - z_lo) * WT_PTS + (a1 - a_lo)) << 16);
        off2[q][1] = ((z1 - z_lo) * WT_PTS + (a0 - a_lo)) | (((z1 - z_lo) * WT_PTS + (a1 - a_lo)) << 16);
        wgt[q][0] = (1 - wz) * (1 - wa); wgt[q][1] = (1 - wz) * wa; wgt[q][2] = wz * (1 - wa); wgt[q][3] = wz * wa;
        srcs[q] = a.src_idx[s];
      }
    }
    const int u0 = uses;
    for (int i = 0; i < nstage - 1 && fb0 + i < fb1; ++i) { stage(fb0 + i, uses % nstage); ++uses; }
    for (int fb = fb0; fb < fb1; ++fb) {
      const int use = u0 + (fb - fb0), st = use % nstage;
      if (fb + nstage - 1 < fb1) { stage(fb + nstage - 1, uses % nstage); ++uses; }
      if (a.bulk) mbar_wait(bar0 + 8 * st, (unsigned)((use / nstage) & 1));
      else __syncthreads();
      const E* pb = buf + (size_t)st * a.K * patch;
      const int64_t fi = a.f0 + fb;
#pragma unroll
      for (int q = 0; q < WT_PASSES; ++q) {
        if (srcs[q] < 0) continue;
        const int s = p0 + q * WT_THREADS + tid;
        const int o00 = off2[q][0] & 0xffff, o01 = off2[q][0] >> 16, o10 = off2[q][1] & 0xffff, o11 = off2[q][1] >> 16;
        const double w00 = wgt[q][0], w01 = wgt[q][1], w10 = wgt[q][2], w11 = wgt[q][3];
        const int64_t src = srcs[q];
        C A[KMAX][NC];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < a.K) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const E* pc = pb + (size_t)k * patch + c * WT_PTS * WT_PTS;
              if constexpr (NC == 1) {
                const double v = w00 * (double)pc[o00] + w01 * (double)pc[o01] + w10 * (double)pc[o10] + w11 * (double)pc[o11];
                A[k][c] = make_c<T>((T)v, T(0));
              } else {
                const E p00 = pc[o00], p01 = pc[o01], p10 = pc[o10], p11 = pc[o11];
                A[k][c] = make_c<T>((T)(w00 * p00.x + w01 * p01.x + w10 * p10.x + w11 * p11.x),
                                    (T)(w00 * p00.y + w01 * p01.y + w10 * p10.y + w11 * p11.y));
              }
            }
          }
        }
        if constexpr (NC == 1) {
          // sqrt(B_i B_j) F with the complex principal root (beams were cast to complex first)
          const T p = A[0][0].x * A[a.K - 1][0].x;
          const C r = p >= T(0) ? make_c<T>(sqrt(p), T(0)) : make_c<T>(T(0), sqrt(-p));
          a.out[(int64_t)fb * a.n_cap + s] = cmul(r, a.flux[fi * a.nsrc_total + src]);
        } else {
          C Cm[4];
          if (a.mode == 1) Cm[0] = a.flux[fi * a.nsrc_total + src];
          else {
#pragma unroll
            for (int c = 0; c < 4; ++c) Cm[c] = a.flux[(fi * 4 + c) * a.nsrc_total + src];
          }
          C* o = a.out + ((int64_t)fb * npairs * P) * a.n_cap + s;
          if (!a.basis) {
            C res[4];
            coherency_product<T>(a.mode, A[0], A[a.K - 1], Cm, res);
#pragma unroll
            for (int c = 0; c < 4; ++c) o[(int64_t)c * a.n_cap] = res[c];
          } else {
            int qq = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
#pragma unroll
              for (int l = k; l < KMAX; ++l) {
                if (l < a.K) {
                  C res[4];
                  coherency_product<T>(a.mode, A[k], A[l], Cm, res);
#pragma unroll
                  for (int c = 0; c < 4; ++c) o[(int64_t)(qq * 4 + c) * a.n_cap] = res[c];
                  ++qq;
                }
              }
          }
        }
      }
      __syncthreads();                                     // the slot is refilled nstage - 1 frequencies later
    }
  }
}

}  // namespace fv
