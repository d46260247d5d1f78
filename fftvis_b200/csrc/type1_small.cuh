// Type-1 transform on SMALL fine grids (the whole nf x nf grid of one transform fits the shared memory of
// one CTA: HERA-331-like arrays, nf = 90) with MANY sources per cell -- BASELINE configs[2] and [4]
// (100k sources, w = 14, four polarisation products, K (K + 1) / 2 basis pairs).
//
// The strip kernel of type1_fused.cuh spreads a hit by read-modify-write of its w x w cells in shared
// memory: 196 cells x 32 B per hit and product, i.e. the shared-memory pipe is the bound and every
// product re-evaluates the 2 w kernel samples.  Here a hit costs FMAs on REGISTERS instead:
//
//   t1s_key_kernel      key (frequency, bin) of every live source: bins are B x B blocks of footprint
//                       origins, so all footprints of a bin lie inside one (w + B - 1)^2 window, B chosen so
//                       that the window fits the 4T x 4T cells a half-warp holds (w = 14: B = 3, 16 x 16)
//   cub radix sort      stable: sources of a bin stay in catalogue order -> deterministic sums
//   t1s_bounds_kernel   first sorted position of every (frequency, bin)
//   t1s_records_kernel  kernel samples of every sorted (frequency, source), evaluated ONCE for all products /
//                       basis pairs and already shifted to window coordinates (zero outside the footprint)
//   t1s_spread_kernel   one CTA per (frequency, transform).  A warp takes a bin; each HALF-warp keeps the
//                       bin's whole window in registers (16 lanes x T x T cells) and takes every other record
//                       of the bin: acc[j][i] += (W kx[i]) ky[j] -- an outer product per lane, so that a record
//                       costs 2 T reals of shared-memory traffic per lane for T^2 complex FMAs (the
//                       shared-memory pipe, not the FP64 pipe, bounds this kernel).  The two halves exchange
//                       half of their partial sums by shuffles and add the window to the shared-memory grid once.
//                       Bins are processed in phases of pairwise disjoint windows (host-built schedule), so
//                       the adds need no atomics and the order of summation is fixed.  Then the row FFTs
//                       and the write-out of the needed columns, as in t1_spread_fftx_kernel.
//
// Same kernel, width, grid size and deconvolution as the strip path (finufft.nufft2d1, reference
// cpu/nufft.py:120-175): only the order of the additions differs.
#pragma once

namespace fv {

template <int WT>
struct T1Small {
  static constexpr int T = (WT + 4) / 4;                 // cells per lane and dimension: ceil((w + 1) / 4)
  static constexpr int SMAX = 4 * T;                     // widest window a half-warp (4 x 4 lanes) holds
  static constexpr int BMAX = SMAX - WT + 1;             // largest bin size (origins per dimension), >= 2
  static constexpr int REC = 8 * T;                      // reals per record: kx[4T], ky[4T] in window coordinates
  static_assert(T >= 2 && T <= 4, "kernel width");
};

// bin size for a grid: the largest B <= BMAX (and <= 6) that divides nf
inline int t1s_bin_size(int64_t nf, int w) {
  const int bmax = std::min(4 * ((w + 4) / 4) - w + 1, 6);
  for (int B = bmax; B >= 2; --B) if (nf % B == 0) return B;
  return 0;
}

template <typename T>
struct T1SmallArgs {
  const int32_t* n_dev;
  int64_t n_cap;
  int nf, pitch, w, nbd, B;      // B x B origins per bin, nbd = nf / B bins per dimension
  T beta, c, halfw;
  int ntr;
  const cplx_t<T>* W;            // (nb, ntr, n_cap)
  const int32_t* ix0; const int32_t* iy0; const T* zx; const T* zy;   // t1_prep_kernel, (nb, n_cap)
  uint32_t* keys; int32_t* vals;          // unsorted (nb * n_cap)
  const uint32_t* skeys; const int32_t* svals;   // sorted
  int32_t* off;                  // (nb * nbins + 1) first sorted position of every (frequency, bin)
  T* rec;                        // (nb * n_cap, REC) records in sorted order
  int nphase;
  const int32_t* ph_off;         // (nphase + 1)
  const uint16_t* ph_bins;       // bins of every phase
  const cplx_t<T>* tw;
  FftStages st;
  int ncols;
  const int32_t* col_pos;
  cplx_t<T>* Tbuf;               // (nb, ntr, ncols, nf)
  int nb;
};

template <typename T> struct t1s_vec;
template <> struct t1s_vec<double> { using type = double2; };
template <> struct t1s_vec<float> { using type = float2; };

template <typename T>
__global__ void __launch_bounds__(256)
t1s_key_kernel(T1SmallArgs<T> a) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_cap) return;
  const int b = blockIdx.y;
  const int64_t o = (int64_t)b * a.n_cap + s;
  const int nbins = a.nbd * a.nbd;
  uint32_t key = (uint32_t)a.nb * (uint32_t)nbins;      // sentinel: sorts behind every live source
  if (s < *a.n_dev) {
    const int cx = wrap_idx(a.ix0[o], a.nf), cy = wrap_idx(a.iy0[o], a.nf);
    key = (uint32_t)b * (uint32_t)nbins + (uint32_t)((cy / a.B) * a.nbd + (cx / a.B));
  }
  a.keys[o] = key;
  a.vals[o] = (int32_t)s;
}

// off[k] = first sorted position whose key is >= k, k = 0 .. nkeys (nkeys = nb * nbins = the sentinel)
__global__ void __launch_bounds__(256)
t1s_bounds_kernel(const uint32_t* __restrict__ skeys, int64_t n, uint32_t nkeys, int32_t* __restrict__ off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint32_t cur = i < n ? min(skeys[i], nkeys) : nkeys;
  if (i == 0) { for (uint32_t k = 0; k <= cur; ++k) off[k] = 0; return; }
  const uint32_t prev = min(skeys[i - 1], nkeys);
  for (uint32_t k = prev + 1; k <= cur; ++k) off[k] = (int32_t)i;
  if (i == n) for (uint32_t k = cur + 1; k <= nkeys; ++k) off[k] = (int32_t)n;
}

// one thread per (live sorted item, dimension): its w kernel samples, shifted to window coordinates
template <typename T, int WT>
__global__ void __launch_bounds__(256)
t1s_records_kernel(T1SmallArgs<T> a, int64_t nitems) {
  using G = T1Small<WT>;
  const int nbins = a.nbd * a.nbd;
  const int64_t nlive = a.off[(int64_t)a.nb * nbins];           // sorted position of the first sentinel key
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 1;
  if (i >= nlive) return;
  const int dim = (int)(t & 1);
  const uint32_t key = a.skeys[i];
  const int b = (int)(key / (uint32_t)nbins), bin = (int)(key - (uint32_t)b * nbins);
  const int by = bin / a.nbd, bx = bin - by * a.nbd;
  const int64_t o = (int64_t)b * a.n_cap + a.svals[i];
  const int o0 = wrap_idx(dim == 0 ? a.ix0[o] : a.iy0[o], a.nf) - a.B * (dim == 0 ? bx : by);   // in [0, B)
  const T z0 = dim == 0 ? a.zx[o] : a.zy[o];
  T* out = a.rec + i * G::REC + dim * G::SMAX;
  // window slot sl sees kernel sample sl - o0 (zero outside [0, w)); two slots per store
#pragma unroll
  for (int sl = 0; sl < G::SMAX; sl += 2) {
    const int j0 = sl - o0, j1 = j0 + 1;
    const T v0 = (j0 >= 0 && j0 < WT) ? es_kernel<T>(z0 + (T)j0, a.beta, a.c, a.halfw) : T(0);
    const T v1 = (j1 >= 0 && j1 < WT) ? es_kernel<T>(z0 + (T)j1, a.beta, a.c, a.halfw) : T(0);
    typename t1s_vec<T>::type v; v.x = v0; v.y = v1;
    *reinterpret_cast<typename t1s_vec<T>::type*>(out + sl) = v;
  }
}

// ---- mbarrier + 1-D bulk copy (TMA) helpers: a warp streams the records of its bins through a private
// double buffer in shared memory; the copy of chunk k + 1 is in flight while chunk k is consumed --------------
constexpr int T1S_CH = 8;          // records per chunk of the per-warp pipeline

// shared memory of one CTA of t1s_spread_kernel
template <typename T>
inline size_t t1s_smem_bytes(int nf, int nbd, int nphase, int nwarps, int rec_len) {
  const size_t grid = sizeof(cplx_t<T>) * ((size_t)nf * (nf + 1) + nf);            // strip + twiddles
  size_t ints = sizeof(int) * ((size_t)nf + 2 * nphase + 4);                        // colp, done / need counters, item counter
  ints += sizeof(int) * ((size_t)nbd * nbd + 2 + nphase + 1) + sizeof(uint16_t) * (((size_t)nbd * nbd + 1) & ~(size_t)1);
  const size_t per_warp = 2 * ((size_t)T1S_CH * rec_len * sizeof(T) + T1S_CH * sizeof(cplx_t<T>)) + 16;
  return ((grid + ints + 15) & ~(size_t)15) + nwarps * per_warp + 16;
}

template <typename T, int WT>
__global__ void __launch_bounds__(512)
t1s_spread_kernel(T1SmallArgs<T> a) {
  using C = cplx_t<T>;
  using G = T1Small<WT>;
  constexpr int CH = T1S_CH;
  extern __shared__ __align__(16) unsigned char t1s_smem[];
  const int nf = a.nf, pitch = a.pitch;
  C* strip = (C*)t1s_smem;                         // nf * pitch
  C* tw = strip + (size_t)nf * pitch;              // nf
  int* colp = (int*)(tw + nf);                     // nf
  int* done = colp + nf;                           // bins of every phase already added to the grid
  int* need = done + a.nphase;                     // bins of every phase
  int* next_item = need + a.nphase;                // work counter over the schedule (phase-major)
  const int nbins = a.nbd * a.nbd;
  int* s_off = next_item + 4;                      // nbins + 1 (+ 1 pad): first sorted record of every bin of this frequency
  int* s_phoff = s_off + nbins + 2;                // nphase + 1
  uint16_t* s_bins = (uint16_t*)(s_phoff + a.nphase + 1);   // schedule: bins in phase-major order
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  size_t fixed = sizeof(C) * ((size_t)nf * pitch + nf) + sizeof(int) * ((size_t)nf + 2 * a.nphase + 4) +
                 sizeof(int) * ((size_t)nbins + 2 + a.nphase + 1) + sizeof(uint16_t) * (((size_t)nbins + 1) & ~(size_t)1);
  fixed = (fixed + 15) & ~(size_t)15;
  constexpr size_t kRecBytes = (size_t)CH * G::REC * sizeof(T);
  constexpr size_t kPerWarp = 2 * (kRecBytes + CH * sizeof(C)) + 16;
  unsigned char* mine = t1s_smem + fixed + (size_t)warp * kPerWarp;
  T* rbuf = (T*)mine;                              // [2][CH * REC]
  C* wbuf = (C*)(mine + 2 * kRecBytes);            // [2][CH]
  const unsigned bar0 = smem_addr(mine + 2 * (kRecBytes + CH * sizeof(C)));   // two mbarriers
  const int bpi = blockIdx.x;                      // (frequency, transform)
  const int b = bpi / a.ntr;
  const int nitems = a.ph_off[a.nphase];
  {
    const int32_t* goff = a.off + (int64_t)b * nbins;
    for (int i = tid; i <= nbins; i += nthr) s_off[i] = goff[i];
    for (int i = tid; i <= a.nphase; i += nthr) s_phoff[i] = a.ph_off[i];
    for (int i = tid; i < nitems; i += nthr) s_bins[i] = a.ph_bins[i];
  }

  for (int i = tid; i < nf * pitch; i += nthr) strip[i] = make_c<T>(T(0), T(0));
  for (int i = tid; i < a.st.tw_len; i += nthr) tw[i] = a.tw[i];
  for (int i = tid; i < a.ncols; i += nthr) colp[i] = a.col_pos[i];
  for (int i = tid; i < a.nphase; i += nthr) { done[i] = 0; need[i] = a.ph_off[i + 1] - a.ph_off[i]; }
  if (tid == 0) *next_item = 0;
  if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const C* Wp = a.W + (int64_t)bpi * a.n_cap;
  const int* off = s_off;
  constexpr int TT = G::T;
  const int half = lane >> 4, hl = lane & 15;
  const int lx = hl & 3, ly = hl >> 2;             // lane tile: window columns lx T .., rows ly T ..
  const int kx_slot = lx * TT, ky_slot = G::SMAX + ly * TT;
  const int s_eff = a.w + a.B - 1;                 // true window side (<= 4 T); cells beyond it belong to nobody

  // fetch cursor (warp-uniform): the item being fetched and the next record of it to request
  int f_i = 0, f_i1 = 0, f_bin = 0, f_ph = 0;
  bool exhausted = false;
  // meta of a chunk: bin, phase, record count, last chunk of its bin
  struct Chunk { int bin, ph, n, last; };
  unsigned parity = 0;                              // bit s = parity to wait for on stage s
  C w_next = make_c<T>(T(0), T(0));

  auto fetch = [&](int stage, Chunk& ck) -> bool {
    while (!exhausted && f_i >= f_i1) {
      int j = 0;
      if (lane == 0) j = atomicAdd(next_item, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= nitems) { exhausted = true; break; }
      f_bin = s_bins[j];
      // phase of item j: the schedule is phase-major, phases are short runs: walk forward from the last one
      while (s_phoff[f_ph + 1] <= j) ++f_ph;
      f_i = off[f_bin]; f_i1 = off[f_bin + 1];
      if (f_i >= f_i1 && lane == 0) atomicAdd(&done[f_ph], 1);      // empty bin: nothing to add to the grid
    }
    if (exhausted) return false;
    const int n = min(CH, f_i1 - f_i);
    ck.bin = f_bin; ck.ph = f_ph; ck.n = n; ck.last = (f_i + n >= f_i1);
    if (lane == 0) {
      const unsigned bytes = (unsigned)(n * G::REC * sizeof(T));
      const unsigned bar = bar0 + 8u * stage;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_addr(rbuf + (size_t)stage * CH * G::REC), a.rec + (int64_t)f_i * G::REC, bytes, bar);
    }
    // strengths of the chunk's sources: one gathered load per lane, consumed one chunk later
    w_next = lane < n ? Wp[a.svals[f_i + lane]] : make_c<T>(T(0), T(0));
    f_i += n;
    return true;
  };

  C acc[TT][TT];                                    // [row j][column i]
#pragma unroll
  for (int j = 0; j < TT; ++j)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[j][i] = make_c<T>(T(0), T(0));
  Chunk cur{}, nxt{};
  int stage = 0;
  bool have = fetch(0, nxt);
  while (have) {
    cur = nxt;
    const C w_cur = w_next;
    __syncwarp();                                   // every lane is done with the other stage's buffers
    have = fetch(stage ^ 1, nxt);
    // this chunk: strengths to shared memory (broadcast reads below), records have landed?
    C* wb = wbuf + stage * CH;
    if (lane < CH) wb[lane] = w_cur;
    mbar_wait(bar0 + 8u * stage, (parity >> stage) & 1u);
    parity ^= 1u << stage;
    __syncwarp();
    const T* rb = rbuf + (size_t)stage * CH * G::REC;
    // half-warp `half` takes records half, half + 2, ...
    for (int r0 = half; r0 < cur.n; r0 += 2) {
      const T* rr = rb + r0 * G::REC;
      const C wv = wb[r0];
      T kx[TT], ky[TT];
      if (TT % 2 == 0) {
        using V = typename t1s_vec<T>::type;
#pragma unroll
        for (int i = 0; i < TT; i += 2) {
          const V u = *reinterpret_cast<const V*>(rr + kx_slot + i), v = *reinterpret_cast<const V*>(rr + ky_slot + i);
          kx[i] = u.x; kx[i + 1 < TT ? i + 1 : i] = u.y; ky[i] = v.x; ky[i + 1 < TT ? i + 1 : i] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TT; ++i) { kx[i] = rr[kx_slot + i]; ky[i] = rr[ky_slot + i]; }
      }
      C e[TT];
#pragma unroll
      for (int i = 0; i < TT; ++i) e[i] = make_c<T>(wv.x * kx[i], wv.y * kx[i]);
#pragma unroll
      for (int j = 0; j < TT; ++j)
#pragma unroll
        for (int i = 0; i < TT; ++i) { acc[j][i].x += e[i].x * ky[j]; acc[j][i].y += e[i].y * ky[j]; }
    }
    if (cur.last) {
      // The two halves hold partial sums of the same window (even / odd records).  Half 0 keeps tile rows
      // [0, KH), half 1 rows [KH, T): each sends the other the rows it does not keep (one xor-16 shuffle per
      // word moves both directions), then adds its rows to the grid -- once every window of the previous phase
      // is in (windows of one phase are pairwise disjoint, a lane's cells are its own): no atomics on the
      // grid, fixed order of summation.
      constexpr int KH = (TT + 1) / 2;
      C tot[KH][TT];
#pragma unroll
      for (int j = 0; j < KH; ++j)
#pragma unroll
        for (int i = 0; i < TT; ++i) {
          const C keep = half ? (KH + j < TT ? acc[KH + j < TT ? KH + j : 0][i] : make_c<T>(T(0), T(0))) : acc[j][i];
          const C send = half ? acc[j][i] : (KH + j < TT ? acc[KH + j < TT ? KH + j : 0][i] : make_c<T>(T(0), T(0)));
          C got;
          got.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
          got.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
          tot[j][i] = make_c<T>(keep.x + got.x, keep.y + got.y);
        }
      if (cur.ph > 0) {
        const volatile int* dp = done + (cur.ph - 1);
        const int want = need[cur.ph - 1];
        while (*dp < want) __nanosleep(64);
        __threadfence_block();
      }
      {
        const int by = cur.bin / a.nbd, bx = cur.bin - by * a.nbd;
#pragma unroll
        for (int j = 0; j < KH; ++j) {
          const int wr = ly * TT + (half ? KH + j : j);            // window row of tot[j]
          const bool row_ok = (half ? KH + j : j) < TT && wr < s_eff;
          int row = a.B * by + wr;
          if (row >= nf) row -= nf;
#pragma unroll
          for (int i = 0; i < TT; ++i) {
            const int wc = lx * TT + i;
            if (row_ok && wc < s_eff) {
              int col = a.B * bx + wc;
              if (col >= nf) col -= nf;
              C* cell = strip + row * pitch + col;
              C v = *cell;
              v.x += tot[j][i].x; v.y += tot[j][i].y;
              *cell = v;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < TT; ++j)
#pragma unroll
          for (int i = 0; i < TT; ++i) acc[j][i] = make_c<T>(T(0), T(0));
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) atomicAdd(&done[cur.ph], 1);
    }
    stage ^= 1;
  }
  __syncthreads();

  smem_fft<T>(strip, nf, pitch, nf, tw, a.st);
  __syncthreads();

  // needed columns -> T[col][row] (rows contiguous)
  const int total = a.ncols * nf;
  const unsigned inv_rows = 0xFFFFFFFFu / (unsigned)nf + 1u;
  C* Tb = a.Tbuf + (int64_t)bpi * a.ncols * nf;
#pragma unroll 4
  for (int i = tid; i < total; i += nthr) {
    const int ci = (int)__umulhi((unsigned)i, inv_rows);
    const int rr = i - ci * nf;
    Tb[(int64_t)ci * nf + rr] = strip[rr * pitch + colp[ci]];
  }
}

}  // namespace fv
