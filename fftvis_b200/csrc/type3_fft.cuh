// Type-3 inner FFT, pruned: the transform that follows the spreader in finufft's type-3 scheme is a
// zero-padded FFT of the deconvolved spread grid (nf modes per dimension inside an ng = sigma * nf
// grid; reference: finufft.nufft2d3 / nufft3d3 inner type-2 step, cpu/nufft.py:48,105).  A library
// FFT of the padded grid moves ~6 passes of the full ng^d array; here each dimension is transformed by
// a shared-memory pass that READS only the non-zero extent of the dimensions not yet transformed and
// applies the deconvolution on the way in:
//
//   x pass  t3_fft_contig_kernel   rows (s3, s2) of the spread grid -> (nf3, nf2, ng1)   [+ 1/phihat]
//   y pass  t3_fft_strided_kernel  (nf3, nf2, ng1) -> (nf3, ng2, ng1)
//   z pass  t3_fft_strided_kernel  (nf3, ng2, ng1) -> (ng3, ng2, ng1)                   (3-D only)
//
// For cfg4 (1620 x 1620 x 30 -> 3240 x 3240 x 60, complex128) this is 26 GB of traffic per transform
// instead of ~120 GB, and the separate deconvolve + pad kernel disappears.  The FFT itself is the
// shared-memory mixed-radix transform of type1_fused.cuh, here with the whole CTA cooperating on each
// stage (vectors are long and few).
#pragma once

namespace fv {

// one DIF stage of radix R over `nvec` vectors with all threads of the CTA (flat butterfly index)
template <typename T, int R>
__device__ __forceinline__ void fft_stage_flat(cplx_t<T>* data, int nvec, int pitch, int N, int n, unsigned inv,
                                               const cplx_t<T>* __restrict__ tws) {
  using C = cplx_t<T>;
  const int m = n / R, per_vec = N / R, total = nvec * per_vec;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int v = idx / per_vec, t = idx - v * per_vec;
    const int blk = inv ? (int)__umulhi((unsigned)t, inv) : t;
    const int j = t - blk * m;
    C* vec = data + v * pitch + blk * n + j;
    C x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = vec[q * m];
    dft_r<T, R>(x);
    vec[0] = x[0];
    if (m > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) vec[q * m] = cmul(x[q], tws[(q - 1) * m + j]);
    } else {
#pragma unroll
      for (int q = 1; q < R; ++q) vec[q] = x[q];
    }
  }
}

// in-place FFT of `nvec` shared-memory vectors, CTA-cooperative (a barrier after every stage)
template <typename T>
__device__ void smem_fft_cta(cplx_t<T>* data, int nvec, int pitch, int N, const cplx_t<T>* __restrict__ tw,
                             const FftStages& st) {
  int n = N;
  for (int s = 0; s < st.nstage; ++s) {
    const int r = st.radix[s];
    const unsigned inv = st.inv_m[s];
    const cplx_t<T>* tws = tw + st.tw_off[s];
    switch (r) {
      case 15: fft_stage_flat<T, 15>(data, nvec, pitch, N, n, inv, tws); break;
      case 8: fft_stage_flat<T, 8>(data, nvec, pitch, N, n, inv, tws); break;
      case 4: fft_stage_flat<T, 4>(data, nvec, pitch, N, n, inv, tws); break;
      case 2: fft_stage_flat<T, 2>(data, nvec, pitch, N, n, inv, tws); break;
      case 5: fft_stage_flat<T, 5>(data, nvec, pitch, N, n, inv, tws); break;
      default: fft_stage_flat<T, 3>(data, nvec, pitch, N, n, inv, tws); break;
    }
    __syncthreads();
    n /= r;
  }
}

template <typename T>
struct T3FftArgs {
  const cplx_t<T>* in;
  cplx_t<T>* out;
  int nin, n;                    // non-zero (centred-mode) count and FFT length along the transformed dimension
  int nvec_cta;                  // vectors (contiguous pass) or inner columns (strided pass) per CTA
  // contiguous pass: vector v of transform q starts at in + q * in_q + v * nin; written to out + q * out_q + v * n
  // strided pass:    element (a, k, c) at in + q * in_q + a * in_a + k * in_k + c (c contiguous, c < ninner)
  int64_t in_q, out_q, in_a, in_k, out_a, out_k;
  int64_t nvec;                  // contiguous pass: vectors per transform
  int ninner, nouter;            // strided pass
  const T* inv1; const T* inv2; const T* inv3;   // contiguous pass: deconvolution rows (inv2/inv3 by vector index)
  int nf2;                       // contiguous pass: vector v <-> (s3, s2) = (v / nf2, v % nf2)
  const cplx_t<T>* tw;
  FftStages st;
  const int32_t* pos;            // digit-reversed position of output j (n entries)
};

// FFT-ordered index of centred mode index k (mode k - nin / 2) inside an n-point grid
__device__ __forceinline__ int t3_mode_slot(int k, int nin, int n) {
  const int m = k - nin / 2;
  return m < 0 ? m + n : m;
}

// x pass: the transformed dimension is the contiguous one
template <typename T>
__global__ void __launch_bounds__(512)
t3_fft_contig_kernel(T3FftArgs<T> a) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t3f_smem[];
  C* vecs = (C*)t3f_smem;                         // nvec_cta * pitch
  const int n = a.n, nin = a.nin, pitch = n + 1;
  C* tw = vecs + (size_t)a.nvec_cta * pitch;      // tw_len <= n
  const int64_t v0 = (int64_t)blockIdx.x * a.nvec_cta;
  const int nv = (int)min((int64_t)a.nvec_cta, a.nvec - v0);
  const int q = blockIdx.y;
  const C* in = a.in + (int64_t)q * a.in_q + v0 * nin;
  for (int i = threadIdx.x; i < a.st.tw_len; i += blockDim.x) tw[i] = a.tw[i];
  for (int i = threadIdx.x; i < nv * pitch; i += blockDim.x) vecs[i] = make_c<T>(T(0), T(0));
  __syncthreads();
  // one vector after the other: the (s3, s2) row of a vector and its deconvolution factors are found once per
  // vector, not by 64-bit divisions per element
  for (int v = 0; v < nv; ++v) {
    const int64_t vg = v0 + v;
    const int s3 = (int)(vg / a.nf2), s2 = (int)(vg - (int64_t)s3 * a.nf2);
    const T sc23 = a.inv3 ? a.inv2[s2] * a.inv3[s3] : a.inv2[s2];
    const C* iv = in + (int64_t)v * nin;
    C* dv = vecs + v * pitch;
    for (int k = threadIdx.x; k < nin; k += blockDim.x) {
      const T sc = a.inv1[k] * sc23;
      C x = iv[k];
      x.x *= sc; x.y *= sc;
      dv[t3_mode_slot(k, nin, n)] = x;
    }
  }
  __syncthreads();
  smem_fft_cta<T>(vecs, nv, pitch, n, tw, a.st);
  C* out = a.out + (int64_t)q * a.out_q + v0 * n;
  for (int v = 0; v < nv; ++v) {
    const C* sv = vecs + v * pitch;
    C* ov = out + (int64_t)v * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) ov[j] = sv[a.pos[j]];
  }
}

// y / z pass: the transformed dimension is strided; a CTA owns nvec_cta consecutive inner columns
template <typename T>
__global__ void __launch_bounds__(512)
t3_fft_strided_kernel(T3FftArgs<T> a) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t3f_smem[];
  C* vecs = (C*)t3f_smem;                         // nvec_cta * pitch, vector = one inner column
  const int n = a.n, nin = a.nin, pitch = n + 1, cb = a.nvec_cta;
  C* tw = vecs + (size_t)cb * pitch;
  const int tiles = (a.ninner + cb - 1) / cb;
  const int outer = blockIdx.x / tiles, tile = blockIdx.x - outer * tiles;
  const int c0 = tile * cb, nc = min(cb, a.ninner - c0);
  const int q = blockIdx.y;
  const C* in = a.in + (int64_t)q * a.in_q + (int64_t)outer * a.in_a + c0;
  for (int i = threadIdx.x; i < a.st.tw_len; i += blockDim.x) tw[i] = a.tw[i];
  for (int i = threadIdx.x; i < nc * pitch; i += blockDim.x) vecs[i] = make_c<T>(T(0), T(0));
  __syncthreads();
  for (int i = threadIdx.x; i < nin * nc; i += blockDim.x) {
    const int k = i / nc, c = i - k * nc;
    vecs[c * pitch + t3_mode_slot(k, nin, n)] = in[(int64_t)k * a.in_k + c];
  }
  __syncthreads();
  smem_fft_cta<T>(vecs, nc, pitch, n, tw, a.st);
  C* out = a.out + (int64_t)q * a.out_q + (int64_t)outer * a.out_a + c0;
  for (int i = threadIdx.x; i < n * nc; i += blockDim.x) {
    const int j = i / nc, c = i - j * nc;
    out[(int64_t)j * a.out_k + c] = vecs[c * pitch + a.pos[j]];
  }
}

// ---- half-length form (ng = 2 nf exactly, nf even: upsampling factor 2, every BASELINE config) -------------------
// The padded vector holds its nin = n / 2 centred modes at the two ends of the n-point grid.  Shifted by nin / 2 it
// is y[j] = data[j], j < nin, followed by nin zeros, and the shift costs a factor (-i)^k on the outputs.  The first
// decimation-in-frequency stage of a vector whose second half is zero is free:
//     X[2k']     = (-1)^k'        FFT_nin( y[j] )[k']
//     X[2k' + 1] = (-1)^k' (-i)   FFT_nin( y[j] exp(+2 pi i j / n) )[k']
// so the pass runs two nin-point transforms per vector on half-length shared-memory vectors (twice the vectors per
// CTA, or twice the CTAs per SM, for the same shared memory), reading the input twice (the second time from L2).
template <typename T>
__device__ __forceinline__ cplx_t<T> t3_half_out(cplx_t<T> v, int k, int half) {
  // (-1)^k, and for the odd outputs a further factor -i
  if (k & 1) { v.x = -v.x; v.y = -v.y; }
  if (half) { const T t = v.x; v.x = v.y; v.y = -t; }
  return v;
}

template <typename T, int MINB>
__global__ void __launch_bounds__(MINB > 1 ? 256 : 512, MINB)
t3_fft_half_strided_kernel(T3FftArgs<T> a, const cplx_t<T>* __restrict__ wn) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t3f_smem[];
  C* vecs = (C*)t3f_smem;                         // nvec_cta * pitch, vector = one inner column, nin points
  const int nin = a.nin, pitch = nin + 1, cb = a.nvec_cta;
  C* tw = vecs + (size_t)cb * pitch;
  const int tiles = (a.ninner + cb - 1) / cb;
  const int outer = blockIdx.x / tiles, tile = blockIdx.x - outer * tiles;
  const int c0 = tile * cb, nc = min(cb, a.ninner - c0);
  const int q = blockIdx.y;
  const C* in = a.in + (int64_t)q * a.in_q + (int64_t)outer * a.in_a + c0;
  C* out = a.out + (int64_t)q * a.out_q + (int64_t)outer * a.out_a + c0;
  const unsigned inv_nc = nc > 1 ? 0xFFFFFFFFu / (unsigned)nc + 1u : 0u;     // exact i / nc for i < 2^16 * nc ... (i < 2^31 / nc)
  for (int i = threadIdx.x; i < a.st.tw_len; i += blockDim.x) tw[i] = a.tw[i];
  for (int half = 0; half < 2; ++half) {
    for (int i = threadIdx.x; i < nin * nc; i += blockDim.x) {
      const int k = inv_nc ? (int)__umulhi((unsigned)i, inv_nc) : i, c = i - k * nc;
      C x = in[(int64_t)k * a.in_k + c];
      if (half) x = cmul(x, wn[k]);
      vecs[c * pitch + k] = x;
    }
    __syncthreads();
    smem_fft_cta<T>(vecs, nc, pitch, nin, tw, a.st);
    for (int i = threadIdx.x; i < nin * nc; i += blockDim.x) {
      const int k = inv_nc ? (int)__umulhi((unsigned)i, inv_nc) : i, c = i - k * nc;
      out[(int64_t)(2 * k + half) * a.out_k + c] = t3_half_out<T>(vecs[c * pitch + a.pos[k]], k, half);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(512)
t3_fft_half_contig_kernel(T3FftArgs<T> a, const cplx_t<T>* __restrict__ wn) {
  using C = cplx_t<T>;
  extern __shared__ __align__(16) unsigned char t3f_smem[];
  C* vecs = (C*)t3f_smem;                         // nvec_cta * pitch
  const int nin = a.nin, n = 2 * nin, pitch = nin + 1;
  C* tw = vecs + (size_t)a.nvec_cta * pitch;
  const int64_t v0 = (int64_t)blockIdx.x * a.nvec_cta;
  const int nv = (int)min((int64_t)a.nvec_cta, a.nvec - v0);
  const int q = blockIdx.y;
  const C* in = a.in + (int64_t)q * a.in_q + v0 * nin;
  C* out = a.out + (int64_t)q * a.out_q + v0 * n;
  const unsigned inv_nin = 0xFFFFFFFFu / (unsigned)nin + 1u;
  for (int i = threadIdx.x; i < a.st.tw_len; i += blockDim.x) tw[i] = a.tw[i];
  for (int half = 0; half < 2; ++half) {
    for (int i = threadIdx.x; i < nv * nin; i += blockDim.x) {
      const int v = (int)__umulhi((unsigned)i, inv_nin), k = i - v * nin;
      const int64_t vg = v0 + v;
      const int s3 = (int)(vg / a.nf2), s2 = (int)(vg - (int64_t)s3 * a.nf2);
      T sc = a.inv1[k] * a.inv2[s2];
      if (a.inv3) sc *= a.inv3[s3];
      C x = in[i];
      x.x *= sc; x.y *= sc;
      if (half) x = cmul(x, wn[k]);
      vecs[v * pitch + k] = x;
    }
    __syncthreads();
    smem_fft_cta<T>(vecs, nv, pitch, nin, tw, a.st);
    for (int i = threadIdx.x; i < nv * nin; i += blockDim.x) {
      const int v = (int)__umulhi((unsigned)i, inv_nin), k = i - v * nin;
      out[(int64_t)v * n + 2 * k + half] = t3_half_out<T>(vecs[v * pitch + a.pos[k]], k, half);
    }
    __syncthreads();
  }
}

}  // namespace fv
