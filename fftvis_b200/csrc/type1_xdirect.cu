// x-direct pass 1 of the fused type-1 path: host side (launches) of the kernels in type1_xdirect.cuh.
#include "nufft_internal.cuh"
#include "type1_xdirect.cuh"

namespace fv {

bool t1_xdirect_built(int w) { return w >= 2 && w <= kMaxW; }
int t1_xdirect_rows() { return XD_ROWS; }

template <int W>
static int launch_xd(fv_plan* P, const T1XdArgs& a, dim3 grid) {
  t1_xdirect_kernel<W, XD_ROWS><<<grid, XD_THREADS, 0, P->stream>>>(a);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

int t1_xdirect_pass1_entry(fv_plan* P, const void* bx, const void* by, const int32_t* n_dev, int64_t n_cap, int nb, int ntr,
                           const void* W, int64_t nf, int w, double beta, const fv_modeset::Tables* tab) {
  const int ns = ceil_div(nf, XD_ROWS);
  const size_t per = (size_t)nb * n_cap;
  const size_t smem_prep = sizeof(int) * ((size_t)(XD_PREP_THREADS / 32) * ns + ns + 1 + XD_PREP_THREADS / 32);
  if (smem_prep > 200 * 1024) { set_error("x-direct pass 1: too many strips"); return FV_ERR_UNSUPPORTED; }
  const size_t lcap = (size_t)XD_LIST_PER_SRC * ((n_cap + XD_PREP_SPLIT - 1) / XD_PREP_SPLIT);
  int rc = ensure(&P->prep, &P->prep_bytes, 4 * (3 * per + (size_t)nb * XD_PREP_SPLIT * (lcap + ns + 1)) + 64);
  if (rc) return rc;
  T1XdPrepArgs pa{};
  pa.bx = (const float*)bx; pa.by = (const float*)by; pa.n_dev = n_dev; pa.n_cap = n_cap; pa.bp = P->bp_dev;
  pa.nf = (int)nf; pa.w = w; pa.nstrips = ns;
  pa.yrow = (int32_t*)P->prep;
  pa.zy = (float*)(pa.yrow + per);
  pa.xt = (uint32_t*)(pa.zy + per);
  pa.list = (int32_t*)(pa.xt + per);
  pa.lcap = (int64_t)lcap;
  pa.strip_off = pa.list + (size_t)nb * XD_PREP_SPLIT * lcap;
  {
    StageScope ts(P, FV_STAGE_ZERO);
    FV_CUDA(cudaFuncSetAttribute(t1_xd_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
    t1_xd_prep_kernel<<<dim3(nb, XD_PREP_SPLIT), XD_PREP_THREADS, smem_prep, P->stream>>>(pa);
    FV_LAUNCH_CHECK();
  }
  T1XdArgs a{};
  a.n_cap = n_cap; a.nf = (int)nf; a.nstrips = ns;
  a.beta = (float)beta; a.c = (float)(4.0 / ((double)w * w)); a.halfw = (float)(w / 2.0);
  a.ntr = ntr; a.W = (const float2*)W; a.yrow = pa.yrow; a.zy = pa.zy; a.xt = pa.xt;
  a.strip_off = pa.strip_off; a.list = pa.list; a.lcap = pa.lcap;
  a.ncols = tab->ncols; a.col_k = tab->col_k; a.Tbuf = (float2*)P->tbuf;
  dim3 grid(ns, nb * ntr, ceil_div(tab->ncols, XD_THREADS));
  StageScope ts(P, FV_STAGE_SPREAD);
  switch (w) {
#define XD_W(N) case N: rc = launch_xd<N>(P, a, grid); break;
    XD_W(2) XD_W(3) XD_W(4) XD_W(5) XD_W(6) XD_W(7) XD_W(8) XD_W(9) XD_W(10) XD_W(11) XD_W(12) XD_W(13) XD_W(14) XD_W(15) XD_W(16)
#undef XD_W
    default: set_error("x-direct pass 1: kernel width not built"); return FV_ERR_UNSUPPORTED;
  }
  return rc;
}

}  // namespace fv
