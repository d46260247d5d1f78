// x-direct pass 1 of the fused type-1 path: host side (launch) of the kernel in type1_xdirect.cuh.
#include "nufft_internal.cuh"
#include "type1_xdirect.cuh"

namespace fv {

constexpr int XD_ROWS = 24;       // strip height: 2 * 24 accumulator registers per thread, three CTAs per SM

bool t1_xdirect_built(int w) { return w >= 2 && w <= kMaxW; }
int t1_xdirect_rows() { return XD_ROWS; }

template <int W>
static int launch_xd(fv_plan* P, const T1XdArgs& a, dim3 grid) {
  t1_xdirect_kernel<W, XD_ROWS><<<grid, XD_THREADS, 0, P->stream>>>(a);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

int t1_xdirect_pass1_entry(fv_plan* P, const int32_t* n_dev, int64_t n_cap, int nb, int ntr, const void* W, int64_t nf,
                           int w, double beta, const int32_t* iy0, const float* zy, const uint32_t* xt,
                           const uint32_t* hm0, const uint32_t* hm1, const fv_modeset::Tables* tab) {
  T1XdArgs a{};
  a.n_dev = n_dev; a.n_cap = n_cap; a.nf = (int)nf; a.R = XD_ROWS; a.w = w;
  a.beta = (float)beta; a.c = (float)(4.0 / ((double)w * w)); a.halfw = (float)(w / 2.0);
  a.ntr = ntr; a.W = (const float2*)W; a.iy0 = iy0; a.zy = zy; a.xt = xt; a.hm0 = hm0; a.hm1 = hm1;
  a.ncols = tab->ncols; a.col_k = tab->col_k; a.Tbuf = (float2*)P->tbuf;
  dim3 grid(ceil_div(nf, XD_ROWS), nb * ntr, ceil_div(tab->ncols, XD_THREADS));
  int rc = FV_OK;
  switch (w) {
#define XD_W(N) case N: rc = launch_xd<N>(P, a, grid); break;
    XD_W(2) XD_W(3) XD_W(4) XD_W(5) XD_W(6) XD_W(7) XD_W(8) XD_W(9) XD_W(10) XD_W(11) XD_W(12) XD_W(13) XD_W(14) XD_W(15) XD_W(16)
#undef XD_W
    default: set_error("x-direct pass 1: kernel width not built"); return FV_ERR_UNSUPPORTED;
  }
  return rc;
}

}  // namespace fv
