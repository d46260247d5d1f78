// Library-wide state of the C ABI: error string, launch counter, version.
#include "common.cuh"

namespace fv {
static thread_local std::string g_error;
int64_t g_launches = 0;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace fv

extern "C" const char* fv_last_error_string(void) { return fv::g_error.c_str(); }
extern "C" int fv_version(void) { return 100; }
extern "C" int64_t fv_launch_count(void) { return fv::g_launches; }

extern "C" int fv_device_count(int* count_host) {
  FV_REQUIRE(count_host, "null pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    *count_host = 0;
    fv::set_error("no CUDA device: fftvis_b200 has no CPU fallback");
    return FV_ERR_NO_DEVICE;
  }
  *count_host = n;
  return FV_OK;
}

// Strided copy between host and device on `stream` (cudaMemcpy2DAsync).  The engine streams each
// finished time slab of the (nf, nt, P, nbls) visibility array into the caller's page-locked result
// while later time steps are still being computed.  direction: 0 = device -> host, 1 = host -> device.
extern "C" int fv_memcpy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width,
                                 int64_t height, int direction, void* stream) {
  FV_REQUIRE(dst && src, "null pointer");
  FV_REQUIRE(width >= 0 && height >= 0 && dpitch >= width && spitch >= width, "bad pitch / extent");
  if (width == 0 || height == 0) return FV_OK;
  const cudaMemcpyKind kind = direction == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
  // cudaMemcpy2DAsync rejects pitches above cudaDevAttrMaxPitch (2 GiB - 1): rows that far apart
  // (e.g. nt * P * nbls * 16 B of a long observation) go as one contiguous copy per row instead
  int dev = 0, max_pitch = 0;
  FV_CUDA(cudaGetDevice(&dev));
  FV_CUDA(cudaDeviceGetAttribute(&max_pitch, cudaDevAttrMaxPitch, dev));
  if (dpitch == width && spitch == width) {
    FV_CUDA(cudaMemcpyAsync(dst, src, (size_t)width * (size_t)height, kind, (cudaStream_t)stream));
  } else if (dpitch > (int64_t)max_pitch || spitch > (int64_t)max_pitch) {
    for (int64_t r = 0; r < height; ++r)
      FV_CUDA(cudaMemcpyAsync((char*)dst + r * dpitch, (const char*)src + r * spitch, (size_t)width, kind,
                              (cudaStream_t)stream));
  } else {
    FV_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, kind,
                              (cudaStream_t)stream));
  }
  return FV_OK;
}
