// cuFFT micro-benchmark for the type-3 inner FFT grids (development aid).
#include <cufft.h>
#include <cuda_runtime.h>
#include <cstdio>
int main() {
  long long sizes[] = {3240, 3375, 3456, 3600, 3840, 4000, 4096, 1620, 1728, 1800, 1920, 2048};
  for (int prec = 1; prec <= 2; ++prec)
  for (int dim = 2; dim <= 3; ++dim)
  for (long long n : sizes) {
    if (dim == 3 && (prec == 1 || n < 3000)) continue;
    long long nz = dim == 3 ? 60 : 1;
    size_t esz = prec == 1 ? 8 : 16;
    size_t bytes = (size_t)n * n * nz * esz;
    void* d; if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc fail\n"); continue; }
    cudaMemset(d, 0, bytes);
    cufftHandle h; cufftCreate(&h);
    long long dims3[3] = {nz, n, n}, dims2[2] = {n, n};
    size_t work = 0;
    cufftResult r = cufftMakePlanMany64(h, dim, dim == 3 ? dims3 : dims2, nullptr, 1, n * n * nz, nullptr, 1, n * n * nz,
                                        prec == 1 ? CUFFT_C2C : CUFFT_Z2Z, 1, &work);
    if (r != CUFFT_SUCCESS) { printf("plan fail %d\n", (int)r); cudaFree(d); continue; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&]() { if (prec == 1) cufftExecC2C(h, (cufftComplex*)d, (cufftComplex*)d, CUFFT_INVERSE); else cufftExecZ2Z(h, (cufftDoubleComplex*)d, (cufftDoubleComplex*)d, CUFFT_INVERSE); };
    run(); run();
    cudaEventRecord(e0);
    int reps = dim == 3 ? 4 : 10;
    for (int i = 0; i < reps; ++i) run();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("prec=%d dim=%d n=%5lld  %9.3f ms  %7.1f MB  work %7.1f MB  eff(r+w) %7.1f GB/s  ns/pt %.3f\n", prec, dim, n, ms / reps, bytes / 1e6,
           work / 1e6, 2.0 * bytes / (ms / reps * 1e-3) / 1e9, ms / reps * 1e6 / ((double)n * n * nz));
    cufftDestroy(h); cudaFree(d);
  }
  return 0;
}
