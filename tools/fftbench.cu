// cuFFT micro-benchmark: batched 2-D C2C in place, several sizes/batches (development aid).
#include <cufft.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
int main() {
  int sizes[] = {90, 960, 972, 1000, 1024, 486, 512};
  int batches[] = {1, 6, 12, 24, 48, 96, 256};
  for (int n : sizes) for (int b : batches) {
    if (n == 90) b *= 16;
    size_t bytes = (size_t)n * n * b * 8;
    void* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
    cufftHandle h; int dims[2] = {n, n};
    cufftPlanMany(&h, 2, dims, nullptr, 1, n * n, nullptr, 1, n * n, CUFFT_C2C, b);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cufftExecC2C(h, (cufftComplex*)d, (cufftComplex*)d, CUFFT_INVERSE);
    cudaEventRecord(e0);
    int reps = 20;
    for (int i = 0; i < reps; ++i) cufftExecC2C(h, (cufftComplex*)d, (cufftComplex*)d, CUFFT_INVERSE);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double us = ms * 1e3 / reps;
    printf("n=%4d batch=%4d  %8.1f us  %7.2f us/fft  %6.1f MB  eff(r+w) %7.1f GB/s\n", n, b, us, us / b, bytes / 1e6, 2.0 * bytes / (us * 1e-6) / 1e9);
    cufftDestroy(h); cudaFree(d);
  }
  // 1-D batched: rows of length n, count rows
  int n1[] = {960, 1024};
  for (int n : n1) for (int rows : {960 * 6, 960 * 48}) {
    size_t bytes = (size_t)n * rows * 8;
    void* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
    cufftHandle h; cufftPlan1d(&h, n, CUFFT_C2C, rows);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cufftExecC2C(h, (cufftComplex*)d, (cufftComplex*)d, CUFFT_INVERSE);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) cufftExecC2C(h, (cufftComplex*)d, (cufftComplex*)d, CUFFT_INVERSE);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double us = ms * 1e3 / 20;
    printf("1-D n=%4d rows=%6d %8.1f us  eff(r+w) %7.1f GB/s\n", n, rows, us, 2.0 * bytes / (us * 1e-6) / 1e9);
    cufftDestroy(h); cudaFree(d);
  }
  return 0;
}
