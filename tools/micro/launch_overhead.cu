// Micro-benchmark (developer tool): what does a launch cost as a function of dynamic shared memory and block size when it
// follows a small-shared-memory kernel?  Timed with CUDA events around back-to-back (small, big) pairs.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_small(int* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
__global__ void __launch_bounds__(1024, 1) k_big(int* p, unsigned long long* t) {
  if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); t[blockIdx.x] = g; }
  if (threadIdx.x == 0 && blockIdx.x == 0) p[1] += 1;
}
int main() {
  int* d; unsigned long long* t; cudaMalloc(&d, 64); cudaMalloc(&t, 8 * 1024); cudaMemset(d, 0, 64);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  int smems[] = {0, 48 * 1024, 100 * 1024, 160 * 1024, 200 * 1024, 226 * 1024};
  int thr[] = {128, 1024};
  cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int carve = 0; carve < 2; ++carve) {
    if (carve) { cudaFuncSetAttribute(k_small, cudaFuncAttributePreferredSharedMemoryCarveout, 100); cudaFuncSetAttribute(k_big, cudaFuncAttributePreferredSharedMemoryCarveout, 100); }
    for (int ti = 0; ti < 2; ++ti)
      for (int si = 0; si < 6; ++si) {
        float big = 0, small = 0;
        for (int it = 0; it < 60; ++it) {
          cudaEventRecord(e0, s);
          k_small<<<148, 128, 0, s>>>(d);
          cudaEventRecord(e1, s);
          k_big<<<148, thr[ti], smems[si], s>>>(d, t);
          cudaEventRecord(e2, s);
          cudaStreamSynchronize(s);
          float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
          if (it >= 10) { small += a; big += b; }
        }
        printf("carveout_pref %d threads %4d smem %3d KB : small %.2f us  big %.2f us  (%s)\n", carve, thr[ti], smems[si] / 1024, small / 50 * 1e3, big / 50 * 1e3,
               cudaGetErrorString(cudaGetLastError()));
      }
  }
  // back-to-back big kernels (no carveout change in between)
  for (int si = 0; si < 6; ++si) {
    float tot = 0;
    for (int it = 0; it < 60; ++it) {
      k_big<<<148, 1024, smems[si], s>>>(d, t);
      cudaEventRecord(e0, s);
      k_big<<<148, 1024, smems[si], s>>>(d, t);
      cudaEventRecord(e1, s);
      cudaStreamSynchronize(s);
      float a; cudaEventElapsedTime(&a, e0, e1);
      if (it >= 10) tot += a;
    }
    printf("big after big, smem %3d KB : %.2f us\n", smems[si] / 1024, tot / 50 * 1e3);
  }
  return 0;
}
