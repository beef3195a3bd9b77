// Micro-benchmark (developer tool): cycles per step of an ordered fp32 sum fed from shared memory, for several ways of issuing
// the loads.  One warp, 15 lanes = 15 columns of 15-float records, as in the ICP kernel's chain consumer.
//   nvcc -arch=sm_100a -O3 -fmad=false -o chain_bench chain_bench.cu && ./chain_bench
#include <cstdio>
#include <cuda_runtime.h>
#define NCOL 15
#define CH 256
#define STEPS (CH * 16)

template <int VAR>
__global__ void k(float* out, long long* cyc) {
  __shared__ __align__(16) float s[(CH + 32) * 16 * 2];
  const int lane = threadIdx.x;
  for (int i = lane; i < (CH + 32) * 16 * 2; i += 32) s[i] = 1.0f + (float)(i % 7) * 0.125f;
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  for (int rep = 0; rep < STEPS / CH; ++rep) {
    if (VAR == 0) {            // scalar LDS, register double buffer of 16
      if (lane < NCOL) {
        const float* p = s + lane;
        float v[16], w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = p[u * NCOL];
        for (int i = 0; i + 16 <= CH; i += 16) {
#pragma unroll
          for (int u = 0; u < 16; ++u) w[u] = p[(i + 16 + u) * NCOL];
#pragma unroll
          for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, v[u]);
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = w[u];
        }
      }
    } else if (VAR == 1) {     // plain unrolled loop
      if (lane < NCOL) {
        const float* p = s + lane;
#pragma unroll 16
        for (int i = 0; i < CH; ++i) acc = __fadd_rn(acc, p[i * NCOL]);
      }
    } else if (VAR == 2) {     // column-major, LDS.128: 4 steps per load, double buffer of 4 loads
      if (lane < NCOL) {
        const float4* p = reinterpret_cast<const float4*>(s + lane * (CH + 4));
        float4 v[4], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[u];
        for (int i = 0; i + 4 <= CH / 4; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) w[u] = p[i + 4 + u];
#pragma unroll
          for (int u = 0; u < 4; ++u) { acc = __fadd_rn(acc, v[u].x); acc = __fadd_rn(acc, v[u].y); acc = __fadd_rn(acc, v[u].z); acc = __fadd_rn(acc, v[u].w); }
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = w[u];
        }
      }
    } else if (VAR == 3) {     // column-major LDS.128, three register sets (loads two groups ahead)
      if (lane < NCOL) {
        const float4* p = reinterpret_cast<const float4*>(s + lane * (CH + 4));
        float4 a[4], b[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = p[u]; b[u] = p[4 + u]; }
        for (int i = 0; i + 4 <= CH / 4; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = p[i + 8 + u];
#pragma unroll
          for (int u = 0; u < 4; ++u) { acc = __fadd_rn(acc, a[u].x); acc = __fadd_rn(acc, a[u].y); acc = __fadd_rn(acc, a[u].z); acc = __fadd_rn(acc, a[u].w); }
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = b[u]; b[u] = c[u]; }
        }
      }
    } else if (VAR == 4) {     // registers only (no loads): the FADD latency floor
      if (lane < NCOL) {
        float x = s[lane];
#pragma unroll 16
        for (int i = 0; i < CH; ++i) acc = __fadd_rn(acc, x);
      }
    } else if (VAR == 5) {     // LDS.128 column-major, one big group of 8 loads (32 steps) double-buffered
      if (lane < NCOL) {
        const float4* p = reinterpret_cast<const float4*>(s + lane * (CH + 4));
        float4 v[8], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[u];
        for (int i = 0; i + 8 <= CH / 4; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = p[i + 8 + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) { acc = __fadd_rn(acc, v[u].x); acc = __fadd_rn(acc, v[u].y); acc = __fadd_rn(acc, v[u].z); acc = __fadd_rn(acc, v[u].w); }
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = w[u];
        }
      }
    }
  }
  long long t1 = clock64();
  if (lane < NCOL) out[lane] = acc;
  if (lane == 0) cyc[0] = t1 - t0;
}

template <int VAR> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 64 * 4); cudaMalloc(&cyc, 8);
  k<VAR><<<1, 32>>>(out, cyc); k<VAR><<<1, 32>>>(out, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-60s %6.2f cycles/step (%s)\n", name, (double)h / STEPS, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<4>("registers only (FADD latency floor)");
  run<1>("scalar LDS, plain loop unrolled 16");
  run<0>("scalar LDS, register double buffer of 16");
  run<2>("LDS.128 column-major, double buffer of 4 loads");
  run<3>("LDS.128 column-major, three register sets");
  run<5>("LDS.128 column-major, double buffer of 8 loads");
  return 0;
}
