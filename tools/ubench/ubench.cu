// Micro-benchmarks of issue rates on sm_100a (cycles per warp instruction per SM sub-partition):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
// One CTA on one SM; W warps per sub-partition (W * 4 warps in the CTA), every warp runs N independent ops per loop.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int OP>
__global__ void k(float* out, u64* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const u64 t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 6) asm volatile("max.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    }
    if (OP == 4 || OP == 5) {
      u64* p = reinterpret_cast<u64*>(a);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
        if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[i]));
      }
    }
  }
  const u64 t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int ops_per_iter) {
  float* out; u64* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int w = 1; w <= 8; w *= 2) {
    k<OP><<<1, 128 * w>>>(out, cyc, iters);
    k<OP><<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    u64 h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-14s warps/SMSP %d: %.2f cycles per warp instruction per SMSP\n", name, w, (double)h / ((double)iters * ops_per_iter * w));
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.approx", 16); run<1>("tanh.approx", 16); run<2>("rcp.approx", 16); run<3>("fma.f32", 16);
  run<4>("fma.f32x2", 8); run<5>("add.f32x2", 8); run<6>("max3.f32", 16);
  return 0;
}
