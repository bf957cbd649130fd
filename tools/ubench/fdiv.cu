// Which operands send __fdiv_rn to its slow path?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fdiv fdiv.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float a, float b, float* out, long long* cyc) {
  float x = a, acc = 0.f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) {
    const float q = __fdiv_rn(x, b);
    acc += q;
    x = __fmaf_rn(q, 0.0f, a);      // dependency, value unchanged
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
  const float cases[][2] = {{1.0f, 3.0f}, {0.004f, 0.16f}, {1234.0f, 384.0f}, {0.0f, 384.0f}, {0.01f, 0.16f}, {3.9e-3f, 1.6e-1f},
                            {-17.0f, 384.0f}, {1e-5f, 0.16f}, {0.004f, 2.5f}, {4e-3f, 1e-3f}};
  for (auto& c : cases) {
    k<<<1, 32>>>(c[0], c[1], out, cyc); cudaDeviceSynchronize();
    k<<<1, 32>>>(c[0], c[1], out, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%12g / %-8g : %6.1f cycles per division\n", c[0], c[1], (double)h / 4096);
  }
  return 0;
}
