// Issue-rate micro-benchmark for the instructions the epilogues are made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs 1024 threads (8 warps per scheduler) on one SM; every thread keeps 8 independent chains going so
// neither latency nor dependencies limit issue.  Reported: warp instructions per cycle per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint64_t* out, long long* cycles, uint32_t seed) {
  uint32_t r[CHAINS * 2];
#pragma unroll
  for (int i = 0; i < CHAINS * 2; ++i) r[i] = seed + threadIdx.x * 31 + i;
  const uint32_t c0 = seed | 1, c1 = seed * 3 + 7;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      uint32_t& a = r[2 * i];
      uint32_t& b = r[2 * i + 1];
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
      if (OP == 1) {
        uint64_t v = ((uint64_t)b << 32) | a, m = ((uint64_t)c0 << 32) | c0, d = ((uint64_t)c1 << 32) | c1;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(m), "l"(d));
        a = (uint32_t)v; b = (uint32_t)(v >> 32);
      }
      if (OP == 2) {
        uint64_t v = ((uint64_t)b << 32) | a, m = ((uint64_t)c0 << 32) | c0;
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(m));
        a = (uint32_t)v; b = (uint32_t)(v >> 32);
      }
      if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c0), "r"(c1));
      if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c0));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a) : "r"(c0));
      if (OP == 7) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a));
      if (OP == 8) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a));
      if (OP == 9) asm volatile("{.reg .b32 t; cvt.rni.s32.f32 t, %0; cvt.pack.sat.s8.s32.b32 %0, t, t, %1;}" : "+r"(a) : "r"(c0));
      if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
      if (OP == 11) asm volatile("max.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
      if (OP == 12) {  // alternate FMA-pipe and ALU-pipe instructions
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(c0), "r"(c1));
      }
      if (OP == 13) {  // FFMA2 + LOP3
        uint64_t v = ((uint64_t)b << 32) | a, m = ((uint64_t)c0 << 32) | c0, d = ((uint64_t)c1 << 32) | c1;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(m), "l"(d));
        a = (uint32_t)v; b = (uint32_t)(v >> 32);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(2 * i + 3) % (CHAINS * 2)]) : "r"(c0), "r"(c1));
      }
    }
  }
  const long long t1 = clock64();
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS * 2; ++i) acc += r[i];
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  uint64_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  k<OP><<<1, 1024>>>(out, cyc, 12345u); cudaDeviceSynchronize();
  k<OP><<<1, 1024>>>(out, cyc, 12345u); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double inst_per_warp = (double)ITERS * CHAINS * per_iter;
  printf("%-28s %8.3f warp-inst / cycle / scheduler   (%lld cycles)\n", name, inst_per_warp * 8 / (double)h, h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FADD2", 1); run<10>("FADD", 1); run<11>("FMNMX", 1);
  run<3>("IMAD", 1); run<4>("LOP3", 1); run<5>("IADD", 1); run<6>("PRMT", 1);
  run<7>("I2FP", 1); run<8>("MUFU.EX2", 1); run<9>("F2I + I2IP (or F2IP)", 2);
  run<12>("FFMA + LOP3 interleaved", 2); run<13>("FFMA2 + LOP3 interleaved", 2);
  return 0;
}
