// p2v_observe.cu - calibration statistics kernels (K12 of SURVEY.md): fused range and candidate-scale
// squared-error reductions for the activation observers.
//
//   p2v_observe_minmax     per-tensor or per-channel min / max            (MinmaxObserver.update / PtfObserver.update,
//                                                                          models/ptq/observer/minmax.py:16-39, ptf.py:14-31)
//   p2v_observe_scale_sse  sum (x - fakequant_k(x))^2 for up to 8 candidate scales, per tensor or per channel
//                          (the PoT search of minmax.py:180-242 on activations, the factor search of ptf.py:110-131)
// HBM-bound streaming reductions: 128-bit loads, fp64 accumulation (so that the arg-min over candidates agrees with
// the single-process fp32 reference and sums from several ranks can be all-reduced), one atomic per block.
#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

constexpr int kMaxCand = 8;

struct CandScales {
  float s[kMaxCand];
  int n;
};

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ---- per-tensor ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
minmax_tensor_kernel(const float* __restrict__ x, int64_t total, float* __restrict__ out_min, float* __restrict__ out_max) {
  float lo = INFINITY, hi = -INFINITY;
  const int64_t n4 = total >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (total & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f32(out_min, lo);
    atomic_max_f32(out_max, hi);
  }
}

__global__ void __launch_bounds__(256)
scale_sse_tensor_kernel(const float* __restrict__ x, int64_t total, const CandScales cand, float qmin, float qmax,
                        double* __restrict__ out) {
  double acc[kMaxCand];
#pragma unroll
  for (int k = 0; k < kMaxCand; ++k) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
#pragma unroll
    for (int k = 0; k < kMaxCand; ++k) {
      if (k < cand.n) {
        // (x / s).round().clamp(qmin, qmax) * s, then (x - xq)^2, every step rounded like the fp32 reference
        const float q = fminf(fmaxf(rintf(fdiv(v, cand.s[k])), qmin), qmax);
        const float d = fsub(v, fmul(q, cand.s[k]));
        acc[k] += (double)fmul(d, d);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxCand; ++k) {
    if (k < cand.n) {
      double a = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(out + k, a);
    }
  }
}

// ---- per-channel, x is [rows, c] with c contiguous ---------------------------------------------------------------
// A block owns a slab of rows; thread t walks the columns t, t + 256, ...: loads are coalesced along c.
constexpr int kSlabRows = 64;

__global__ void __launch_bounds__(256)
minmax_channel_kernel(const float* __restrict__ x, int64_t rows, int c, float* __restrict__ out_min,
                      float* __restrict__ out_max) {
  const int64_t r0 = (int64_t)blockIdx.x * kSlabRows;
  const int64_t r1 = r0 + kSlabRows < rows ? r0 + kSlabRows : rows;
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = __ldg(x + r * c + col);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    atomic_min_f32(out_min + col, lo);
    atomic_max_f32(out_max + col, hi);
  }
}

__global__ void __launch_bounds__(256)
scale_sse_channel_kernel(const float* __restrict__ x, int64_t rows, int c, const CandScales cand, float qmin, float qmax,
                         double* __restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * kSlabRows;
  const int64_t r1 = r0 + kSlabRows < rows ? r0 + kSlabRows : rows;
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    double acc[kMaxCand];
#pragma unroll
    for (int k = 0; k < kMaxCand; ++k) acc[k] = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = __ldg(x + r * c + col);
#pragma unroll
      for (int k = 0; k < kMaxCand; ++k) {
        if (k < cand.n) {
          const float q = fminf(fmaxf(rintf(fdiv(v, cand.s[k])), qmin), qmax);
          const float d = fsub(v, fmul(q, cand.s[k]));
          acc[k] += (double)fmul(d, d);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxCand; ++k)
      if (k < cand.n) atomicAdd(out + (int64_t)k * c + col, acc[k]);
  }
}

}  // namespace p2v

using namespace p2v;

static int grid_1d(int64_t work) {
  int64_t g = (work + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// out_min / out_max must be pre-filled with +inf / -inf by the caller (they are accumulated, which also lets
// several calibration batches share one running range).
extern "C" int p2v_observe_minmax(const float* x, int64_t rows, int channels, int per_channel, float* out_min,
                                  float* out_max, void* stream) {
  P2V_REQUIRE(x && out_min && out_max && rows > 0 && channels > 0, "p2v_observe_minmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (per_channel) {
    minmax_channel_kernel<<<(int)((rows + kSlabRows - 1) / kSlabRows), 256, 0, st>>>(x, rows, channels, out_min, out_max);
  } else {
    P2V_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "p2v_observe_minmax: x must be 16-byte aligned");
    const int64_t total = rows * channels;
    minmax_tensor_kernel<<<grid_1d(total / 4 + 1), 256, 0, st>>>(x, total, out_min, out_max);
  }
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

// out: fp64, [n_scales] (per tensor) or [n_scales, channels] (per channel), zero-filled by the caller (accumulated).
extern "C" int p2v_observe_scale_sse(const float* x, int64_t rows, int channels, int per_channel, const float* scales_host,
                                     int n_scales, float qmin, float qmax, double* out, void* stream) {
  P2V_REQUIRE(x && scales_host && out && rows > 0 && channels > 0, "p2v_observe_scale_sse: bad arguments");
  P2V_REQUIRE(n_scales > 0 && n_scales <= kMaxCand, "p2v_observe_scale_sse: 1..%d candidate scales", kMaxCand);
  CandScales cand;
  cand.n = n_scales;
  for (int k = 0; k < kMaxCand; ++k) cand.s[k] = k < n_scales ? scales_host[k] : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  if (per_channel)
    scale_sse_channel_kernel<<<(int)((rows + kSlabRows - 1) / kSlabRows), 256, 0, st>>>(x, rows, channels, cand, qmin, qmax, out);
  else
    scale_sse_tensor_kernel<<<grid_1d(rows * channels), 256, 0, st>>>(x, rows * channels, cand, qmin, qmax, out);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
