// p2v_modules.cu - the reference's operators as stand-alone kernels on fp32 tensors (module-level use).
//
// The whole-model engine keeps activations as int8 codes; these entry points serve a Q-module that is called
// on its own after model_quant() (forward hooks of the analysis scripts, hand-built graphs, a bit_config with
// fp32 layers): fp32 dequantized values in, fp32 dequantized values out, same scalar arithmetic
// (p2v_math.cuh) as the engine kernels.
//
//   p2v_layernorm_int_f32      QIntLayerNorm.forward, mode 'int'           (models/ptq/layers.py:255-289)
//   p2v_softmax_log_int_f32    QIntSoftmax.forward, log-int-softmax        (models/ptq/layers.py:323-376)
//   p2v_requant_eltwise        QAct on the sum of two code tensors         (vit_fquant.py:449,466 residual adds)
//   p2v_select_histogram       one pass of an exact radix select           (torch.quantile / np.percentile of
//                                                                           models/ptq/observer/percentile.py:27-38)
// All four are streaming, HBM-bound row kernels: one warp per row (LN, softmax) or grid-stride (the others).
#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

static int grid_cap(int64_t blocks) {
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ---- QIntLayerNorm on fp32 ---------------------------------------------------------------------------------
// x_q = RNE(x / in_scale[c]) * in_mask[c]; exact integer row sums; then the dyadic affine of ln_code<false>.
// One warp per row; the row is read twice (the second read hits L1/L2).
__global__ void __launch_bounds__(256)
layernorm_int_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int d,
                         const float* __restrict__ in_scale, const float* __restrict__ in_mask, float in_scale1,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ out_scale, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const float* src = x + row * d;
    long long sum = 0, sumsq = 0;
    bool bad = false;
    for (int c = lane; c < d; c += 32) {
      const float q = fmul(rne(fdiv(src[c], in_scale[c])), in_mask[c]);
      bad |= !(fabsf(q) < 1048576.f);   // keeps C * sum(x^2) inside 63 bits; real codes are < 2^11
      const long long v = (long long)q;
      sum += v;
      sumsq += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    }
    if (bad) atomicExch(overflow, 1);
    const LnRow st = ln_row_stats(sum, sumsq, d, in_scale1);
    float* dst = out + row * d;
    for (int c = lane; c < d; c += 32) {
      const float q = fmul(rne(fdiv(src[c], in_scale[c])), in_mask[c]);
      const float os = out_scale[c];
      dst[c] = fmul(ln_code<false>(q, st, gamma[c], beta[c], os, 0.f), os);
    }
  }
}

// ---- QIntSoftmax (log-int-softmax) on fp32 --------------------------------------------------------------------
struct ExpConst {
  float scale, x0, b, c, floor_x;   // floor_x = n * x0 (the clamp of layers.py:355)
  int nbits;                        // n = 32
};

// the integer exp of one element, op for op as layers.py:353-361
__device__ __forceinline__ float int_exp(float x_int, const ExpConst& k) {
  x_int = fmaxf(x_int, k.floor_x);
  const float q = floorf(fdiv(x_int, k.x0));
  const float r = fsub(x_int, fmul(k.x0, q));
  float z = fadd(r, k.b);
  z = fmul(r, z);
  z = fadd(z, k.c);
  const float e = floorf(fmul(z, pow2i(k.nbits - (int)q)));
  return e < 0.f ? 0.f : e;
}

__global__ void __launch_bounds__(256)
softmax_log_int_f32_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ codes,
                           int64_t rows, int n, const ExpConst k, int levels) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const float* src = x + row * n;
    float mx = -INFINITY;
    for (int c = lane; c < n; c += 32) mx = fmaxf(mx, fdiv(src[c], k.scale));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // the integer exps are integers of <= 24 significant bits: their fp64 sum is exact, and rounding it once to
    // fp32 is what an exact integer accumulation gives (the engine's attention kernel does the same)
    double sum = 0.0;
    for (int c = lane; c < n; c += 32) sum += (double)int_exp(fsub(fdiv(src[c], k.scale), mx), k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float fsum = __double2float_rn(sum);
    for (int c = lane; c < n; c += 32) {
      const float e = int_exp(fsub(fdiv(src[c], k.scale), mx), k);
      const int code = softmax_log_code(fsum, e, levels);
      if (codes != nullptr) codes[row * n + c] = (uint8_t)code;
      if (out != nullptr) out[row * n + c] = code >= levels ? 0.f : pow2i(-code);
    }
  }
}

// ---- QAct over the sum of two code tensors ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
requant_eltwise_kernel(const int8_t* __restrict__ a, const int8_t* __restrict__ b, int8_t* __restrict__ out,
                       int64_t rows, int d, const float* __restrict__ a_scale, const float* __restrict__ b_scale,
                       const float* __restrict__ out_scale, float out_zp) {
  const int64_t total = rows * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    float v = fmul((float)a[i], a_scale[c]);
    if (b != nullptr) v = fadd(v, fmul((float)b[i], b_scale[c]));
    out[i] = (int8_t)quant_div(v, out_scale[c], out_zp, -128, 127);
  }
}

// ---- exact order statistics: one radix-select pass --------------------------------------------------------------
// Monotone key of a float: flip all bits of negatives, the sign bit of the rest (ascending float == ascending key).
__device__ __forceinline__ uint32_t order_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;

// histogram of key bits [shift, shift + 11) over the elements whose key agrees with `prefix` on `prefix_mask`
__global__ void __launch_bounds__(512)
select_histogram_kernel(const float* __restrict__ x, int64_t total, uint32_t prefix, uint32_t prefix_mask, int shift,
                        unsigned long long* __restrict__ hist) {
  __shared__ unsigned int local[kSelBins];
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) local[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = order_key(__ldg(x + i));
    if ((key & prefix_mask) == prefix) atomicAdd(&local[(key >> shift) & (kSelBins - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
    if (local[i]) atomicAdd(&hist[i], (unsigned long long)local[i]);
}

// ---- int4-packed weights -> int8 codes ---------------------------------------------------------------------------
// byte i holds code 2i in its low nibble and code 2i + 1 in its high nibble, both two's complement in [-8, 7].
// 16 packed bytes in, 32 codes out per thread and step.
__global__ void __launch_bounds__(256)
unpack_int4_kernel(const uint8_t* __restrict__ packed, int8_t* __restrict__ out, int64_t nbytes) {
  const int64_t n16 = nbytes >> 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(packed) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // spread the eight nibbles of w[k] over two words, then sign-extend every byte: (x ^ 8) - 8 per nibble
      const uint32_t lo = w[k] & 0x0f0f0f0fu, hi = (w[k] >> 4) & 0x0f0f0f0fu;
      const uint32_t a = __byte_perm(lo, hi, 0x5140), b = __byte_perm(lo, hi, 0x7362);   // bytes: l0 h0 l1 h1 | l2 h2 l3 h3
      o[2 * k] = __vsub4(a ^ 0x08080808u, 0x08080808u);
      o[2 * k + 1] = __vsub4(b ^ 0x08080808u, 0x08080808u);
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + 2 * i;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (nbytes & 15)) {   // tail bytes
    const int64_t i = (n16 << 4) + threadIdx.x;
    const int b = packed[i];
    out[2 * i] = (int8_t)(((b & 15) ^ 8) - 8);
    out[2 * i + 1] = (int8_t)((((b >> 4) & 15) ^ 8) - 8);
  }
}

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_layernorm_int_f32(const float* x, float* out, int64_t rows, int d, const float* in_scale,
                                     const float* in_mask, float in_scale1, const float* gamma, const float* beta,
                                     const float* out_scale, int* overflow_flag, void* stream) {
  P2V_REQUIRE(x && out && in_scale && in_mask && gamma && beta && out_scale && overflow_flag,
              "p2v_layernorm_int_f32: null pointer");
  P2V_REQUIRE(rows > 0 && d > 0, "p2v_layernorm_int_f32: bad shape rows=%lld d=%d", (long long)rows, d);
  P2V_REQUIRE(in_scale1 > 0.f, "p2v_layernorm_int_f32: in_scale1 must be positive");
  layernorm_int_f32_kernel<<<grid_cap((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      x, out, rows, d, in_scale, in_mask, in_scale1, gamma, beta, out_scale, overflow_flag);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_softmax_log_int_f32(const float* x, float* out, uint8_t* codes, int64_t rows, int n, float scale,
                                       float x0_int, float b_int, float c_int, int exp_bits, int levels,
                                       void* stream) {
  P2V_REQUIRE(x && (out || codes), "p2v_softmax_log_int_f32: null pointer");
  P2V_REQUIRE(rows > 0 && n > 0, "p2v_softmax_log_int_f32: bad shape rows=%lld n=%d", (long long)rows, n);
  P2V_REQUIRE(scale > 0.f && x0_int < 0.f, "p2v_softmax_log_int_f32: scale must be positive and x0_int negative");
  P2V_REQUIRE(exp_bits > 0 && exp_bits <= 60 && levels > 0 && levels <= 64, "p2v_softmax_log_int_f32: bad exp_bits/levels");
  ExpConst k;
  k.scale = scale;
  k.x0 = x0_int;
  k.b = b_int;
  k.c = c_int;
  k.floor_x = (float)exp_bits * x0_int;
  k.nbits = exp_bits;
  softmax_log_int_f32_kernel<<<grid_cap((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, out, codes, rows, n, k,
                                                                                        levels);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_requant_eltwise(const int8_t* a, const int8_t* b, int8_t* out, int64_t rows, int d,
                                   const float* a_scale, const float* b_scale, const float* out_scale, float out_zp,
                                   void* stream) {
  P2V_REQUIRE(a && out && a_scale && out_scale, "p2v_requant_eltwise: null pointer");
  P2V_REQUIRE((b == nullptr) == (b_scale == nullptr), "p2v_requant_eltwise: b and b_scale go together");
  P2V_REQUIRE(rows > 0 && d > 0, "p2v_requant_eltwise: bad shape");
  requant_eltwise_kernel<<<grid_cap((rows * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      a, b, out, rows, d, a_scale, b_scale, out_scale, out_zp);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_unpack_int4(const uint8_t* packed, int8_t* out, int64_t nbytes, void* stream) {
  P2V_REQUIRE(packed && out && nbytes > 0, "p2v_unpack_int4: bad arguments");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "p2v_unpack_int4: buffers must be 16-byte aligned");
  unpack_int4_kernel<<<grid_cap((nbytes / 16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(packed, out, nbytes);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_select_histogram(const float* x, int64_t total, uint32_t prefix, uint32_t prefix_mask, int shift,
                                    unsigned long long* hist, void* stream) {
  P2V_REQUIRE(x && hist, "p2v_select_histogram: null pointer");
  P2V_REQUIRE(total > 0 && shift >= 0 && shift <= 32 - kSelBits, "p2v_select_histogram: bad arguments");
  select_histogram_kernel<<<grid_cap((total + 511) / 512), 512, 0, (cudaStream_t)stream>>>(x, total, prefix,
                                                                                         prefix_mask, shift, hist);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
