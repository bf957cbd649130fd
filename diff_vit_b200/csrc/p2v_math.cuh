// p2v_math.cuh - the scalar arithmetic of the P2-ViT integer forward, written once for host and device.
//
// Every function here restates, operation for operation, what the reference's fp32 fake-quant graph
// computes on integer-valued tensors, so that integer codes produced by the sm_100a kernels equal the
// reference's codes.  The functions are __host__ __device__: the kernels inline them, and
// tests/hostmath builds them with g++ (-ffp-contract=off) so the arithmetic can be checked against the
// oracle on a CPU-only box.  On the device every fp32 step uses an explicitly rounded intrinsic so that
// nvcc cannot contract a mul+add into an FMA where the reference rounds twice.
//
// Reference lines are cited per function (paths relative to the reference repo).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define P2V_HD __host__ __device__ __forceinline__
#else
#define P2V_HD static inline
#endif

namespace p2v {

#if defined(__CUDA_ARCH__)
P2V_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
P2V_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
P2V_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
P2V_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
P2V_HD float fsqrt(float a) { return __fsqrt_rn(a); }
P2V_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
// host build: compiled with -ffp-contract=off, so each operator rounds once
P2V_HD float fmul(float a, float b) { return a * b; }
P2V_HD float fadd(float a, float b) { return a + b; }
P2V_HD float fsub(float a, float b) { return a - b; }
P2V_HD float fdiv(float a, float b) { return a / b; }
P2V_HD float fsqrt(float a) { return sqrtf(a); }
P2V_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

// round-half-to-even, the rounding of torch.round
P2V_HD float rne(float v) { return rintf(v); }

P2V_HD int clamp_i(float v, int lo, int hi) {
  // clamp AFTER round (models/ptq/quantizer/uniform.py:86-87); v is already integral
  v = v < (float)lo ? (float)lo : v;
  v = v > (float)hi ? (float)hi : v;
  return (int)v;
}

P2V_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
P2V_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// 2^n for integer n in [-126, 127]
P2V_HD float pow2i(int n) { return u2f((uint32_t)(n + 127) << 23); }

// ---------------------------------------------------------------------------------------------
// Activation quantizer: codes = clamp(RNE(x / s + zp), qmin, qmax)
// (models/ptq/quantizer/uniform.py:82-88).  `rs` is the exact reciprocal when s is a power of two
// (x * 2^-e == x / 2^e in fp32), which is what P2-ViT's minmax observer produces.
// ---------------------------------------------------------------------------------------------
P2V_HD int quant_div(float x, float s, float zp, int qmin, int qmax) {
  return clamp_i(rne(fadd(fdiv(x, s), zp)), qmin, qmax);
}
P2V_HD int quant_pot(float x, float rs, int qmin, int qmax) { return clamp_i(rne(fmul(x, rs)), qmin, qmax); }

// erf-GELU as ATen evaluates it: x * 0.5 * (1 + erf(x * sqrt(1/2))) (nn.GELU() default,
// models/layers_quant.py:331).  The 0.5 scaling is exact, so only the association of the last
// product matters; the add must not be fused into it.
P2V_HD float gelu_erf(float x) {
  const float kAlpha = 0.70710678118654752440f;
  return fmul(fmul(x, 0.5f), fadd(1.0f, erff(fmul(x, kAlpha))));
}

// ---------------------------------------------------------------------------------------------
// GEMM epilogue (models/ptq/layers.py:173-178 followed by the QAct that consumes it).
//   y  = acc * (s_in * s_w[n]) + bias[n]      one rounding: the fp32 GEMM of the reference is exact
//                                             on integer-valued operands (|acc| < 2^24), then adds bias
//   y  = gelu(y)                              fc1 only
//   q  = clamp(RNE(y / s_out[n] + zp))        next QAct
// Residual form (models/vit_fquant.py:431,468): the branch output is dequantized again and added to the
// dequantized residual stream, then re-quantized by the block-level PTF QAct, all in fp32:
//   q2 = clamp(RNE((res * s_res[n] + q * s_out[n]) / s_out2[n]))
// ---------------------------------------------------------------------------------------------
struct EpiChannel {
  float acc_scale;   // s_in * s_w[n]
  float bias;        // fp32 bias, never quantized
  float out_scale;   // s_out[n]
  float out_rscale;  // 1 / s_out[n], exact when power of two
  float res_scale;   // s_res[n]
  float out2_scale;  // s_out2[n]
};

enum : uint32_t {
  EPI_GELU = 1u,       // apply erf-GELU before re-quantization
  EPI_RESIDUAL = 2u,   // add the residual stream and re-quantize on the block-level grid
  EPI_OUT_POT = 4u,    // s_out is a power of two for every channel: multiply by out_rscale
  EPI_OUT_F32 = 8u,    // additionally emit the dequantized value (logits)
};

template <uint32_t FLAGS>
P2V_HD int epilogue_code(int acc, const EpiChannel& c, float out_zp) {
  float y = ffma((float)acc, c.acc_scale, c.bias);
  if (FLAGS & EPI_GELU) y = gelu_erf(y);
  if (FLAGS & EPI_OUT_POT) return quant_pot(y, c.out_rscale, -128, 127);
  return quant_div(y, c.out_scale, out_zp, -128, 127);
}

P2V_HD int residual_code(int q, int res, const EpiChannel& c) {
  float sum = fadd(fmul((float)res, c.res_scale), fmul((float)q, c.out_scale));
  return clamp_i(rne(fdiv(sum, c.out2_scale)), -128, 127);
}

// ---------------------------------------------------------------------------------------------
// Integer LayerNorm (models/ptq/layers.py:255-289).  Row statistics are exact integers; everything
// after follows the reference's fp32 op order.
// ---------------------------------------------------------------------------------------------
struct LnRow {
  float t;  // in_scale1 / std
  float u;  // mean / std
};

// sum, sumsq: exact integer sums of x_q * mask over the row.  C: channel count.
// scale_over_c = fl(in_scale1 / C), the row-independent factor of the standard deviation
P2V_HD LnRow ln_row_stats(long long sum, long long sumsq, int C, float in_scale1, float scale_over_c) {
  float fs = (float)sum, fq = (float)sumsq, fc = (float)C;
  float mean = fmul(fdiv(fs, fc), in_scale1);                       // x_q.mean(-1) * in_scale1
  float var = fsub(fmul(fc, fq), fmul(fs, fs));                     // C*sum(x^2) - sum(x)^2
  float stdv = fmul(scale_over_c, fsqrt(var));                      // (in_scale1 / C) * sqrt(.)
  LnRow r;
  r.t = fdiv(in_scale1, stdv);
  r.u = fdiv(mean, stdv);
  return r;
}
P2V_HD LnRow ln_row_stats(long long sum, long long sumsq, int C, float in_scale1) {
  return ln_row_stats(sum, sumsq, C, in_scale1, fdiv(in_scale1, (float)C));
}

// One element: xq = code * mask (integer-valued), gamma/beta the LN affine, out_scale the LN output
// grid (out_quantizer.scale * SmoothQuant channel scale) with its exact reciprocal when POT.
// Returns the LN code on the out grid, before any clamp (it is an fp32-held integer in the reference).
template <bool POT>
P2V_HD float ln_code(float xq, const LnRow& row, float gamma, float beta, float out_scale, float out_rscale) {
  float a1 = fmul(row.t, gamma);
  float A = POT ? fmul(a1, out_rscale) : fdiv(a1, out_scale);
  float absA = fabsf(A);
  float sign = A > 0.f ? 1.f : (A < 0.f ? -1.f : 0.f);
  // get_MN (layers.py:234-238): N = clamp(7 - floor(log2 |A|), 0, 31), M = clamp(floor(|A| 2^N), 0, 255).
  // floor(log2) is taken from the exponent field; torch evaluates log2 in fp32, which differs only when
  // |A| sits within ~2 ulp below a power of two (measure-zero, inside the <=1 LSB allowance).
  int e = (int)((f2u(absA) >> 23) & 0xffu) - 127;
  int N = 7 - e;
  N = N < 0 ? 0 : (N > 31 ? 31 : N);
  float p2N = pow2i(N);
  float M = floorf(fmul(absA, p2N));
  M = M > 255.f ? 255.f : M;
  float b0 = fsub(beta, fmul(row.u, gamma));
  float b1 = POT ? fmul(b0, out_rscale) : fdiv(b0, out_scale);
  float Bq = rne(fmul(b1, p2N));
  // (sign * M * xq + B) / 2^N : sign*M*xq is exact (< 2^24), so the add is the only rounding
  float y = fadd(fmul(fmul(sign, M), xq), Bq);
  return rne(fmul(y, pow2i(-N)));
}

// ---------------------------------------------------------------------------------------------
// log-int-softmax tail (models/ptq/layers.py:323-329,367-376): given the exact integer row sum and the
// element's integer exp, the 4-bit log2 code k (2^bits means "probability 0").
// ---------------------------------------------------------------------------------------------
P2V_HD int softmax_log_code(float row_sum, float exp_int, int levels) {
  float r = rne(fdiv(row_sum, exp_int));        // torch.round(exp_int_sum / exp_int)
  uint32_t b = f2u(r);
  int big = (int)((b >> 23) & 0xffu) - 127;     // floor(log2 r), r >= 1 integral
  big += (int)((b >> 22) & 1u);                 // +1 when (r - 2^big) >= 2^(big-1)
  big = big < 0 ? 0 : big;
  return big >= levels ? levels : big;
}

}  // namespace p2v
