// p2v_requant.cuh - re-quantization helpers shared by the GEMM epilogues and the element kernels: saturating
// RNE pack to int8 and the division-free (tie-guarded) RNE(y / s + zp).
#pragma once
#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

// RNE + saturate + pack of four fp32 values into four int8 (two F2IP.S8.F32 on sm_100)
__device__ __forceinline__ uint32_t pack_sat4(float v0, float v1, float v2, float v3) {
  // cvt.pack d, a, b, c: d[7:0] = sat(b), d[15:8] = sat(a), d[31:16] = c[15:0]
  uint32_t hi, r;
  const int i0 = __float2int_rn(v0), i1 = __float2int_rn(v1), i2 = __float2int_rn(v2), i3 = __float2int_rn(v3);
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(i3), "r"(i2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(i1), "r"(i0), "r"(hi));
  return r;
}

// RNE(fl(y / s) + zp), the unclamped code of a quantizer with a non-power-of-two scale, without paying an
// IEEE division per element: t = y * fl(1/s) + zp is within a few ulp of the exact operand of the
// rounding, so rint(t) can differ from the exact result only when t lies within kTieGuard of a
// half-integer; only those elements (~0.02 %) take the exact division.  For |t| >= 512 both paths
// saturate to the same int8 code, so the absolute guard is sufficient: below 128.5 the operand t = y * fl(1/s)
// is within 1.5 ulp(128) = 2.3e-5 of the exactly divided one, and the guard is 5x that.
constexpr float kTieGuard = 0.0001220703125f;  // 2^-13
static __device__ __noinline__ float div_round_exact(float y, float s, float zp) { return rintf(fadd(fdiv(y, s), zp)); }

// Round-half-even on the FMA pipe: (t + 1.5 * 2^23) - 1.5 * 2^23 is RNE(t) for |t| < 2^22.  Larger |t| come back
// as some large value of the same sign, which saturates to the same int8 code as RNE(t) would.  (FRND runs on
// the conversion unit, which ncu showed ~40 % busy in every GEMM epilogue.)
__device__ __forceinline__ float rne_small(float t) { return fsub(fadd(t, 12582912.0f), 12582912.0f); }
__device__ __forceinline__ float div_round(float y, float s, float rs, float zp) {
  const float t = fadd(fmul(y, rs), zp);
  float r = rne_small(t);
  if (fabsf(fabsf(fsub(t, r)) - 0.5f) < kTieGuard) r = div_round_exact(y, s, zp);   // rare: keeps the hot loop small
  return r;
}

// Four at a time: the common path is branch-free so the four dependency chains interleave; one rarely taken branch
// per group re-does the flagged elements exactly.  The tie test of the group is one comparison: with df = t - RNE(t)
// in [-1/2, 1/2], q = df^2 - 1/4 = -(1/2 - |df|)(1/2 + |df|) is within kTieGuard of zero whenever |df| is within
// kTieGuard of 1/2 (and hardly ever otherwise), so "min |q| over the group < kTieGuard" flags a superset of the
// per-element test |(|df| - 1/2)| < kTieGuard at 1.4 instead of 3.3 instructions per element (one packed FMA per
// pair, a three-input minimum per group).
__device__ __forceinline__ void div_round4(const float (&y)[4], const float (&s)[4], const float (&rs)[4], float zp,
                                           float (&r)[4]) {
  // element pairs on the packed fp32 instructions (each half rounds like the scalar _rn op)
  const float2 zp2 = make_float2(zp, zp), kMagic = make_float2(12582912.0f, 12582912.0f);
  const float2 kMagicNeg = make_float2(-12582912.0f, -12582912.0f), kNegOne = make_float2(-1.0f, -1.0f);
  float2 df[2];
#pragma unroll
  for (int e = 0; e < 4; e += 2) {
    // (a contraction of this mul + add into one fma moves t by at most one ulp, far inside the tie guard)
    const float2 t = fadd2(fmul2(make_float2(y[e], y[e + 1]), make_float2(rs[e], rs[e + 1])), zp2);
    const float2 rr = fadd2(fadd2(t, kMagic), kMagicNeg);      // rne_small
    df[e >> 1] = ffma2(rr, kNegOne, t);                            // t - r, exact
    r[e] = rr.x;
    r[e + 1] = rr.y;
  }
  const float2 q0 = ffma2(df[0], df[0], make_float2(-0.25f, -0.25f)), q1 = ffma2(df[1], df[1], make_float2(-0.25f, -0.25f));
  if (fminf(fminf(fabsf(q0.x), fabsf(q0.y)), fminf(fabsf(q1.x), fabsf(q1.y))) < kTieGuard) {
    // rare
    const float d4[4] = {df[0].x, df[0].y, df[1].x, df[1].y};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (!(fabsf(fabsf(d4[e]) - 0.5f) >= kTieGuard)) r[e] = div_round_exact(y[e], s[e], zp);
  }
}

}  // namespace p2v
