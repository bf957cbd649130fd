// p2v_rowops.cu - the HBM-bound element / row kernels of the integer forward:
//   input QAct + im2col, token assembly, integer LayerNorm (+ consumer QAct), standalone fake-quant.
// All fp32 steps follow the reference's op order through p2v_math.cuh.
#include "p2v_common.cuh"
#include "p2v_math.cuh"
#include "p2v_requant.cuh"

namespace p2v {

__device__ __forceinline__ uint32_t pack4i(int q0, int q1, int q2, int q3) {
  return (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) |
         ((uint32_t)(q3 & 0xff) << 24);
}

// RNE(x / s + zp) clamped to int8 without an IEEE division per element: t = x * fl(1/s) + zp is within a few ulp of
// the exactly divided operand, so the rounding can differ only within 2^-13 of a tie; those rare elements are
// redone with the exact division (same guard as the GEMM epilogue, p2v_gemm.cu div_round).
__device__ __forceinline__ int quant_div_guarded(float x, float s, float rs, float zp) {
  const float t = fadd(fmul(x, rs), zp);
  float r = rintf(t);
  if (fabsf(fabsf(fsub(t, r)) - 0.5f) < 0.0001220703125f) r = rintf(fadd(fdiv(x, s), zp));
  return clamp_i(r, -128, 127);
}

// ---- input quantizer + patchify ------------------------------------------------------------------
// One thread converts 16 consecutive pixels of one image row (64 B fp32 in, 16 B codes out).  With
// p % 16 == 0 those 16 pixels are 16 consecutive K entries of one patch row: K = (c*p + kh)*p + kw.
__global__ void quant_patchify_kernel(const float* __restrict__ x, int8_t* __restrict__ codes, int b, int c,
                                      int h, int w, int p, float scale, float zp, int64_t total16) {
  const int gw = w / p, gh = h / p;
  const int k = c * p * p;
  const int w16 = w / 16;
  const float rs = __frcp_rn(scale);
  // 32-bit index arithmetic (total16 < 2^31 is checked at launch): three 64-bit divisions per 16 pixels cost more
  // than the conversion of those pixels (ncu: 19 instructions per pixel before, most of them division sequences)
  const int n16 = (int)total16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
    const int xs = i % w16;
    int r = i / w16;
    const int y = r % h;
    r /= h;
    const int ch = r % c;
    const int img = r / c;
    const float4* src = reinterpret_cast<const float4*>(x + (((int64_t)img * c + ch) * h + y) * w + xs * 16);
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(src + j);
    uint32_t o[4];
    const float s4[4] = {scale, scale, scale, scale}, rs4[4] = {rs, rs, rs, rs};
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // four pixels per guarded rounding, packed fp32 + one saturating pack (p2v_requant.cuh)
      const float y4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      float r4[4];
      div_round4(y4, s4, rs4, zp, r4);
      o[j] = pack_sat4(r4[0], r4[1], r4[2], r4[3]);
    }
    const int px = xs * 16;
    const int pw = px / p, kw = px % p, ph = y / p, kh = y % p;
    const int64_t row = ((int64_t)img * gh + ph) * gw + pw;
    *reinterpret_cast<uint4*>(codes + row * k + ((int64_t)ch * p + kh) * p + kw) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// The same from 8-bit pixels: x = (pixel / 255 - mean[c]) / std[c] (torchvision ToTensor + Normalize, the
// preprocessing of the reference's loaders, test_quant.py:96-110), then the input quantizer.  A pixel has 256 values,
// so each block evaluates the fp32 expression once per (channel, value) into a shared-memory table of codes and the
// image is a byte gather: a quarter of the fp32 input's HBM and host-link traffic, same codes.
struct U8Norm {
  float mean[4], stdv[4];
};
__global__ void quant_patchify_u8_kernel(const uint8_t* __restrict__ x, int8_t* __restrict__ codes, int b, int c,
                                         int h, int w, int p, float scale, float zp, U8Norm nm, int64_t total16) {
  __shared__ uint8_t lut[4][256];
  const float rs = __frcp_rn(scale);
  for (int i = threadIdx.x; i < c * 256; i += blockDim.x) {
    const int ch = i >> 8, v = i & 255;
    const float xn = fdiv(fsub(fdiv((float)v, 255.0f), nm.mean[ch]), nm.stdv[ch]);
    lut[ch][v] = (uint8_t)(int8_t)quant_div_guarded(xn, scale, rs, zp);
  }
  __syncthreads();
  const int gw = w / p, gh = h / p;
  const int k = c * p * p;
  const int w16 = w / 16;
  const int n16 = (int)total16;   // 32-bit index arithmetic, as in quant_patchify_kernel
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
    const int xs = i % w16;
    int r = i / w16;
    const int y = r % h;
    r /= h;
    const int ch = r % c;
    const int img = r / c;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)img * c + ch) * h + y) * w + xs * 16));
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = (uint32_t)lut[ch][in[j] & 0xff] | ((uint32_t)lut[ch][(in[j] >> 8) & 0xff] << 8) |
             ((uint32_t)lut[ch][(in[j] >> 16) & 0xff] << 16) | ((uint32_t)lut[ch][in[j] >> 24] << 24);
    const int px = xs * 16;
    const int pw = px / p, kw = px % p, ph = y / p, kh = y % p;
    const int64_t row = ((int64_t)img * gh + ph) * gw + pw;
    *reinterpret_cast<uint4*>(codes + row * k + ((int64_t)ch * p + kh) * p + kw) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- token assembly ------------------------------------------------------------------------------------
// x = qact_embed(cat(cls, patches)) + qact_pos(pos_embed); out = qact1(x)   (models/vit_fquant.py:718-733)
// One thread owns one 4-channel group (blockDim.x is a multiple of d / 4) and walks tokens grid-stride, so the
// per-channel constants and reciprocals are set up once per thread.
__global__ void __launch_bounds__(256)
embed_assemble_kernel(const int8_t* __restrict__ pe, int8_t* __restrict__ out, int b, int np,
                      int d, float pe_scale, float pe_zp, float embed_scale, float embed_zp,
                      const float* __restrict__ cls_value, const float* __restrict__ pos_value,
                      const float* __restrict__ out_scale) {
  const int d4 = d / 4;
  const int c0 = (threadIdx.x % d4) * 4;
  const int tok_per_block = blockDim.x / d4;
  const int64_t tokens = (int64_t)b * (np + 1);
  const float4 so4 = *reinterpret_cast<const float4*>(out_scale + c0);
  const float so[4] = {so4.x, so4.y, so4.z, so4.w};
  const float rso[4] = {__frcp_rn(so4.x), __frcp_rn(so4.y), __frcp_rn(so4.z), __frcp_rn(so4.w)};
  const float4 cl4 = *reinterpret_cast<const float4*>(cls_value + c0);
  const float cls[4] = {cl4.x, cl4.y, cl4.z, cl4.w};
  const float embed_rs = __frcp_rn(embed_scale);
  const float es4[4] = {embed_scale, embed_scale, embed_scale, embed_scale}, ers4[4] = {embed_rs, embed_rs, embed_rs, embed_rs};
  auto clamp_q = [](float r) { return fminf(fmaxf(r, -128.f), 127.f); };
  // 32-bit token arithmetic (b * (np + 1) < 2^31 is checked at launch): a 64-bit modulo per token costs more than
  // the whole element math
  const int ntok = (int)tokens, np1 = np + 1;
  for (int tok = blockIdx.x * tok_per_block + threadIdx.x / d4; tok < ntok; tok += gridDim.x * tok_per_block) {
    const int img = tok / np1;
    const int t = tok - img * np1;
    uint32_t word = 0;
    if (t > 0) word = __ldg(reinterpret_cast<const uint32_t*>(pe + ((int64_t)img * np + (t - 1)) * d + c0));
    const float4 ps4 = __ldg(reinterpret_cast<const float4*>(pos_value + (int64_t)t * d + c0));
    const float pos[4] = {ps4.x, ps4.y, ps4.z, ps4.w};
    float xe[4];
    if (t == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) xe[j] = cls[j];
    } else {   // patch-embedding codes -> qact_embed codes -> values, four channels per guarded rounding
      float pv[4], qe[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pv[j] = fmul(fsub((float)(int8_t)((word >> (8 * j)) & 0xff), pe_zp), pe_scale);
      div_round4(pv, es4, ers4, embed_zp, qe);
#pragma unroll
      for (int j = 0; j < 4; ++j) xe[j] = fmul(fsub(clamp_q(qe[j]), embed_zp), embed_scale);
    }
    float sum[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sum[j] = fadd(xe[j], pos[j]);
    div_round4(sum, so, rso, 0.f, q);
    *reinterpret_cast<uint32_t*>(out + (int64_t)tok * d + c0) = pack_sat4(q[0], q[1], q[2], q[3]);
  }
}

// ---- integer LayerNorm ---------------------------------------------------------------------------------
// One warp per row, grid-stride over rows; each lane owns the same 4-channel groups in every row, so for
// d <= 128 * kRegGroups the per-channel constants are loaded once into registers and reused for all rows
// of the warp.  Power-of-two output grids (every minmax-calibrated model) fold 1/s_out into gamma and
// beta: fl(t*gamma) * 2^-e == fl(t * (gamma * 2^-e)), so the folded form rounds exactly like the
// reference's (t * gamma) / s_out and (beta - u * gamma) / s_out.
constexpr int kLnMaxGroups = 16;  // d <= 2048 (general kernel: the 4C LayerNorm of the last PatchMerging is 1536 wide in
                                  // swin_tiny / small, 2048 in swin_base); the register-resident power-of-two kernel serves d <= 1024

__device__ __forceinline__ uint32_t pack_sat4f(float v0, float v1, float v2, float v3) {
  uint32_t hi, r;
  const int i0 = __float2int_rn(v0), i1 = __float2int_rn(v1), i2 = __float2int_rn(v2), i3 = __float2int_rn(v3);
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(i3), "r"(i2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(i1), "r"(i0), "r"(hi));
  return r;
}

// LN code of one element on a power-of-two grid with pre-folded affine (go = gamma/s_out, bo = beta/s_out).
// Same arithmetic as p2v_math.cuh ln_code<true>, with the dyadic exponent handled in the bit domain:
//   2^N = clamp(2^(7 - e), 1, 2^31),  e = exponent(|A|)   ->  bits(2^(7-e)) = (261 << 23) - (bits(|A|) & 0x7f800000)
__device__ __forceinline__ float ln_code_folded(float xq, const LnRow& row, float go, float bo) {
  const float A = fmul(row.t, go);
  const uint32_t ab = f2u(A) & 0x7fffffffu;
  float p2N = u2f(0x82800000u - (ab & 0x7f800000u));          // 2^(7 - e); e = -127 (A == 0 / denormal) clamps below
  p2N = fminf(fmaxf(p2N, 1.0f), 2147483648.0f);               // N in [0, 31]
  const float inv2N = u2f(0x7f000000u - f2u(p2N));            // 2^-N
  const float M = fminf(floorf(fmul(u2f(ab), p2N)), 255.f);
  // sign(A) * M == copysign(M, A): A == 0 forces M == 0, and (-0) * xq + Bq == Bq
  const float sM = u2f(f2u(M) | (f2u(A) & 0x80000000u));
  const float Bq = rne(fmul(fsub(bo, fmul(row.u, go)), p2N));
  // sM * xq is exact (< 2^24), so the fused multiply-add rounds once, like the reference's separate add
  return rne(fmul(ffma(sM, xq, Bq), inv2N));
}

// The same code with every step on the integer / FMA pipes (ncu: the three FRND per element of ln_code_folded kept
// the conversion unit 58 % busy).  For exponent(|A|) in [-24, 7], i.e. N = 7 - e in [0, 31] without clamping:
//   M = floor(|A| 2^N) is the leading 8 bits of |A|'s significand: bits(sign(A) M) = (bits(A) & 0x807f0000) | 0x43000000
//   2^N and 2^-N are exponent arithmetic, and both roundings are done with the 1.5 * 2^23 constant:
//   RNE(x) = (x + 1.5 * 2^23) - 1.5 * 2^23 for |x| < 2^22, the scaling by the power of two folded into the add.
// ok is cleared when A is out of that range or |Bq| > 2^21 (then |code| <= |Bq| + 255 * 1016 could reach 2^22);
// the caller redoes such elements with ln_code_folded.  Both paths give identical codes where both apply.
__device__ __forceinline__ float ln_code_fast(float xq, const LnRow& row, float go, float bo, bool& ok) {
  constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23
  const float A = fmul(row.t, go);
  const uint32_t fa = f2u(A);
  const uint32_t ex = fa & 0x7f800000u;
  const uint32_t p2n = 0x82800000u - ex;                      // bits of 2^(7 - e)
  const float sM = u2f((fa & 0x807f0000u) | 0x43000000u);
  const float b = fsub(bo, fmul(row.u, go));
  const float Bq = fsub(ffma(b, u2f(p2n), kMagic), kMagic);   // RNE(b 2^N): b 2^N is exact, one rounding in the fma
  ok = ok & (ex - 0x33800000u < 0x10000000u) & (fabsf(Bq) <= 2097152.0f);   // no short circuit: branch-free
  const float y = ffma(sM, xq, Bq);                            // exact product (< 2^18) + integer: one rounding
  return fsub(ffma(y, u2f(0x7f000000u - p2n), kMagic), kMagic);   // RNE(y 2^-N)
}

// Two adjacent elements at once with the packed fp32 instructions of sm_100 (fma.rn.f32x2 and friends: each half
// rounds exactly like the scalar _rn op, so the codes are those of ln_code_fast); the bit-domain steps stay scalar.
__device__ __forceinline__ void ln_code_fast2(const float (&xq)[2], const LnRow& row, const float (&go)[2],
                                              const float (&bo)[2], bool& ok, float (&code)[2]) {
  const float2 kMagic = make_float2(12582912.0f, 12582912.0f), kMagicNeg = make_float2(-12582912.0f, -12582912.0f);
  const float2 g2 = make_float2(go[0], go[1]);
  const float2 A = fmul2(make_float2(row.t, row.t), g2);
  const uint32_t fa0 = f2u(A.x), fa1 = f2u(A.y);
  const uint32_t ex0 = fa0 & 0x7f800000u, ex1 = fa1 & 0x7f800000u;
  const uint32_t p0 = 0x82800000u - ex0, p1 = 0x82800000u - ex1;               // bits of 2^(7 - e)
  const float2 sM = make_float2(u2f((fa0 & 0x807f0000u) | 0x43000000u), u2f((fa1 & 0x807f0000u) | 0x43000000u));
  // bo - u go with two roundings as in the reference: packed product, scalar subtractions (a packed add fed by a
  // packed mul gets contracted into one FFMA2 by ptxas, explicit .rn or not)
  const float2 ug = fmul2(make_float2(row.u, row.u), g2);
  // bo - ug as ug * (-1) + bo: one packed instruction, one rounding of the already rounded product
  const float2 b = ffma2(ug, make_float2(-1.0f, -1.0f), make_float2(bo[0], bo[1]));
  const float2 Bq = fadd2(ffma2(b, make_float2(u2f(p0), u2f(p1)), kMagic), kMagicNeg);
  ok = ok & (ex0 - 0x33800000u < 0x10000000u) & (ex1 - 0x33800000u < 0x10000000u) &
       (fabsf(Bq.x) <= 2097152.0f) & (fabsf(Bq.y) <= 2097152.0f);
  const float2 y = ffma2(sM, make_float2(xq[0], xq[1]), Bq);
  const float2 c = fadd2(ffma2(y, make_float2(u2f(0x7f000000u - p0), u2f(0x7f000000u - p1)), kMagic), kMagicNeg);
  code[0] = c.x;
  code[1] = c.y;
}

// byte J of w, sign-extended, in one PRMT (selector nibble 8 | J replicates the byte's msb; __byte_perm would mask
// that bit away, hence the PTX)
__device__ __forceinline__ int sext_byte(uint32_t w, int j) {
  const uint32_t sel = j == 0 ? 0x8880u : (j == 1 ? 0x9991u : (j == 2 ? 0xaaa2u : 0xbbb3u));   // constant after unrolling
  int v;
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(w), "r"(sel));
  return v;
}

// G = 4-channel groups per lane.  FULL: d == 128 * G (no partial group).  DUMP: also write the unclamped LN codes.
// SM: the per-channel constants live in shared memory (4 x 128 G floats, dynamic) instead of registers - for
// d > 384 (DeiT-B / ViT-B: 768), where 16 G constant registers per lane no longer fit beside the row.
// LPR = lanes per row: narrow rows (Swin's d = 96 / 192 = 4 * 3 * LPR) put 32 / LPR rows side by side in a warp, so the
// per-row chain (statistics, three IEEE divisions, the square root) serves that many rows per issue slot; the row sums
// are then reduced by shuffles inside each LPR-lane segment.  LPR = 32 is the one-row-per-warp kernel.
template <int G, bool FULL, bool DUMP, bool SM = false, int LPR = 32>
__global__ void __launch_bounds__(256, 2)
layernorm_int_pot_kernel(const int8_t* __restrict__ in, int64_t in_row_stride, int8_t* __restrict__ out,
                         int32_t* __restrict__ ln_codes, int rows, int d, const p2v_layernorm p) {
  pdl_launch_dependents();
  constexpr int RPW = 32 / LPR;                         // rows per warp and step
  const int lane = LPR == 32 ? (threadIdx.x & 31) : (threadIdx.x & (LPR - 1));   // lane within its row
  const int rsel = LPR == 32 ? 0 : ((threadIdx.x & 31) / LPR);
  const int warps_total = gridDim.x * (blockDim.x >> 5) * RPW;
  const int groups = d >> 2;
  constexpr int GR = SM ? 1 : G;     // register copies only without SM
  float go[GR][4], bo[GR][4], pm[GR][4];
  int mk[GR][4];
  extern __shared__ __align__(16) float ln_const[];   // SM: go | bo | pm | mask, 128 G entries each, zero beyond d
  if (SM) {
    for (int c = threadIdx.x; c < 128 * G; c += blockDim.x) {
      const bool in_range = c < d;
      const float rs = in_range ? p.ln_out_rscale[c] : 0.f;
      ln_const[c] = in_range ? fmul(p.gamma[c], rs) : 0.f;
      ln_const[128 * G + c] = in_range ? fmul(p.beta[c], rs) : 0.f;
      ln_const[2 * 128 * G + c] = in_range ? p.post_mul[c] : 0.f;
      ln_const[3 * 128 * G + c] = in_range ? p.in_mask[c] : 0.f;
    }
    __syncthreads();
  }
#pragma unroll
  for (int g = 0; g < GR; ++g) {
    const int grp = g * LPR + lane;
    if (SM) {
    } else if (FULL || grp < groups) {
      const int c0 = grp * 4;
      const float4 ga = *reinterpret_cast<const float4*>(p.gamma + c0);
      const float4 be = *reinterpret_cast<const float4*>(p.beta + c0);
      const float4 rs = *reinterpret_cast<const float4*>(p.ln_out_rscale + c0);
      const float4 m4 = *reinterpret_cast<const float4*>(p.post_mul + c0);
      const float4 im = *reinterpret_cast<const float4*>(p.in_mask + c0);
      go[g][0] = fmul(ga.x, rs.x); go[g][1] = fmul(ga.y, rs.y); go[g][2] = fmul(ga.z, rs.z); go[g][3] = fmul(ga.w, rs.w);
      bo[g][0] = fmul(be.x, rs.x); bo[g][1] = fmul(be.y, rs.y); bo[g][2] = fmul(be.z, rs.z); bo[g][3] = fmul(be.w, rs.w);
      pm[g][0] = m4.x; pm[g][1] = m4.y; pm[g][2] = m4.z; pm[g][3] = m4.w;
      mk[g][0] = (int)im.x; mk[g][1] = (int)im.y; mk[g][2] = (int)im.z; mk[g][3] = (int)im.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { go[g][j] = 0.f; bo[g][j] = 0.f; pm[g][j] = 0.f; mk[g][j] = 0; }
    }
  }
  const float scale_over_c = fdiv(p.in_scale1, (float)d);   // row-independent part of ln_row_stats
  pdl_wait();   // everything above is static; the rows are the previous kernel's output
  // software pipeline: the next row's codes are in flight while this row is normalised
  uint32_t next_w[G];
  int wrow = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;   // first row of this warp's step (uniform)
  int row = wrow + rsel;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int grp = g * LPR + lane;
    next_w[g] = (row < rows && (FULL || grp < groups)) ? ld_act_u32(in + (int64_t)row * in_row_stride + grp * 4) : 0u;
  }
  for (; wrow < rows; wrow += warps_total, row += warps_total) {
    uint32_t cur_w[G];
    const int nrow = row + warps_total;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int grp = g * LPR + lane;
      cur_w[g] = next_w[g];
      next_w[g] = (nrow < rows && (FULL || grp < groups)) ? ld_act_u32(in + (int64_t)nrow * in_row_stride + grp * 4) : 0u;
    }
    float xq[G][4];
    int sum = 0, sumsq = 0;   // |x| <= 1024, d <= 128 G: per-lane partial sums stay far below 2^31; padded groups add 0
#pragma unroll
    for (int g = 0; g < G; ++g) {
      int mg[4];
      if (SM) {
        const float4 m4 = *reinterpret_cast<const float4*>(ln_const + 3 * 128 * G + (g * LPR + lane) * 4);
        mg[0] = (int)m4.x; mg[1] = (int)m4.y; mg[2] = (int)m4.z; mg[3] = (int)m4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) mg[j] = mk[SM ? 0 : g][j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // one PRMT sign-extends byte j (selector nibble 8 | j replicates its msb)
        const int v = sext_byte(cur_w[g], j) * mg[j];
        xq[g][j] = (float)v;
        sum += v;
        sumsq += v * v;
      }
    }
    // warp totals with the REDUX unit (one instruction per 32-bit value instead of five shuffle / add rounds); the
    // sum of squares can pass 2^32 for large masks, so its two 16-bit halves are reduced separately
    long long sumsq64;
    if (LPR == 32) {
      sum = __reduce_add_sync(0xffffffffu, sum);
      const unsigned sq_lo = __reduce_add_sync(0xffffffffu, (unsigned)sumsq & 0xffffu);
      const unsigned sq_hi = __reduce_add_sync(0xffffffffu, (unsigned)sumsq >> 16);
      sumsq64 = ((long long)sq_hi << 16) + (long long)sq_lo;
    } else {
      sumsq64 = (long long)(unsigned)sumsq;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {   // xor offsets below LPR stay inside the row's lane segment
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sumsq64 += __shfl_xor_sync(0xffffffffu, sumsq64, o);
      }
    }
    const LnRow st = ln_row_stats((long long)sum, sumsq64, d, p.in_scale1, scale_over_c);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int grp = g * LPR + lane;
      if ((FULL || grp < groups) && (LPR == 32 || row < rows)) {
        float v[4], code[4], gg[4], bb[4], pp[4];
        if (SM) {
          const float4 g4 = *reinterpret_cast<const float4*>(ln_const + grp * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(ln_const + 128 * G + grp * 4);
          const float4 p4 = *reinterpret_cast<const float4*>(ln_const + 2 * 128 * G + grp * 4);
          gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
          bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
          pp[0] = p4.x; pp[1] = p4.y; pp[2] = p4.z; pp[3] = p4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) { gg[j] = go[SM ? 0 : g][j]; bb[j] = bo[SM ? 0 : g][j]; pp[j] = pm[SM ? 0 : g][j]; }
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float x2[2] = {xq[g][j], xq[g][j + 1]}, g2[2] = {gg[j], gg[j + 1]}, b2[2] = {bb[j], bb[j + 1]};
          float c2[2];
          ln_code_fast2(x2, st, g2, b2, ok, c2);
          code[j] = c2[0];
          code[j + 1] = c2[1];
        }
        if (!ok) {   // rare: a dyadic exponent outside [0, 31] before clamping, or a huge offset
#pragma unroll
          for (int j = 0; j < 4; ++j) code[j] = ln_code_folded(xq[g][j], st, gg[j], bb[j]);
        }
        if (DUMP) {
#pragma unroll
          for (int j = 0; j < 4; ++j) ln_codes[(int64_t)row * d + grp * 4 + j] = (int)code[j];
        }
        if (p.pre_clamp) {   // an int8 QAct on the LayerNorm's own grid sits in front of the re-gridding (Swin)
#pragma unroll
          for (int j = 0; j < 4; ++j) code[j] = fminf(fmaxf(code[j], -128.f), 127.f);
        }
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          // code * 2^k is exact: one rounding, like mul then add
          const float2 v2 = ffma2(make_float2(code[j], code[j + 1]), make_float2(pp[j], pp[j + 1]),
                                       make_float2(p.post_zp, p.post_zp));
          v[j] = v2.x;
          v[j + 1] = v2.y;
        }
        *reinterpret_cast<uint32_t*>(out + (int64_t)row * d + grp * 4) = pack_sat4f(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

// General path (non-power-of-two grids, or d > 128 * 3): constants re-read per row, IEEE divisions.
template <bool POT>
__global__ void __launch_bounds__(256)
layernorm_int_kernel(const int8_t* __restrict__ in, int64_t in_row_stride, int8_t* __restrict__ out,
                     int32_t* __restrict__ ln_codes, int rows, int d, const p2v_layernorm p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int groups = d >> 2;  // 4-channel groups in the row
  const int8_t* src = in + (int64_t)row * in_row_stride;
  float xq[kLnMaxGroups][4];
  int sum = 0;
  long long sumsq = 0;
#pragma unroll
  for (int g = 0; g < kLnMaxGroups; ++g) {
    const int grp = g * 32 + lane;
    if (grp < groups) {
      const uint32_t word = *reinterpret_cast<const uint32_t*>(src + grp * 4);
      const float4 mk = *reinterpret_cast<const float4*>(p.in_mask + grp * 4);
      const float m[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = (int)(int8_t)((word >> (8 * j)) & 0xff) * (int)m[j];
        xq[g][j] = (float)v;
        sum += v;
        sumsq += (long long)(v * v);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
  }
  const LnRow st = ln_row_stats((long long)sum, sumsq, d, p.in_scale1);
#pragma unroll
  for (int g = 0; g < kLnMaxGroups; ++g) {
    const int grp = g * 32 + lane;
    if (grp < groups) {
      const int c0 = grp * 4;
      const float4 ga = *reinterpret_cast<const float4*>(p.gamma + c0);
      const float4 be = *reinterpret_cast<const float4*>(p.beta + c0);
      const float4 os = *reinterpret_cast<const float4*>(p.ln_out_scale + c0);
      const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
      const float osc[4] = {os.x, os.y, os.z, os.w};
      float ors[4] = {0.f, 0.f, 0.f, 0.f}, pm[4] = {1.f, 1.f, 1.f, 1.f}, pd[4] = {1.f, 1.f, 1.f, 1.f};
      if (POT) {
        const float4 r4 = *reinterpret_cast<const float4*>(p.ln_out_rscale + c0);
        const float4 m4 = *reinterpret_cast<const float4*>(p.post_mul + c0);
        ors[0] = r4.x; ors[1] = r4.y; ors[2] = r4.z; ors[3] = r4.w;
        pm[0] = m4.x; pm[1] = m4.y; pm[2] = m4.z; pm[3] = m4.w;
      } else {
        const float4 d4 = *reinterpret_cast<const float4*>(p.post_div1 + c0);
        pd[0] = d4.x; pd[1] = d4.y; pd[2] = d4.z; pd[3] = d4.w;
      }
      int q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float code = ln_code<POT>(xq[g][j], st, gam[j], bet[j], osc[j], ors[j]);
        if (ln_codes != nullptr) ln_codes[(int64_t)row * d + c0 + j] = (int)code;
        if (p.pre_clamp) code = fminf(fmaxf(code, -128.f), 127.f);
        float v;
        if (POT) {
          v = rne(fadd(fmul(code, pm[j]), p.post_zp));
        } else {
          // value on the LN grid, / SmoothQuant scale of the consumer, / its QAct scale (+ zp)
          v = rne(fadd(fdiv(fdiv(fmul(code, osc[j]), pd[j]), p.post_div2), p.post_zp));
        }
        q[j] = clamp_i(v, -128, 127);
      }
      *reinterpret_cast<uint32_t*>(out + (int64_t)row * d + c0) = pack4i(q[0], q[1], q[2], q[3]);
    }
  }
}

// ---- standalone QAct on fp32 tensors -------------------------------------------------------------------
__global__ void fake_quant_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int8_t* __restrict__ codes,
                                      int64_t total, int channels, int64_t inner, const float* __restrict__ scale,
                                      const float* __restrict__ zero_point, int qmin, int qmax) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)((i / inner) % channels);
    const float s = scale[ch], z = zero_point ? zero_point[ch] : 0.f;
    const int q = quant_div(x[i], s, z, qmin, qmax);
    if (codes != nullptr) codes[i] = (int8_t)q;
    if (out != nullptr) out[i] = fmul(fsub((float)q, z), s);
  }
}

int quant_patchify_small_launch(const float* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                float zero_point, cudaStream_t st);   // p2v_swin.cu
int quant_patchify_small_u8_launch(const uint8_t* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                   float zero_point, const float* mean, const float* stdv, cudaStream_t st);

static int grid_for(int64_t work, int block) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_quant_patchify(const float* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                  float zero_point, void* stream) {
  P2V_REQUIRE(x && codes, "p2v_quant_patchify: null pointer");
  P2V_REQUIRE(b > 0 && c > 0 && p > 0 && h % p == 0 && w % p == 0, "p2v_quant_patchify: bad shape %dx%dx%dx%d p=%d", b,
              c, h, w, p);
  if (p % 16 != 0 || w % 16 != 0)   // small patches (Swin: 4 x 4): four pixels per thread, csrc/p2v_swin.cu
    return quant_patchify_small_launch(x, codes, b, c, h, w, p, scale, zero_point, (cudaStream_t)stream);
  const int64_t total16 = (int64_t)b * c * h * (w / 16);
  P2V_REQUIRE(total16 < (1ll << 31) - (int64_t)kNumSMs * 16 * 256, "p2v_quant_patchify: batch too large for 32-bit pixel indexing");
  quant_patchify_kernel<<<grid_for(total16, 256), 256, 0, (cudaStream_t)stream>>>(x, codes, b, c, h, w, p, scale,
                                                                                  zero_point, total16);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_quant_patchify_u8(const uint8_t* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                     float zero_point, const float* mean, const float* stdv, void* stream) {
  P2V_REQUIRE(x && codes && mean && stdv, "p2v_quant_patchify_u8: null pointer");
  P2V_REQUIRE(b > 0 && c > 0 && c <= 4 && p > 0 && h % p == 0 && w % p == 0,
              "p2v_quant_patchify_u8: bad shape %dx%dx%dx%d p=%d (at most 4 channels)", b, c, h, w, p);
  if (p % 16 != 0 || w % 16 != 0)   // small patches (Swin: 4 x 4): four pixels per thread, csrc/p2v_swin.cu
    return quant_patchify_small_u8_launch(x, codes, b, c, h, w, p, scale, zero_point, mean, stdv, (cudaStream_t)stream);
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "p2v_quant_patchify_u8: x must be 16-byte aligned");
  U8Norm nm = {};
  for (int i = 0; i < c; ++i) {
    nm.mean[i] = mean[i];
    nm.stdv[i] = stdv[i];
  }
  const int64_t total16 = (int64_t)b * c * h * (w / 16);
  P2V_REQUIRE(total16 < (1ll << 31) - (int64_t)kNumSMs * 16 * 256, "p2v_quant_patchify_u8: batch too large for 32-bit pixel indexing");
  quant_patchify_u8_kernel<<<grid_for(total16, 256), 256, 0, (cudaStream_t)stream>>>(x, codes, b, c, h, w, p, scale,
                                                                                     zero_point, nm, total16);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_embed_assemble(const int8_t* pe, int8_t* out, int b, int np, int d, float pe_scale, float pe_zp,
                                  float embed_scale, float embed_zp, const float* cls_value, const float* pos_value,
                                  const float* out_scale, void* stream) {
  P2V_REQUIRE(pe && out && cls_value && pos_value && out_scale, "p2v_embed_assemble: null pointer");
  P2V_REQUIRE(b > 0 && np > 0 && d > 0 && d % 4 == 0, "p2v_embed_assemble: bad shape b=%d np=%d d=%d", b, np, d);
  const int d4 = d / 4;
  P2V_REQUIRE(d4 <= 256, "p2v_embed_assemble: d=%d exceeds 1024 channels", d);
  P2V_REQUIRE((int64_t)b * (np + 1) < (1ll << 31), "p2v_embed_assemble: too many tokens");
  const int block = (256 / d4) * d4;   // a whole number of tokens per block
  const int64_t tokens = (int64_t)b * (np + 1);
  embed_assemble_kernel<<<grid_for(tokens, block / d4), block, 0, (cudaStream_t)stream>>>(
      pe, out, b, np, d, pe_scale, pe_zp, embed_scale, embed_zp, cls_value, pos_value, out_scale);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_layernorm_int(const int8_t* in, int64_t in_row_stride, int8_t* out, int32_t* ln_codes, int rows,
                                 int d, const p2v_layernorm* p, void* stream) {
  P2V_REQUIRE(in && out && p, "p2v_layernorm_int: null pointer");
  P2V_REQUIRE(rows > 0 && d > 0 && d % 4 == 0 && d <= 128 * kLnMaxGroups, "p2v_layernorm_int: bad shape rows=%d d=%d",
              rows, d);
  P2V_REQUIRE(in_row_stride % 4 == 0, "p2v_layernorm_int: row stride must be a multiple of 4 bytes");
  P2V_REQUIRE(p->in_mask && p->gamma && p->beta && p->ln_out_scale, "p2v_layernorm_int: missing vectors");
  const int warps = 8;
  const int grid = (rows + warps - 1) / warps;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->pot && d > 1024) {   // wider than the register / shared-memory resident kernel (Swin's last PatchMerging: 4 x 384)
    P2V_REQUIRE(p->ln_out_rscale && p->post_mul, "p2v_layernorm_int: pot path needs ln_out_rscale and post_mul");
    layernorm_int_kernel<true><<<grid, warps * 32, 0, st>>>(in, in_row_stride, out, ln_codes, rows, d, *p);
  } else if (p->pot) {
    P2V_REQUIRE(p->ln_out_rscale && p->post_mul, "p2v_layernorm_int: pot path needs ln_out_rscale and post_mul");
    const int groups = (d / 4 + 31) / 32;
    const int pgrid = grid < kNumSMs * 2 ? grid : kNumSMs * 2;   // persistent warps (2 resident CTAs per SM, 119 registers: three CTAs at 80 registers spilled and were slower): constants stay in registers
#define P2V_LN_LAUNCH(G_)                                                                                          \
  do {                                                                                                             \
    const bool full = d == 128 * (G_);                                                                             \
    if (full && ln_codes) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, true, true>, dim3(pgrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));        \
    else if (full) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, true, false>, dim3(pgrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));            \
    else if (ln_codes) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, false, true>, dim3(pgrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));        \
    else P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, false, false>, dim3(pgrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));                     \
  } while (0)
#define P2V_LN_LAUNCH_SM(G_)                                                                                       \
  do {                                                                                                             \
    const size_t smem = 4 * 128 * (G_) * sizeof(float);                                                            \
    if (ln_codes) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, false, true, true>, dim3(pgrid), dim3(warps * 32), smem, st, in, in_row_stride, out, ln_codes, rows, d, *p)); \
    else P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<G_, false, false, true>, dim3(pgrid), dim3(warps * 32), smem, st, in, in_row_stride, out, ln_codes, rows, d, *p));        \
  } while (0)
    // rows of 96 / 192 channels (Swin stages 1 and 2): four / two rows per warp
    if (d == 96 || d == 192) {
      const int lpr = d / 12, rgrid = (rows + warps * (32 / lpr) - 1) / (warps * (32 / lpr));
      const int ngrid = rgrid < kNumSMs * 2 ? rgrid : kNumSMs * 2;
      if (lpr == 8 && ln_codes) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<3, true, true, false, 8>, dim3(ngrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));
      else if (lpr == 8) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<3, true, false, false, 8>, dim3(ngrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));
      else if (ln_codes) P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<3, true, true, false, 16>, dim3(ngrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));
      else P2V_CHECK_CUDA(launch_pdl(2, layernorm_int_pot_kernel<3, true, false, false, 16>, dim3(ngrid), dim3(warps * 32), 0, st, in, in_row_stride, out, ln_codes, rows, d, *p));
    } else
    if (groups == 1) P2V_LN_LAUNCH(1);
    else if (groups == 2) P2V_LN_LAUNCH(2);
    else if (groups == 3) P2V_LN_LAUNCH(3);
    else if (groups <= 4) P2V_LN_LAUNCH_SM(4);
    else if (groups <= 6) P2V_LN_LAUNCH_SM(6);      // d = 768: DeiT-B / ViT-B
    else P2V_LN_LAUNCH_SM(8);                       // d <= 1024
#undef P2V_LN_LAUNCH
#undef P2V_LN_LAUNCH_SM
  } else {
    P2V_REQUIRE(p->post_div1, "p2v_layernorm_int: non-pot path needs post_div1");
    layernorm_int_kernel<false><<<grid, warps * 32, 0, st>>>(in, in_row_stride, out, ln_codes, rows, d, *p);
  }
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_fake_quant_f32(const float* x, float* out, int8_t* codes, int64_t outer, int channels,
                                  int64_t inner, const float* scale, const float* zero_point, int qmin, int qmax,
                                  void* stream) {
  P2V_REQUIRE(x && (out || codes) && scale, "p2v_fake_quant_f32: null pointer");
  P2V_REQUIRE(outer > 0 && channels > 0 && inner > 0, "p2v_fake_quant_f32: bad shape");
  const int64_t total = outer * channels * inner;
  fake_quant_f32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, codes, total, channels, inner,
                                                                                 scale, zero_point, qmin, qmax);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
