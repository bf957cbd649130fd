// p2v_attention.cu - fused integer attention for one layer:
//   S = Q K^T (int32) -> qact_attn1 int8 codes -> log-int-softmax 4-bit log2 codes -> P V -> qact2 codes
// with the [n, n] score matrix never leaving the SM.
//
// Replaces models/vit_fquant.py:308-326 and QIntSoftmax.forward (models/ptq/layers.py:323-376).
//
// Integer formulation (probed bit-exact against the reference):
//   score code   sc = clamp(RNE(acc * 2^(2 e_qkv - 3 - e_score)))          acc = sum_d q*k   (|acc| < 2^20)
//   integer exp  e(d) for d = rowmax - sc in [0, 255]: a 256-entry table built on the host with the
//                reference's own expressions; the row sum is exact (integers < 2^51 accumulated in fp64)
//   log2 code    k = log_round(RNE(sum / e)) in [0, 15], 16 = probability 0
//   AV           acc2 = sum_j v_j * 2^(15 - k_j)  (int32), out = clamp(RNE(acc2 * 2^(e_qkv - 15 - e_out)))
// The probabilities 2^(15-k) span 16 bits, so P is split into two u8 planes (k <= 7 -> 2^(7-k) in units of
// 2^8; k >= 8 -> 2^(15-k)) and the AV product is two u8 x s8 tensor-core MMAs recombined as (hi << 8) + lo.
//
// One CTA = one (image, head), K/V staged once; each of its 7 warps walks 16-row query tiles, all keys; three CTAs
// per SM.  Both products use mma.sync m16n8k32 (IMMA): the score fragment layout (row g, cols 2t,2t+1 of each 8-key
// tile) is re-used directly as the A operand of the AV product by permuting the key order of V when it is transposed
// into shared memory, so the 4-bit codes never leave the SM.  The kernel is bound by CUDA-core issue and
// shared-memory wavefronts of the softmax code path, not by the tensor pipe (profiles/r1_summary.md items 11-12, 19);
// what each phase does to stay cheap is described where it happens.
#include <math.h>

#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

constexpr int kHd = 64;            // head dim (all DeiT/ViT configs of the reference)
constexpr int kAttWarps = 7;
constexpr int kAttRows = kAttWarps * 16;  // 112 query rows per CTA
constexpr int kMaxKeys = 224;      // keys padded to a multiple of 32
constexpr int kMaxTiles = kMaxKeys / 8;
constexpr int kKStride = 64;       // bytes per K row in smem; the 16 words of a row are stored t-major (see stage K)
constexpr int kVtStride = 240;     // bytes per V^T row in smem (224 + pad)

__device__ __forceinline__ void mma_s8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A B + C with C a broadcast constant kept in its own registers (the scores start from a bias, see below)
__device__ __forceinline__ void mma_s8s8_init(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int c) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c));
}
// RNE + unsigned saturation of two floats into bytes 0 and 1, `upper`'s low half into bytes 2 and 3
__device__ __forceinline__ uint32_t pack2_u8(float lo, float hi, uint32_t upper) {
  uint32_t r;
  const int a = __float2int_rn(hi), b = __float2int_rn(lo);
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(upper));
  return r;
}
__device__ __forceinline__ uint32_t pack2_s8(int lo, int hi) {
  uint32_t r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(lo), "r"(0));
  return r;
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Table lookups of the hot loops through 32-bit shared-window addresses: base + (byte << k) is a single LEA,
// where indexing a generic pointer costs a subtract, a scaled add and the window base.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// position of key j inside the permuted K dimension of the AV product (see file header)
__device__ __forceinline__ int av_perm(int j) {
  const int s = j >> 5, jj = j & 31;
  const int ti = jj >> 3, col = jj & 7;
  const int t = col >> 1, i0 = col & 1;
  const int h = ti >> 1, i = ((ti & 1) << 1) | i0;
  return (s << 5) + (h << 4) + (t << 2) + i;
}

// Fast evaluation of the log2 code k = log_round(RNE(S / e)) (p2v_math.cuh softmax_log_code).
// With y = S/e + 1/2 the code is a step function of y with steps at 2, 3, 6, 12, 24, ...:
//   k = floor(log2(y / 3)) + 2 for y >= 3, and the probability in units of 2^-15 is 1 << (15 - k).
// u = fma(S, fl(1/(3e)), 1/6) approximates y/3 to a few ulp, so k = exponent(u) + 2 unless u lies within
// kCodeGuard ulps of a power of two (a step) or below 1 (y < 3: a key holding more than 40 % of the row
// mass, where the one irregular step y = 2 lives).  Flagged elements (~1e-5 of all plus the dominant keys
// of peaked rows) are recomputed with the exact IEEE-division formula; both paths give the same code.
constexpr uint32_t kCodeGuard = 64;  // ulps: |u - 2^j| / 2^j < 2^-17, far above the ~4 ulp error of u

__device__ __noinline__ uint32_t exact_prob16(float fsum, float e, int levels) {
  const int k = softmax_log_code(fsum, e, levels);
  return k >= 16 ? 0u : (0x8000u >> k);
}

// guard test of one element: next to a step of the code function, or below 1 (dominant key of a peaked row)
__device__ __forceinline__ bool code_guarded(uint32_t bits) {
  return (((bits + kCodeGuard) & 0x7fffffu) < 2 * kCodeGuard) || bits < 0x3f800000u;
}

// four keys of one packed code word -> four 16-bit probabilities.  Returns the minimum over the four elements of
//   (int)((bits - 0x3f800000 + kCodeGuard) & 0x807fff80):  <= 0 exactly when code_guarded() holds for one of them
// (negative: u < 1; zero: within kCodeGuard ulps of a power of two), so the caller tests one number per word pair.
// revR = rev_r3 + (255 - rowmax) (biased), so the table is addressed by the code byte directly.
__device__ __forceinline__ int prob16x4(uint32_t w, uint32_t revR, float fsum, uint32_t (&v)[4], uint32_t (&bits)[4]) {
  static_assert(kCodeGuard == 64, "the mask below encodes a 64-ulp guard");
  int worst = 0x7fffffff;
  const float2 fs2 = make_float2(fsum, fsum), sixth2 = make_float2(0.16666667f, 0.16666667f);
#pragma unroll
  for (int i = 0; i < 4; i += 2) {
    const float r0 = lds_f32(revR + (__byte_perm(w, 0, 0x4440 + i) << 2));
    const float r1 = lds_f32(revR + (__byte_perm(w, 0, 0x4441 + i) << 2));
    const float2 u = ffma2(fs2, make_float2(r0, r1), sixth2);   // two keys per fma.rn.f32x2
    bits[i] = __float_as_uint(u.x);
    bits[i + 1] = __float_as_uint(u.y);
#pragma unroll
    for (int k = i; k < i + 2; ++k) {
      worst = min(worst, (int)((bits[k] - 0x3f7fffc0u) & 0x807fff80u));
      // 1 << (15 - k), k = E - 125: shift counts >= 32 (k >= 16, or the wrapped negative) give 0
      asm("shl.b32 %0, %1, %2;" : "=r"(v[k]) : "r"(1u), "r"(140u - (bits[k] >> 23)));
    }
  }
  return worst;
}

struct AttSmem {
  alignas(16) uint8_t Ks[kMaxKeys * kKStride];
  alignas(16) uint8_t Vt[kHd * kVtStride];
  alignas(16) uint8_t codes[kAttWarps][16 * kVtStride];   // biased score codes in AV key order, per warp
  float lut_f[256];    // e(d), d = rowmax - code (exact path)
  // the two tables of the hot loops are stored reversed (entry 255 - d), so that with the per-row base
  // rev + (255 - rowmax) they are addressed by the biased code byte itself: one LEA per lookup
  float rev_r3[256];   // 1 / (3 e(d))
  int zks[kMaxKeys];   // kZp: z * sum_d k[j][d] per key (asymmetric q/k/v codes)
  uint32_t rev_ehi[256];   // high word of (double)e(d) when its low word is zero (e has <= 21 significant bits: it is
                           // z * 2^(32-q) with z a 10..20-bit integer), so the row sums need no fp32 -> fp64 convert
  int wide;                // some entry has more significant bits: the sums convert rev_e instead
  float rev_e[256];    // e(d): an integer of <= 24 significant bits, exact in fp32 (4-byte entries: a 32-lane lookup
                       // spreads over all 32 banks, where fp64 entries left 16 bank pairs and twice the conflicts)
};

// One CTA = one (image, head): K and V are staged once, each of the 7 warps walks 16-row query tiles.
#ifndef P2V_ATT_MIN_CTAS
#define P2V_ATT_MIN_CTAS 3
#endif
// kPot: the score multiplier is a power of two (every minmax-calibrated model), which lets the int32 -> fp32
// conversion of the scores ride on the accumulator (see the score loop).
// kZp: the q/k/v codes carry a zero point z (asymmetric observers).  The MMAs still multiply the raw codes;
//   sum (q - z)(k - z) = acc - z (sum q + sum k) + 64 z^2   and   sum p (v - z) = acc2 - z sum p
// are completed with per-row / per-key sums (kZp implies !kPot: the corrected sum can reach 2^22).
template <bool kDump, bool kPot, bool kZp>
__global__ void __launch_bounds__(kAttWarps * 32, P2V_ATT_MIN_CTAS)
attention_int_kernel(const int8_t* __restrict__ qkv, int8_t* __restrict__ out, int n, int heads,
                     const p2v_attention p, int out_shift) {
  extern __shared__ __align__(16) uint8_t att_smem_raw[];
  AttSmem& sm = *reinterpret_cast<AttSmem*>(att_smem_raw);

  const int bh = blockIdx.x;
  const int img = bh / heads, head = bh % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nkp = (n + 31) & ~31;   // padded key count
  const int ntiles = nkp >> 3;
  const int64_t row_stride = (int64_t)3 * heads * kHd;
  const int8_t* base = qkv + (int64_t)img * n * row_stride;
  const int iz = kZp ? (int)p.in_zp : 0;

  // ---- stage K (row-major, padded stride) and V (transposed + permuted) in shared memory ---------------
  if (tid == 0) sm.wide = 0;
  __syncthreads();
  for (int i = tid; i < 256; i += blockDim.x) {
    const float e = p.exp_lut[i];
    sm.lut_f[i] = e;
    sm.rev_r3[255 - i] = __fdiv_rn(1.0f, 3.0f * e);   // 3e is exact (<= 24 significant bits)
    sm.rev_e[255 - i] = e;
    sm.rev_ehi[255 - i] = (uint32_t)__double2hiint((double)e);
    if (__double2loint((double)e) != 0) sm.wide = 1;
  }
  // K rows with their 16 words stored t-major: word w (columns 4w .. 4w+3) goes to slot (w & 3) * 4 + (w >> 2), so
  // the four B-fragment words of lane t (columns 4t, 16 + 4t, 32 + 4t, 48 + 4t) are one 16-byte load
  for (int i = tid; i < nkp * 4; i += blockDim.x) {
    const int j = i >> 2, part = i & 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < n) v = __ldg(reinterpret_cast<const uint4*>(base + j * row_stride + (heads + head) * kHd) + part);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm.Ks + j * kKStride) + part;
    dst[0] = v.x;
    dst[4] = v.y;
    dst[8] = v.z;
    dst[12] = v.w;
    if (kZp) {   // the four lanes of a key add their 16-byte partial sums
      int ks = __dp4a((int)v.x, 0x01010101, __dp4a((int)v.y, 0x01010101, __dp4a((int)v.z, 0x01010101, __dp4a((int)v.w, 0x01010101, 0))));
      const uint32_t mask = __activemask();
      ks += __shfl_xor_sync(mask, ks, 1);
      ks += __shfl_xor_sync(mask, ks, 2);
      if (part == 0) sm.zks[j] = iz * ks;
    }
  }
  // V^T with the key permutation of the AV product: 4 consecutive positions kappa = 16h + 4t + {0,1,2,3}
  // hold keys {j0, j0+1, j0+8, j0+9}, j0 = 32s + 16h + 2t.  One thread transposes a 4-key x 4-channel
  // block with byte permutes and writes four 32-bit words.
  for (int i = tid; i < (nkp >> 2) * (kHd >> 2); i += blockDim.x) {
    const int c4 = i & 15, kq = i >> 4;                 // channel quad, key quad
    const int s5 = kq >> 3, h = (kq >> 2) & 1, tt = kq & 3;
    const int j0 = (s5 << 5) + (h << 4) + (tt << 1);
    uint32_t r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + (q >> 1) * 8 + (q & 1);
      r[q] = j < n ? __ldg(reinterpret_cast<const uint32_t*>(base + j * row_stride + (2 * heads + head) * kHd) + c4) : 0u;
    }
    const uint32_t lo01 = __byte_perm(r[0], r[1], 0x5140), hi01 = __byte_perm(r[0], r[1], 0x7362);
    const uint32_t lo23 = __byte_perm(r[2], r[3], 0x5140), hi23 = __byte_perm(r[2], r[3], 0x7362);
    const int col = (s5 << 5) + (h << 4) + (tt << 2);
    uint8_t* dst = sm.Vt + (c4 * 4) * kVtStride + col;
    *reinterpret_cast<uint32_t*>(dst) = __byte_perm(lo01, lo23, 0x5410);
    *reinterpret_cast<uint32_t*>(dst + kVtStride) = __byte_perm(lo01, lo23, 0x7632);
    *reinterpret_cast<uint32_t*>(dst + 2 * kVtStride) = __byte_perm(hi01, hi23, 0x5410);
    *reinterpret_cast<uint32_t*>(dst + 3 * kVtStride) = __byte_perm(hi01, hi23, 0x7632);
  }
  __syncthreads();

  // Score codes of this warp's 16 rows live in shared memory in AV key order: the two codes a thread
  // produces for tile j (row g, keys 8j+2t, 8j+2t+1) land at byte 32s + 16h + 4t + 2(j&1) of the row
  // (s = j/4, h = (j/2)&1), so that every later pass reads back, as one 32-bit word, exactly the four
  // codes this same thread wrote: no cross-thread hazard, and the word is the A fragment of the AV MMA.
  uint8_t* const otile = sm.codes[warp];   // output staging tile of the re-quantize step (codes are consumed by then)
  constexpr int kOutPitch = 80;
  uint8_t* crowA = sm.codes[warp] + g * kVtStride + t * 4;
  uint8_t* crowB = crowA + 8 * kVtStride;
  // key index of byte i of word w (w = 2s + h) for this lane
  auto key_of = [&](int w, int i) { return ((w >> 1) << 5) + ((w & 1) << 4) + ((i >> 1) << 3) + t * 2 + (i & 1); };
  const int nsteps = ntiles >> 2;           // 32-key steps
  const int full_steps = n >> 5;            // steps whose 32 keys are all < n
  const float zp_biased = p.score_zp + 128.f;
  const uint32_t rev_e_saddr = (uint32_t)__cvta_generic_to_shared(sm.rev_e);
  const uint32_t rev_ehi_saddr = (uint32_t)__cvta_generic_to_shared(sm.rev_ehi);
  const bool wide = sm.wide != 0;
  const uint32_t rev_r3_saddr = (uint32_t)__cvta_generic_to_shared(sm.rev_r3);
  const int cinit = kPot ? 0x4B400000 : 0;   // 1.5 * 2^23 as an accumulator bias (see the score loop)
  const float zq = kPot ? (float)((double)zp_biased - 12582912.0 * (double)p.score_mul) : zp_biased;
  const int64_t out_stride = (int64_t)heads * kHd;
  const int izp = (int)p.out_zp;
  const int half_m1 = out_shift > 0 ? (1 << (out_shift - 1)) - 1 : 0;

  for (int r0 = warp * 16; r0 < n; r0 += kAttWarps * 16) {
    const int rowA = r0 + g, rowB = r0 + g + 8;
    // ---- Q fragments straight from global memory (each row is read by exactly one warp) -------------------
    uint32_t qa[2][4];
    {
      const int8_t* qA = base + (int64_t)min(rowA, n - 1) * row_stride + head * kHd + t * 4;
      const int8_t* qB = base + (int64_t)min(rowB, n - 1) * row_stride + head * kHd + t * 4;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qA + ks * 32));
        qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qB + ks * 32));
        qa[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qA + ks * 32 + 16));
        qa[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qB + ks * 32 + 16));
      }
    }

    int rcA = 0, rcB = 0;   // kZp: z * sum_d q[row][d] - 64 z^2, subtracted from every raw score of the row
    if (kZp) {
      int qsA = 0, qsB = 0;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qsA = __dp4a((int)qa[ks][0], 0x01010101, __dp4a((int)qa[ks][2], 0x01010101, qsA));
        qsB = __dp4a((int)qa[ks][1], 0x01010101, __dp4a((int)qa[ks][3], 0x01010101, qsB));
      }
      qsA += __shfl_xor_sync(0xffffffffu, qsA, 1);
      qsA += __shfl_xor_sync(0xffffffffu, qsA, 2);
      qsB += __shfl_xor_sync(0xffffffffu, qsB, 1);
      qsB += __shfl_xor_sync(0xffffffffu, qsB, 2);
      rcA = iz * qsA - kHd * iz * iz;
      rcB = iz * qsB - kHd * iz * iz;
    }

    // ---- S = Q K^T -> biased int8 score codes: clamp(RNE(acc * mul + zp), -128, 127) + 128 ----------------
    // acc * mul is exact for the power-of-two multiplier and adding the integer zp + 128 keeps it exact,
    // so one fma followed by an unsigned saturating RNE conversion is the reference's round-then-clamp.
    // kPot: the accumulators start from 0x4B400000, the bit pattern of 1.5 * 2^23, so that the int32 result
    // re-read as fp32 IS 1.5 * 2^23 + acc (|acc| <= 2^20 keeps it in the binade with ulp 1); the constant is
    // taken out again inside the fma's addend, which stays exactly representable.  No int -> float converts.
    // The row maximum is taken on the fp32 values (the conversion is monotone) and converted once.
    float fmaxA = -INFINITY, fmaxB = -INFINITY;
    const float2 mul2 = make_float2(p.score_mul, p.score_mul), zq2 = make_float2(zq, zq);
    auto score_tile = [&](int j, float (&f)[4]) {
      const uint4 kf = *reinterpret_cast<const uint4*>(sm.Ks + (j * 8 + g) * kKStride + t * 16);
      int c[4];
      mma_s8s8_init(c, qa[0], kf.x, kf.y, cinit);
      mma_s8s8(c, qa[1], kf.z, kf.w);
      if (kZp) {
        const int2 zk = *reinterpret_cast<const int2*>(&sm.zks[j * 8 + t * 2]);
        c[0] -= rcA + zk.x; c[1] -= rcA + zk.y; c[2] -= rcB + zk.x; c[3] -= rcB + zk.y;
      }
      // two scores per fma.rn.f32x2 (rows A and B of the tile each hold an adjacent pair)
      const float2 fa = ffma2(make_float2(kPot ? __int_as_float(c[0]) : (float)c[0], kPot ? __int_as_float(c[1]) : (float)c[1]), mul2, zq2);
      const float2 fb = ffma2(make_float2(kPot ? __int_as_float(c[2]) : (float)c[2], kPot ? __int_as_float(c[3]) : (float)c[3]), mul2, zq2);
      f[0] = fa.x; f[1] = fa.y; f[2] = fb.x; f[3] = fb.y;
    };
#pragma unroll 1
    for (int s = 0; s < full_steps; ++s) {
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {     // a pair of 8-key tiles = one 32-bit word of codes per row
        float f0[4], f1[4];
        score_tile(s * 4 + jp * 2, f0);
        score_tile(s * 4 + jp * 2 + 1, f1);
        fmaxA = fmaxf(fmaxf(fmaxA, f0[0]), f0[1]);
        fmaxB = fmaxf(fmaxf(fmaxB, f0[2]), f0[3]);
        fmaxA = fmaxf(fmaxf(fmaxA, f1[0]), f1[1]);
        fmaxB = fmaxf(fmaxf(fmaxB, f1[2]), f1[3]);
        *reinterpret_cast<uint32_t*>(crowA + (s << 5) + (jp << 4)) = pack2_u8(f0[0], f0[1], pack2_u8(f1[0], f1[1], 0u));
        *reinterpret_cast<uint32_t*>(crowB + (s << 5) + (jp << 4)) = pack2_u8(f0[2], f0[3], pack2_u8(f1[2], f1[3], 0u));
      }
    }
    if (full_steps < nsteps) {   // the ragged last step: padded keys get code 0 (never above a real biased code)
      const int s = full_steps;
#pragma unroll 1
      for (int jj = 0; jj < 4; ++jj) {
        const int j = s * 4 + jj;
        uint32_t pa = 0, pb = 0;
        if (j * 8 < n) {
          float f[4];
          score_tile(j, f);
          const int col = j * 8 + t * 2;
          const bool ok0 = col < n, ok1 = col + 1 < n;
          if (ok0) { fmaxA = fmaxf(fmaxA, f[0]); fmaxB = fmaxf(fmaxB, f[2]); }
          if (ok1) { fmaxA = fmaxf(fmaxA, f[1]); fmaxB = fmaxf(fmaxB, f[3]); }
          pa = pack2_u8(f[0], f[1], 0u) & (ok1 ? 0xffffu : (ok0 ? 0xffu : 0u));
          pb = pack2_u8(f[2], f[3], 0u) & (ok1 ? 0xffffu : (ok0 ? 0xffu : 0u));
        }
        const int pos = (s << 5) + ((jj >> 1) << 4) + ((jj & 1) << 1);
        *reinterpret_cast<uint16_t*>(crowA + pos) = (uint16_t)pa;
        *reinterpret_cast<uint16_t*>(crowB + pos) = (uint16_t)pb;
      }
    }
    int maxA, maxB;   // biased maxima
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(maxA) : "f"(fmaxA));
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(maxB) : "f"(fmaxB));
    maxA = max(maxA, __shfl_xor_sync(0xffffffffu, maxA, 1));
    maxA = max(maxA, __shfl_xor_sync(0xffffffffu, maxA, 2));
    maxB = max(maxB, __shfl_xor_sync(0xffffffffu, maxB, 1));
    maxB = max(maxB, __shfl_xor_sync(0xffffffffu, maxB, 2));
    __syncwarp();

    // ---- exact integer row sums of the integer exp ----------------------------------------------------------
    // the fp32 table values are widened and added in fp64 (exact: < 2^51), on the fp64 pipe next to the integer/fp32
    // work of the other warps
    double sumA = 0.0, sumB = 0.0;
    const uint32_t revA = rev_e_saddr + ((255 - maxA) << 2);
    const uint32_t revB = rev_e_saddr + ((255 - maxB) << 2);
    {
      double pa[4] = {0.0, 0.0, 0.0, 0.0}, pb[4] = {0.0, 0.0, 0.0, 0.0};   // independent chains: the adds are exact
      if (!wide) {   // the table word IS the high half of the double: no conversion instruction
        const uint32_t hiA = rev_ehi_saddr + ((255 - maxA) << 2), hiB = rev_ehi_saddr + ((255 - maxB) << 2);
#pragma unroll 2
        for (int w = 0; w < 2 * full_steps; ++w) {
          const uint32_t wa = *reinterpret_cast<const uint32_t*>(crowA + w * 16);
          const uint32_t wb = *reinterpret_cast<const uint32_t*>(crowB + w * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pa[i] += __hiloint2double((int)lds_u32(hiA + (__byte_perm(wa, 0, 0x4440 + i) << 2)), 0);
            pb[i] += __hiloint2double((int)lds_u32(hiB + (__byte_perm(wb, 0, 0x4440 + i) << 2)), 0);
          }
        }
      } else
#pragma unroll 2
      for (int w = 0; w < 2 * full_steps; ++w) {
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(crowA + w * 16);
        const uint32_t wb = *reinterpret_cast<const uint32_t*>(crowB + w * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pa[i] += (double)lds_f32(revA + (__byte_perm(wa, 0, 0x4440 + i) << 2));
          pb[i] += (double)lds_f32(revB + (__byte_perm(wb, 0, 0x4440 + i) << 2));
        }
      }
      sumA = (pa[0] + pa[1]) + (pa[2] + pa[3]);
      sumB = (pb[0] + pb[1]) + (pb[2] + pb[3]);
    }
    for (int w = 2 * full_steps; w < 2 * nsteps; ++w) {
      const uint32_t wa = *reinterpret_cast<const uint32_t*>(crowA + w * 16);
      const uint32_t wb = *reinterpret_cast<const uint32_t*>(crowB + w * 16);
      for (int i = 0; i < 4; ++i) {
        if (key_of(w, i) < n) {
          sumA += (double)lds_f32(revA + (((wa >> (8 * i)) & 0xff) << 2));
          sumB += (double)lds_f32(revB + (((wb >> (8 * i)) & 0xff) << 2));
        }
      }
    }
    sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
    sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
    sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
    sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
    const float fsumA = __double2float_rn(sumA), fsumB = __double2float_rn(sumB);   // exact integer -> RNE, as u64 -> f32
    const uint32_t lutRA = rev_r3_saddr + ((255 - maxA) << 2);
    const uint32_t lutRB = rev_r3_saddr + ((255 - maxB) << 2);

    // ---- log2 codes -> two u8 probability planes -> P V ------------------------------------------------------
    int8_t* dsc = kDump ? p.dump_scores + ((int64_t)bh * n) * n : nullptr;
    uint8_t* dsm = kDump ? p.dump_softmax + ((int64_t)bh * n) * n : nullptr;
    // One accumulator set for both probability planes: the first sweep multiplies the high-byte plane and
    // parks the low-byte plane in shared memory over the (now consumed) score codes; the accumulators are then
    // scaled by 2^8 and a second sweep adds the low-byte plane.  32 instead of 64 accumulator registers per
    // thread lets three CTAs share an SM.
    int acc[8][4];
#pragma unroll
    for (int jn = 0; jn < 8; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[jn][e] = 0;
    int psA = 0, psB = 0;   // kZp: sum of the 16-bit probabilities of the row (this lane's keys)

#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
      uint32_t pa_hi[4], pa_lo[4];  // a0..a3 of the two planes
      const bool full = s < full_steps;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {        // hh = 0: keys of a0/a1; hh = 1: keys of a2/a3
        const int w = 2 * s + hh;
        if (!full && w * 16 >= n) {   // sixteen padded keys: probability 0 in both planes
          pa_hi[2 * hh] = pa_hi[2 * hh + 1] = pa_lo[2 * hh] = pa_lo[2 * hh + 1] = 0u;
          continue;
        }
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(crowA + w * 16);
        const uint32_t wb = *reinterpret_cast<const uint32_t*>(crowB + w * 16);
        uint32_t va[4], vb[4], ba[4], bb[4];
        const int fa = prob16x4(wa, lutRA, fsumA, va, ba);
        const int fb = prob16x4(wb, lutRB, fsumB, vb, bb);
        if (min(fa, fb) <= 0) {   // rare: a value next to a step of the code function, or a dominant key
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (code_guarded(ba[i])) va[i] = exact_prob16(fsumA, sm.lut_f[maxA - (int)((wa >> (8 * i)) & 0xff)], p.softmax_levels);
            if (code_guarded(bb[i])) vb[i] = exact_prob16(fsumB, sm.lut_f[maxB - (int)((wb >> (8 * i)) & 0xff)], p.softmax_levels);
          }
        }
        if (!full) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = key_of(w, i) < n;
            va[i] = ok ? va[i] : 0u;
            vb[i] = ok ? vb[i] : 0u;
          }
        }
        if (kDump) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = key_of(w, i);
            if (col < n) {
              const int ka = va[i] ? __clz(va[i]) - 16 : p.softmax_levels;
              const int kb = vb[i] ? __clz(vb[i]) - 16 : p.softmax_levels;
              if (rowA < n) { dsc[(int64_t)rowA * n + col] = (int8_t)((int)((wa >> (8 * i)) & 0xff) - 128); dsm[(int64_t)rowA * n + col] = (uint8_t)ka; }
              if (rowB < n) { dsc[(int64_t)rowB * n + col] = (int8_t)((int)((wb >> (8 * i)) & 0xff) - 128); dsm[(int64_t)rowB * n + col] = (uint8_t)kb; }
            }
          }
        }
        if (kZp) {
          psA += (int)(va[0] + va[1] + va[2] + va[3]);
          psB += (int)(vb[0] + vb[1] + vb[2] + vb[3]);
        }
        // 16-bit probabilities -> low-byte plane and high-byte plane, 4 keys per register
        const uint32_t a01 = va[0] | (va[1] << 16), a23 = va[2] | (va[3] << 16);
        const uint32_t b01 = vb[0] | (vb[1] << 16), b23 = vb[2] | (vb[3] << 16);
        pa_lo[2 * hh] = __byte_perm(a01, a23, 0x6420);
        pa_hi[2 * hh] = __byte_perm(a01, a23, 0x7531);
        pa_lo[2 * hh + 1] = __byte_perm(b01, b23, 0x6420);
        pa_hi[2 * hh + 1] = __byte_perm(b01, b23, 0x7531);
      }
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const uint8_t* vp = sm.Vt + (jn * 8 + g) * kVtStride + s * 32 + t * 4;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vp), b1 = *reinterpret_cast<const uint32_t*>(vp + 16);
        mma_u8s8(acc[jn], pa_hi[0], pa_hi[1], pa_hi[2], pa_hi[3], b0, b1);
      }
      // this thread wrote these four code words and is their only reader: overwrite them with the low plane
      *reinterpret_cast<uint32_t*>(crowA + (2 * s) * 16) = pa_lo[0];
      *reinterpret_cast<uint32_t*>(crowB + (2 * s) * 16) = pa_lo[1];
      *reinterpret_cast<uint32_t*>(crowA + (2 * s + 1) * 16) = pa_lo[2];
      *reinterpret_cast<uint32_t*>(crowB + (2 * s + 1) * 16) = pa_lo[3];
    }
#pragma unroll
    for (int jn = 0; jn < 8; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[jn][e] <<= 8;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
      const uint32_t l0 = *reinterpret_cast<const uint32_t*>(crowA + (2 * s) * 16);
      const uint32_t l1 = *reinterpret_cast<const uint32_t*>(crowB + (2 * s) * 16);
      const uint32_t l2 = *reinterpret_cast<const uint32_t*>(crowA + (2 * s + 1) * 16);
      const uint32_t l3 = *reinterpret_cast<const uint32_t*>(crowB + (2 * s + 1) * 16);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const uint8_t* vp = sm.Vt + (jn * 8 + g) * kVtStride + s * 32 + t * 4;
        mma_u8s8(acc[jn], l0, l1, l2, l3, *reinterpret_cast<const uint32_t*>(vp), *reinterpret_cast<const uint32_t*>(vp + 16));
      }
    }

    // ---- re-quantize and store ----------------------------------------------------------------------------
    if (kZp) {
      psA += __shfl_xor_sync(0xffffffffu, psA, 1);
      psA += __shfl_xor_sync(0xffffffffu, psA, 2);
      psB += __shfl_xor_sync(0xffffffffu, psB, 1);
      psB += __shfl_xor_sync(0xffffffffu, psB, 2);
      psA *= iz;
      psB *= iz;
    }
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      int q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int a = acc[jn][e] - (kZp ? (e < 2 ? psA : psB) : 0);
        if (out_shift > 0) {
          // RNE(acc / 2^sh) in integers: add half minus one plus the parity of the truncated result
          q[e] = ((a + half_m1 + ((a >> out_shift) & 1)) >> out_shift) + izp;
        } else {
          const double v = rint((double)a * p.out_mul) + (double)p.out_zp;
          q[e] = (int)fmin(fmax(v, -128.0), 127.0);
        }
      }
      // into this warp's (now free) code buffer as a 16 x 64-byte tile with an 80-byte pitch (conflict-free)
      *reinterpret_cast<uint16_t*>(otile + g * kOutPitch + jn * 8 + t * 2) = (uint16_t)pack2_s8(q[0], q[1]);
      *reinterpret_cast<uint16_t*>(otile + (g + 8) * kOutPitch + jn * 8 + t * 2) = (uint16_t)pack2_s8(q[2], q[3]);
    }
    __syncwarp();
    // ... and out as whole 64-byte rows: four lanes x 16 bytes per row, eight rows per store instruction (the
    // 2-byte stores of the fragment layout wrote 2.9 M partial sectors per launch for a 19 MB output, ncu r1q)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rr = (lane >> 2) + 8 * i;
      if (r0 + rr < n)
        *reinterpret_cast<uint4*>(out + ((int64_t)img * n + r0 + rr) * out_stride + head * kHd + (lane & 3) * 16) =
            *reinterpret_cast<const uint4*>(otile + rr * kOutPitch + (lane & 3) * 16);
    }
    __syncwarp();   // the next row tile overwrites this warp's code buffer
  }
}

int attention_tc_configure();

template <bool kDump, bool kPot, bool kZp>
static int attention_configure_one() {
  P2V_CHECK_CUDA(cudaFuncSetAttribute(attention_int_kernel<kDump, kPot, kZp>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(AttSmem)));
  // three CTAs per SM need the full shared-memory carveout
  P2V_CHECK_CUDA(cudaFuncSetAttribute(attention_int_kernel<kDump, kPot, kZp>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return P2V_OK;
}
static int attention_configure() {
  static unsigned long long done = 0;   // one bit per device
  int dev = 0;
  if (needs_configure(done, &dev)) {
    int rc;
    if ((rc = attention_configure_one<false, false, false>()) || (rc = attention_configure_one<false, true, false>()) ||
        (rc = attention_configure_one<true, false, false>()) || (rc = attention_configure_one<true, true, false>()) ||
        (rc = attention_configure_one<false, false, true>()) || (rc = attention_configure_one<true, false, true>()))
      return rc;
    if ((rc = attention_tc_configure())) return rc;
    mark_configured(done, dev);
  }
  return P2V_OK;
}
int attention_tc_configure();
bool attention_tc_applicable(const int8_t* qkv, const int8_t* out, int b, int n, int heads, const p2v_attention* p);
int attention_tc_launch(const int8_t* qkv, int8_t* out, int b, int n, int heads, const p2v_attention* p, cudaStream_t st);

int attention_configure_once() { return attention_configure(); }

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_attention_int(const int8_t* qkv, int8_t* out, int b, int n, int heads, const p2v_attention* p,
                                 void* stream) {
  P2V_REQUIRE(qkv && out && p && p->exp_lut, "p2v_attention_int: null pointer");
  P2V_REQUIRE(b > 0 && heads > 0 && n > 0, "p2v_attention_int: bad shape b=%d n=%d heads=%d", b, n, heads);
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "p2v_attention_int: qkv and out must be 16-byte aligned");
  P2V_REQUIRE(n <= kMaxKeys, "p2v_attention_int: n=%d tokens exceeds the %d-key tile of this kernel", n, kMaxKeys);
  P2V_REQUIRE((p->dump_scores == nullptr) == (p->dump_softmax == nullptr),
              "p2v_attention_int: dump_scores and dump_softmax must be given together");
  // power-of-two grids (every minmax-calibrated model): the tcgen05 / TMEM / TMA kernel of p2v_attention_tc.cu
  if (!p->force_legacy && attention_tc_applicable(qkv, out, b, n, heads, p)) {
    int rc0 = attention_configure_once();
    if (rc0) return rc0;
    return attention_tc_launch(qkv, out, b, n, heads, p, (cudaStream_t)stream);
  }
  dim3 grid(b * heads);
  // power-of-two output multiplier 2^-sh (every minmax-calibrated model): integer RNE shift in the kernel
  int out_shift = 0, ex = 0;
  if (p->out_mul > 0 && frexp(p->out_mul, &ex) == 0.5 && ex <= 0 && ex >= -29 && p->out_zp == (float)(int)p->out_zp)
    out_shift = 1 - ex;
  int rc = attention_configure_once();
  if (rc) return rc;
  const size_t smem = sizeof(AttSmem);
  // power-of-two score multiplier 2^-sh, sh >= 0, and an integer zero point: the biased-accumulator conversion
  int ex2 = 0;
  const bool zp = p->in_zp != 0.f;
  P2V_REQUIRE(p->in_zp == (float)(int)p->in_zp && fabsf(p->in_zp) <= 128.f, "p2v_attention_int: in_zp must be an int8 code");
  const bool pot = !zp && p->score_mul > 0.f && frexpf(p->score_mul, &ex2) == 0.5f && ex2 <= 1 && ex2 >= -30 &&
                   p->score_zp == (float)(int)p->score_zp;
  const bool dump = p->dump_scores != nullptr;
  const cudaStream_t st = (cudaStream_t)stream;
#define P2V_ATT_LAUNCH(D, P, Z) attention_int_kernel<D, P, Z><<<grid, kAttWarps * 32, smem, st>>>(qkv, out, n, heads, *p, out_shift)
  if (zp && dump) P2V_ATT_LAUNCH(true, false, true);
  else if (zp) P2V_ATT_LAUNCH(false, false, true);
  else if (dump && pot) P2V_ATT_LAUNCH(true, true, false);
  else if (dump) P2V_ATT_LAUNCH(true, false, false);
  else if (pot) P2V_ATT_LAUNCH(false, true, false);
  else P2V_ATT_LAUNCH(false, false, false);
#undef P2V_ATT_LAUNCH
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
