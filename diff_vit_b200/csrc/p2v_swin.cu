// p2v_swin.cu - the kernels the quantized Swin forward (BASELINE config 5) needs beside the GEMM / LayerNorm family:
//
//   p2v_window_attention_int   windowed / shifted-window integer attention      (models/swin_quant.py:177-221)
//   p2v_quant_patchify (p < 16) input QAct + im2col for the 4 x 4 patch conv     (models/layers_quant.py:211-224)
//   p2v_gather_row_segments    the 2 x 2 neighbourhood concat of PatchMerging   (models/swin_quant.py:445-457)
//   p2v_avgpool_requant        token average + qact3                            (models/swin_quant.py:810-813)
//
// Window attention, per (window, head): one thread owns one of the n = ws^2 <= 64 query rows (five items per CTA).
//   q' = fl32(q * head_dim^-1/2)   - the reference scales the DEQUANTIZED q in fp32 before the product, so q' is not on
//                                    an integer grid.  fl32(code * s * c) = s * fl32(code * c) for a power-of-two s, and
//                                    every fl32(code * c), |code| <= 128, is a multiple of 2^-qshift below 2^31 in that
//                                    unit: the product is accumulated EXACTLY as sum_c M[q_c] * k_c in 64-bit integers
//                                    (M = fl32(code * c) * 2^qshift) and rounded to fp32 once - the reference's expression
//                                    without the summation-order noise of an fp32 BLAS (oracle accum = 'fp64').
//   a1 = qact_attn1(S)             RNE(S / s_a1), power-of-two grid
//   a2 = qact2(a1 s_a1 + bias)     relative-position bias: the qact_table-quantized table, gathered per (row, key)
//   x  = a2 - 100 / s_2 [masked]   the -100 of the shifted-window mask, added after qact2 (swin_quant.py:205-209)
//   log-int-softmax over the keys  I-BERT integer exp as a table over d = max - x, exact row sum, 4-bit log2 code
//   O  = sum_j 2^(15-k_j) v_j      int32, then qact3
// Rows are addressed through the layer's window permutation (roll + window_partition composed on the host), for loads
// and for stores alike, so q/k/v and the output stay in token order and no partition / reverse copies exist.
#include <limits.h>

#include "p2v_common.cuh"
#include "p2v_math.cuh"
#include "p2v_requant.cuh"

namespace p2v {

constexpr int kWaMaxN = 64;     // tokens per window
constexpr int kWaHeadDim = 32;  // every Swin variant of the reference: C / heads = 32
constexpr int kWaVtStride = 36; // words per 4-key group of V^T: 16-byte aligned rows, <= 2-way bank conflicts on the fill

__device__ __forceinline__ uint32_t pack4_sat_s8(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}
__device__ __forceinline__ int sx8(uint32_t w, int j) { return (int)(int8_t)((w >> (8 * j)) & 0xffu); }
// c + sum of (unsigned byte of a) * (signed byte of b)
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// The reference's score of one (row, key) pair, exactly: sum_c fl32(q_c * qscale) * k_c in 64-bit fixed point, one
// rounding to fp32.  Only rows' keys whose fast evaluation lands next to a rounding tie of qact_attn1 come here.
__device__ __noinline__ float wa_exact_score(uint4 q0, uint4 q1, const uint32_t* krow, float qscale, int qshift,
                                             double acc_scale) {
  const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  const float unit = __uint_as_float((uint32_t)(127 + qshift) << 23);   // 2^qshift
  long long acc = 0;
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    const uint32_t kw = krow[x];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // fl32(code * qscale) * 2^qshift: an integer below 2^31 (see the file header), so the conversion is exact
      const int m = __float2int_rn(fmul(fmul((float)sx8(qw[x], e), qscale), unit));
      acc += (long long)m * sx8(kw, e);
    }
  }
  return __double2float_rn(__dmul_rn(__ll2double_rn(acc), acc_scale));
}

__device__ __noinline__ int wa_exact_code(float fsum, float e, int levels) { return softmax_log_code(fsum, e, levels); }

// Fast paths with exact fall-backs (both return the reference's codes; tests/test_gpu_swin.py compares every code of
// every layer with the oracle):
//   * score: S' = (sum_c q_c k_c) * (qscale * s^2) with the integer product from eight dp4a.  fl32 rounds within
//     2^-24 relative, so S' differs from the exact S by at most 2^-24 qscale sum_c |q_c| * 128 * s^2 - a per-row bound
//     from the prologue -, and RNE(S' / s_a1) is the exact code unless S' / s_a1 lies within that bound (plus the fp32
//     rounding of S' itself) of a half-integer; those pairs (~1e-4) are redone by wa_exact_score.
//   * both re-quantisations round with the 1.5 * 2^23 constant on the FMA pipe (clamp and RNE commute for integer
//     bounds); the int8 code is the low byte of the biased sum: no FRND / F2I (conversion unit, quarter rate) in the
//     loop.
//   * log2 code: k = exponent(fma(sum, 1 / (3e), 1/6)) + 2 evaluated for a low and a high bracket of 1 / (3e)
//     (+-2^-20, table `r3`); where the two exponents differ, or k <= 1 (the irregular first step of the code
//     function), the pair takes softmax_log_code's IEEE division (see p2v_attention_tc.cu for the argument).
// Several (window, head) items share a CTA of 256 threads: thread tid holds row tid % n of the CTA's item tid / n, so 5 items
// of 49 rows fill 245 of the 256 lanes (one item per 64-thread CTA left 15 of 64 idle).  Each item has its own K / V^T
// tiles; the score / distance columns are per thread.
constexpr int kWaThreads = 256;
constexpr int kWaMaxItems = 5;
struct WaSmem {
  alignas(16) uint32_t kp[kWaMaxItems][kWaMaxN][8];                 // K rows, packed int8
  alignas(16) uint32_t vt[kWaMaxItems][kWaMaxN / 4][kWaVtStride];   // V^T: word [g][c] = v[4g .. 4g + 3][c]
  int8_t xs[kWaMaxN][kWaThreads];                                   // [key][thread] qact2 codes
  uint16_t ds[kWaMaxN][kWaThreads];                                 // [key][thread] distance to the row maximum (clamped)
  uint8_t rid[kWaMaxItems][kWaMaxN];
};

template <bool kDump>
__global__ void __launch_bounds__(kWaThreads, 3)
window_attention_kernel(const int8_t* __restrict__ qkv, int8_t* __restrict__ out, const p2v_window_attention a,
                        int items_total, int items_per_cta) {
  extern __shared__ __align__(16) uint8_t wa_smem_raw[];
  WaSmem& sm = *reinterpret_cast<WaSmem*>(wa_smem_raw);
  constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23
  const int n = a.n, C = a.channels;
  const int tid = threadIdx.x;
  const int it = tid / n, i = tid - it * n;           // item within the CTA, row within the item
  const int item = blockIdx.x * items_per_cta + it;
  const bool valid = it < items_per_cta && item < items_total;
  const int head = valid ? item % a.heads : 0;
  const int wg = valid ? item / a.heads : 0;      // window over the whole batch
  const int img = wg / a.windows, w = wg % a.windows;
  uint32_t (*kp)[8] = sm.kp[valid ? it : 0];
  uint32_t (*vt)[kWaVtStride] = sm.vt[valid ? it : 0];
  const uint8_t* rid = sm.rid[valid ? it : 0];
  // keys n .. 63 of every item: zero K rows and V columns, so the loops below can run over whole groups of four keys
  for (int z = tid; z < kWaMaxItems * (kWaMaxN - n); z += kWaThreads) {
    const int zi = z / (kWaMaxN - n), zr = n + z % (kWaMaxN - n);
    *reinterpret_cast<uint4*>(&sm.kp[zi][zr][0]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&sm.kp[zi][zr][4]) = make_uint4(0, 0, 0, 0);
    for (int c = 0; c < kWaHeadDim; ++c) reinterpret_cast<uint8_t*>(&sm.vt[zi][zr >> 2][c])[zr & 3] = 0;
  }
  int64_t src_row = 0;
  uint32_t qw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float guard = 0.f;
  if (valid) {
    src_row = (int64_t)img * a.tokens + a.perm[w * n + i];
    const int8_t* base = qkv + src_row * (3 * (int64_t)C) + head * kWaHeadDim;
    uint32_t qabs = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 q4 = *reinterpret_cast<const uint4*>(base + 16 * h);
      const uint4 k4 = *reinterpret_cast<const uint4*>(base + C + 16 * h);
      const uint4 v4 = *reinterpret_cast<const uint4*>(base + 2 * C + 16 * h);
      const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
      qw[4 * h] = q4.x; qw[4 * h + 1] = q4.y; qw[4 * h + 2] = q4.z; qw[4 * h + 3] = q4.w;
      *reinterpret_cast<uint4*>(&kp[i][4 * h]) = k4;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        qabs = __dp4a(__vabsss4(qw[4 * h + x]), 0x01010101u, qabs);   // sum |q_c| (|-128| counted as 127: + 32 below)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          reinterpret_cast<uint8_t*>(&vt[i >> 2][16 * h + 4 * x + e])[i & 3] = (uint8_t)(vw[x] >> (8 * e));
      }
    }
    // |S' - S| / s_a1 <= 2^-24 qscale sum|q_c| * 128 s^2 / s_a1 = sum|q_c| * err_mul; + 2^-14 for the roundings of S'
    // itself (|t| < 2^9 wherever the code is not saturated)
    guard = fadd(fmul((float)(qabs + 32u), a.err_mul), 6.103515625e-05f);
    sm.rid[it][i] = a.region != nullptr ? a.region[w * n + i] : (uint8_t)0;
  }
  __syncthreads();
  if (!valid) return;

  const int64_t dump_row = ((int64_t)item * n + i) * n;
  const float* bias_p = a.bias + (int64_t)head * n * n + i;
  const int my_rid = rid[i];
  const int mask_int = a.mask_int;
  const float a1_rscale = a.a1_rscale, a1_scale = a.a1_scale, a2_rscale = a.a2_rscale;
  const float qkrs = fmul(a.qk_scale, a1_rscale);   // exact: a1_rscale is a power of two
  // ---- scores: integer product, the two re-quantisations, row maximum of the masked codes ----
  int mx = INT_MIN;
#pragma unroll 2
  for (int j = 0; j < n; ++j) {
    const uint4 k0 = *reinterpret_cast<const uint4*>(&kp[j][0]);
    const uint4 k1 = *reinterpret_cast<const uint4*>(&kp[j][4]);
    int acc = __dp4a((int)qw[0], (int)k0.x, 0);
    acc = __dp4a((int)qw[1], (int)k0.y, acc);
    acc = __dp4a((int)qw[2], (int)k0.z, acc);
    acc = __dp4a((int)qw[3], (int)k0.w, acc);
    acc = __dp4a((int)qw[4], (int)k1.x, acc);
    acc = __dp4a((int)qw[5], (int)k1.y, acc);
    acc = __dp4a((int)qw[6], (int)k1.z, acc);
    acc = __dp4a((int)qw[7], (int)k1.w, acc);
    const float bias = __ldg(bias_p + j * n);
    // qact_attn1: 1.5 * 2^23 + RNE(acc * qkrs) from ONE fused multiply-add (|acc * qkrs| < 2^22, checked by the host),
    // the distance to the rounded value from a second one; the clamp to int8 follows the rounding (they commute)
    const float accf = (float)acc;
    float c1 = fsub(ffma(accf, qkrs, kMagic), kMagic);
    if (fabsf(fsub(fabsf(ffma(accf, qkrs, -c1)), 0.5f)) < guard) {   // next to a rounding tie: the exact product decides
      const float te = fmul(wa_exact_score(make_uint4(qw[0], qw[1], qw[2], qw[3]), make_uint4(qw[4], qw[5], qw[6], qw[7]),
                                           &kp[j][0], a.qscale, a.qshift, a.acc_scale), a1_rscale);
      c1 = fsub(fadd(fminf(fmaxf(te, -256.f), 256.f), kMagic), kMagic);
    }
    c1 = fminf(fmaxf(c1, -128.f), 127.f);
    // + bias (c1 * s_a1 is exact, so the fused form rounds like the reference's add), qact2
    const float tb = fminf(fmaxf(fmul(ffma(c1, a1_scale, bias), a2_rscale), -128.f), 127.f);
    const int x = (int)(int8_t)(__float_as_uint(fadd(tb, kMagic)) & 0xffu);   // low byte of 1.5 * 2^23 + RNE(tb)
    sm.xs[j][tid] = (int8_t)x;
    if (kDump) {
      a.dump_a1[dump_row + j] = (int8_t)c1;
      a.dump_a2[dump_row + j] = (int8_t)x;
    }
    mx = max(mx, x - (rid[j] != my_rid ? mask_int : 0));
  }
  // ---- row sum of the integer exp; the distances are kept for the last pass ----
  const int dmax = a.lut_n - 1;
  double sum = 0.0;
#pragma unroll 4
  for (int j = 0; j < n; ++j) {
    const int xm = (int)sm.xs[j][tid] - (rid[j] != my_rid ? mask_int : 0);
    const int d = min(mx - xm, dmax);
    sm.ds[j][tid] = (uint16_t)d;
    sum += __ldg(a.exp_lut64 + d);
  }
  for (int j = n; j < ((n + 3) & ~3); ++j) sm.ds[j][tid] = (uint16_t)dmax;   // padding keys: probability 0 (and V rows of 0)
  const float fsum = __double2float_rn(sum);
  // ---- log2 codes -> probabilities 2^(15-k) as two byte planes, four keys to a word; P V by dp4a against V^T ----
  int oh[kWaHeadDim], ol[kWaHeadDim];
#pragma unroll
  for (int c = 0; c < kWaHeadDim; ++c) oh[c] = ol[c] = 0;
  const int levels = a.softmax_levels;
  const float2* r3p = reinterpret_cast<const float2*>(a.r3);
  for (int g = 0; 4 * g < n; ++g) {
    uint32_t ph = 0, pl = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * g + e;
      const int d = sm.ds[j][tid];
      const float2 r3 = __ldg(r3p + d);
      const uint32_t ulo = __float_as_uint(ffma(fsum, r3.x, 0.16666667f)), uhi = __float_as_uint(ffma(fsum, r3.y, 0.16666667f));
      int k = (int)(ulo >> 23) - 125;                      // u > 0: the sign bit is clear
      if (((ulo ^ uhi) >> 23) != 0u || ulo < 0x3f800000u)  // brackets disagree, or k <= 1
        k = wa_exact_code(fsum, a.exp_lut[d], levels);
      if (kDump && j < n) a.dump_softmax[dump_row + j] = (uint8_t)min(k, levels);
      const uint32_t p = k >= levels ? 0u : (0x8000u >> k);
      ph = __byte_perm(ph, p, e == 0 ? 0x3215 : (e == 1 ? 0x3250 : (e == 2 ? 0x3510 : 0x5210)));   // byte 1 of p -> byte e
      pl = __byte_perm(pl, p, e == 0 ? 0x3214 : (e == 1 ? 0x3240 : (e == 2 ? 0x3410 : 0x4210)));   // byte 0 of p -> byte e
    }
#pragma unroll
    for (int c = 0; c < kWaHeadDim; c += 4) {
      const uint4 vv = *reinterpret_cast<const uint4*>(&vt[g][c]);
      oh[c] = dp4a_us(ph, vv.x, oh[c]);         ol[c] = dp4a_us(pl, vv.x, ol[c]);
      oh[c + 1] = dp4a_us(ph, vv.y, oh[c + 1]); ol[c + 1] = dp4a_us(pl, vv.y, ol[c + 1]);
      oh[c + 2] = dp4a_us(ph, vv.z, oh[c + 2]); ol[c + 2] = dp4a_us(pl, vv.z, ol[c + 2]);
      oh[c + 3] = dp4a_us(ph, vv.w, oh[c + 3]); ol[c + 3] = dp4a_us(pl, vv.w, ol[c + 3]);
    }
  }
  // ---- qact3 and the store, back through the permutation ----
  uint32_t w8[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = oh[4 * x + e] * 256 + ol[4 * x + e];
      const float val = fmul(__int2float_rn(o), a.out_unit);    // the fp32 value of (attn @ v): one rounding
      q[e] = (int)rintf(fmul(val, a.out_rscale));
    }
    w8[x] = pack4_sat_s8(q[0], q[1], q[2], q[3]);
  }
  int8_t* dst = out + src_row * C + head * kWaHeadDim;
  *reinterpret_cast<uint4*>(dst) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
  *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
}

// ---- input quantizer + im2col for small patches (p % 4 == 0) -----------------------------------------------------
// One thread converts four consecutive pixels of one patch row: codes[patch][(ch * p + kh) * p + kw .. kw + 3].
__global__ void __launch_bounds__(256)
quant_patchify_small_kernel(const float* __restrict__ x, int8_t* __restrict__ codes, int c, int h, int w, int p,
                            float scale, float zp, int64_t total_words) {
  const int wp = p / 4, gw = w / p, gh = h / p;
  const float rs = __frcp_rn(scale);
  const float s4[4] = {scale, scale, scale, scale}, rs4[4] = {rs, rs, rs, rs};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_words; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int kw4 = (int)(t % wp); t /= wp;
    const int kh = (int)(t % p); t /= p;
    const int ch = (int)(t % c); t /= c;
    const int px = (int)(t % gw); t /= gw;
    const int py = (int)(t % gh);
    const int64_t img = t / gh;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((img * c + ch) * h + (py * p + kh)) * (int64_t)w + px * p + kw4 * 4));
    // four pixels per guarded rounding (multiply by fl(1 / s), exact IEEE division only next to a rounding tie)
    const float y4[4] = {v.x, v.y, v.z, v.w};
    float r4[4];
    div_round4(y4, s4, rs4, zp, r4);
    reinterpret_cast<uint32_t*>(codes)[i] = pack_sat4(r4[0], r4[1], r4[2], r4[3]);
  }
}

// ---- out[img][r][s * seg .. (s + 1) * seg) = in[img][idx[r * segs + s]][0 .. seg), 16 bytes per thread ------------
__global__ void __launch_bounds__(256)
gather_row_segments_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, const int32_t* __restrict__ idx,
                           int rows_in, int rows_out, int segs, int chunks, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int q = (int)(t % chunks); t /= chunks;
    const int s = (int)(t % segs); t /= segs;
    const int r = (int)(t % rows_out);
    const int64_t img = t / rows_out;
    const int64_t src = img * rows_in + __ldg(idx + r * segs + s);
    reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(in + (src * chunks + q) * 16);
  }
}

// ---- adaptive_avg_pool1d over the tokens + the QAct that follows --------------------------------------------------
// The fp32 sum of code * s_in over <= 2^16 tokens is exact for a power-of-two s_in (an integer below 2^24 times s_in),
// so summing the codes in int32 and converting once equals ATen's fp32 accumulation in any order.
__global__ void __launch_bounds__(256)
avgpool_requant_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int tokens, int channels, int64_t total,
                       float in_scale, float out_scale, float out_zp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const int64_t img = i / channels;
    const int8_t* src = in + img * tokens * channels + c;
    int s = 0;
    for (int t = 0; t < tokens; ++t) s += (int)src[(int64_t)t * channels];
    const float mean = fdiv(fmul((float)s, in_scale), (float)tokens);
    out[i] = (int8_t)quant_div(mean, out_scale, out_zp, -128, 127);
  }
}

static int grid_for_work(int64_t work, int block) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int quant_patchify_small_launch(const float* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                float zero_point, cudaStream_t st) {
  P2V_REQUIRE(p % 4 == 0 && w % 4 == 0, "p2v_quant_patchify: patch size and width must be multiples of 4");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0,
              "p2v_quant_patchify: x must be 16-byte aligned");
  const int64_t total = (int64_t)b * c * h * (w / 4);
  quant_patchify_small_kernel<<<grid_for_work(total, 256), 256, 0, st>>>(x, codes, c, h, w, p, scale, zero_point, total);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

// The same from 8-bit pixels (see quant_patchify_u8_kernel in p2v_rowops.cu): the loader's (p / 255 - mean[c]) / std[c]
// and the input quantizer are evaluated once per (channel, pixel value) into a shared-memory table of codes per block,
// and the image becomes a byte gather, four pixels per thread.
struct U8NormSmall {
  float mean[4], stdv[4];
};
__global__ void __launch_bounds__(256)
quant_patchify_small_u8_kernel(const uint8_t* __restrict__ x, int8_t* __restrict__ codes, int c, int h, int w, int p,
                               float scale, float zp, U8NormSmall nm, int64_t total_words) {
  __shared__ uint8_t lut[4][256];
  for (int i = threadIdx.x; i < c * 256; i += blockDim.x) {
    const int ch = i >> 8, v = i & 255;
    const float xn = fdiv(fsub(fdiv((float)v, 255.0f), nm.mean[ch]), nm.stdv[ch]);
    lut[ch][v] = (uint8_t)(int8_t)quant_div(xn, scale, zp, -128, 127);
  }
  __syncthreads();
  const int wp = p / 4, gw = w / p, gh = h / p;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_words; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int kw4 = (int)(t % wp); t /= wp;
    const int kh = (int)(t % p); t /= p;
    const int ch = (int)(t % c); t /= c;
    const int px = (int)(t % gw); t /= gw;
    const int py = (int)(t % gh);
    const int64_t img = t / gh;
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(x + ((img * c + ch) * h + (py * p + kh)) * (int64_t)w + px * p + kw4 * 4));
    reinterpret_cast<uint32_t*>(codes)[i] = (uint32_t)lut[ch][v & 0xff] | ((uint32_t)lut[ch][(v >> 8) & 0xff] << 8) |
                                            ((uint32_t)lut[ch][(v >> 16) & 0xff] << 16) | ((uint32_t)lut[ch][v >> 24] << 24);
  }
}

int quant_patchify_small_u8_launch(const uint8_t* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                                   float zero_point, const float* mean, const float* stdv, cudaStream_t st) {
  P2V_REQUIRE(p % 4 == 0 && w % 4 == 0, "p2v_quant_patchify_u8: patch size and width must be multiples of 4");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0,
              "p2v_quant_patchify_u8: x must be 4-byte aligned");
  U8NormSmall nm = {};
  for (int i = 0; i < c; ++i) {
    nm.mean[i] = mean[i];
    nm.stdv[i] = stdv[i];
  }
  const int64_t total = (int64_t)b * c * h * (w / 4);
  quant_patchify_small_u8_kernel<<<grid_for_work(total, 256), 256, 0, st>>>(x, codes, c, h, w, p, scale, zero_point, nm, total);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_window_attention_int(const int8_t* qkv, int8_t* out, int images, const p2v_window_attention* p,
                                        void* stream) {
  P2V_REQUIRE(qkv && out && p && p->perm && p->bias && p->exp_lut && p->exp_lut64 && p->r3,
              "p2v_window_attention_int: null pointer");
  const bool dump = p->dump_a1 != nullptr || p->dump_a2 != nullptr || p->dump_softmax != nullptr;
  P2V_REQUIRE(!dump || (p->dump_a1 && p->dump_a2 && p->dump_softmax), "p2v_window_attention_int: the three dumps go together");
  P2V_REQUIRE(images > 0 && p->n > 0 && p->n <= kWaMaxN && p->heads > 0 && p->windows > 0,
              "p2v_window_attention_int: bad shape images=%d n=%d heads=%d windows=%d", images, p->n, p->heads, p->windows);
  P2V_REQUIRE(p->channels == p->heads * kWaHeadDim, "p2v_window_attention_int: channels=%d is not heads * 32", p->channels);
  P2V_REQUIRE(p->tokens == p->windows * p->n, "p2v_window_attention_int: tokens=%d is not windows * n", p->tokens);
  P2V_REQUIRE(p->qk_scale > 0.f && p->a1_rscale > 0.f && (double)p->qk_scale * p->a1_rscale <= 4.0,
              "p2v_window_attention_int: qk_scale / s_a1 = %g: scores beyond 2^21 grid steps", (double)p->qk_scale * p->a1_rscale);
  P2V_REQUIRE(p->lut_n <= 65536, "p2v_window_attention_int: table of %d entries (at most 65536)", p->lut_n);
  P2V_REQUIRE(p->lut_n >= 1 && p->qshift >= 0 && p->qshift <= 60 && p->softmax_levels >= 1 && p->softmax_levels <= 16,
              "p2v_window_attention_int: bad table / shift / levels");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "p2v_window_attention_int: qkv and out must be 16-byte aligned");
  const int64_t items = (int64_t)images * p->windows * p->heads;
  P2V_REQUIRE(items < (1ll << 31), "p2v_window_attention_int: too many (window, head) items");
  static unsigned long long configured = 0;
  int dev = 0;
  if (needs_configure(configured, &dev)) {
    P2V_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaSmem)));
    P2V_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WaSmem)));
    mark_configured(configured, dev);
  }
  const int per_cta = kWaThreads / p->n < kWaMaxItems ? kWaThreads / p->n : kWaMaxItems;
  const unsigned grid = (unsigned)((items + per_cta - 1) / per_cta);
  if (dump) window_attention_kernel<true><<<grid, kWaThreads, sizeof(WaSmem), (cudaStream_t)stream>>>(qkv, out, *p, (int)items, per_cta);
  else window_attention_kernel<false><<<grid, kWaThreads, sizeof(WaSmem), (cudaStream_t)stream>>>(qkv, out, *p, (int)items, per_cta);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_gather_row_segments(const int8_t* in, int8_t* out, const int32_t* idx, int images, int rows_in,
                                       int rows_out, int segs, int seg_bytes, void* stream) {
  P2V_REQUIRE(in && out && idx, "p2v_gather_row_segments: null pointer");
  P2V_REQUIRE(images > 0 && rows_in > 0 && rows_out > 0 && segs > 0 && seg_bytes > 0 && seg_bytes % 16 == 0,
              "p2v_gather_row_segments: bad shape (segments are multiples of 16 bytes)");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "p2v_gather_row_segments: buffers must be 16-byte aligned");
  const int chunks = seg_bytes / 16;
  const int64_t total = (int64_t)images * rows_out * segs * chunks;
  gather_row_segments_kernel<<<grid_for_work(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, idx, rows_in, rows_out,
                                                                                         segs, chunks, total);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_avgpool_requant(const int8_t* in, int8_t* out, int images, int tokens, int channels, float in_scale,
                                   float out_scale, float out_zp, void* stream) {
  P2V_REQUIRE(in && out, "p2v_avgpool_requant: null pointer");
  P2V_REQUIRE(images > 0 && tokens > 0 && tokens <= 65536 && channels > 0, "p2v_avgpool_requant: bad shape");
  int ex = 0;
  P2V_REQUIRE(in_scale > 0.f && frexpf(in_scale, &ex) == 0.5f, "p2v_avgpool_requant: in_scale must be a power of two");
  const int64_t total = (int64_t)images * channels;
  avgpool_requant_kernel<<<grid_for_work(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, tokens, channels, total,
                                                                                     in_scale, out_scale, out_zp);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

// ---- the whole Swin forward as one C call ---------------------------------------------------------------------------
// Stateless: the descriptor points at the caller's device tensors (weights, per-channel vectors, permutations, tables),
// the workspace is the caller's, every launch goes to the caller's stream - so the call can be captured into a CUDA
// graph by the caller (diff_vit_b200/swin_engine.py does).  Same launch sequence as SwinIntegerEngine._run, which
// stays as the code-dump path of the tests.
namespace {

inline int64_t align_up(int64_t v) { return (v + 1023) & ~(int64_t)1023; }

struct SwinLayout {
  int64_t patches, pe, stream[3], ln, qkv, att, hidden, cat, cat_ln, pooled, total;
};

// Largest extent of every buffer over the stages (rows * channels halves from stage to stage; the hidden width is
// whatever fc1 says)
int swin_layout(const p2v_swin_desc* d, int b, SwinLayout* L) {
  P2V_REQUIRE(d && d->stages && d->num_stages > 0 && b > 0, "p2v_swin: bad descriptor");
  P2V_REQUIRE(d->patch_size > 0 && d->img_size % d->patch_size == 0, "p2v_swin: image / patch size");
  const int grid = d->img_size / d->patch_size;
  int64_t rows0 = (int64_t)b * grid * grid, act = 0, qkv = 0, hid = 0, cat = 0;
  for (int s = 0; s < d->num_stages; ++s) {
    const p2v_swin_stage_desc& st = d->stages[s];
    P2V_REQUIRE(st.blocks && st.depth > 0 && st.dim > 0 && st.height > 0 && st.width > 0, "p2v_swin: stage %d", s);
    const int64_t rows = (int64_t)b * st.height * st.width;
    act = act > rows * st.dim ? act : rows * st.dim;
    qkv = qkv > rows * 3 * st.dim ? qkv : rows * 3 * st.dim;
    for (int j = 0; j < st.depth; ++j) hid = hid > rows * st.blocks[j].fc1.n ? hid : rows * st.blocks[j].fc1.n;
    if (st.has_merge) {
      cat = cat > rows * st.dim ? cat : rows * st.dim;
      act = act > rows / 4 * st.reduction.n ? act : rows / 4 * st.reduction.n;
    }
  }
  act = act > rows0 * d->embed_dim ? act : rows0 * d->embed_dim;
  int64_t off = 0;
  auto take = [&off](int64_t bytes) { const int64_t o = off; off += align_up(bytes); return o; };
  L->patches = take(rows0 * d->in_chans * d->patch_size * d->patch_size);
  L->pe = take(rows0 * d->embed_dim);
  for (int i = 0; i < 3; ++i) L->stream[i] = take(act);
  L->ln = take(act);
  L->qkv = take(qkv);
  L->att = take(act);
  L->hidden = take(hid);
  L->cat = take(cat);
  L->cat_ln = take(cat);
  L->pooled = take((int64_t)b * d->stages[d->num_stages - 1].dim);
  L->total = off;
  return P2V_OK;
}

int swin_gemm(const int8_t* a, const p2v_linear_desc& lin, int8_t* out, int64_t m, const int8_t* residual, float* out_f32,
              void* st) {
  p2v_epilogue e = lin.epi;
  e.residual = residual;
  e.aux_codes = nullptr;
  e.out_f32 = out_f32;
  e.flags = (e.flags & ~(P2V_EPI_RESIDUAL | P2V_EPI_OUT_F32)) | (residual ? P2V_EPI_RESIDUAL : 0u) | (out_f32 ? P2V_EPI_OUT_F32 : 0u);
  P2V_REQUIRE(m < (1ll << 31), "p2v_swin: %lld rows", (long long)m);
  return p2v_gemm_i8(a, lin.k, lin.w, out, lin.n, (int)m, lin.n, lin.k, &e, st);
}

}  // namespace

#define P2V_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != P2V_OK) return _rc; \
  } while (0)

extern "C" int64_t p2v_swin_workspace_bytes(const p2v_swin_desc* d, int b) {
  SwinLayout L;
  if (swin_layout(d, b, &L) != P2V_OK) return -1;
  return L.total;
}

extern "C" int p2v_swin_launches_per_forward(const p2v_swin_desc* d) {
  if (!d || !d->stages) return -1;
  int n = 3 + 3;   // patchify, patch GEMM, LN | final LN, avgpool, head
  for (int s = 0; s < d->num_stages; ++s) n += 7 * d->stages[s].depth + (d->stages[s].has_merge ? 3 : 0);
  return n;
}

static int swin_forward_impl(const p2v_swin_desc* d, const float* x, const uint8_t* x8, const float* mean, const float* stdv,
                             float* logits, int8_t* logit_codes, int b, void* workspace, void* stream) {
  P2V_REQUIRE((x || x8) && logits && logit_codes && workspace, "p2v_swin_forward: null pointer");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "p2v_swin_forward: workspace must be 1024-byte aligned");
  SwinLayout L;
  P2V_TRY(swin_layout(d, b, &L));
  int8_t* ws = static_cast<int8_t*>(workspace);
  int8_t* patches = ws + L.patches;
  int8_t* pe = ws + L.pe;
  int8_t* ln = ws + L.ln;
  int8_t* qkv = ws + L.qkv;
  int8_t* att = ws + L.att;
  int8_t* hidden = ws + L.hidden;
  const int grid = d->img_size / d->patch_size;
  int64_t rows = (int64_t)b * grid * grid;
  if (x8 != nullptr)
    P2V_TRY(p2v_quant_patchify_u8(x8, patches, b, d->in_chans, d->img_size, d->img_size, d->patch_size, d->input_scale, 0.f,
                                  mean, stdv, stream));
  else
    P2V_TRY(p2v_quant_patchify(x, patches, b, d->in_chans, d->img_size, d->img_size, d->patch_size, d->input_scale, 0.f, stream));
  P2V_TRY(swin_gemm(patches, d->patch_embed, pe, rows, nullptr, nullptr, stream));
  int cur = 0;                                        // which of the three stream buffers holds the residual stream
  int8_t* xs = ws + L.stream[cur];
  P2V_TRY(p2v_layernorm_int(pe, d->embed_dim, xs, nullptr, (int)rows, d->embed_dim, &d->pe_norm, stream));
  for (int s = 0; s < d->num_stages; ++s) {
    const p2v_swin_stage_desc& st = d->stages[s];
    const int C = st.dim;
    rows = (int64_t)b * st.height * st.width;
    for (int j = 0; j < st.depth; ++j) {
      const p2v_swin_block_desc& blk = st.blocks[j];
      int8_t* x1 = ws + L.stream[(cur + 1) % 3];
      int8_t* xn = ws + L.stream[(cur + 2) % 3];
      P2V_TRY(p2v_layernorm_int(xs, C, ln, nullptr, (int)rows, C, &blk.norm1, stream));
      P2V_TRY(swin_gemm(ln, blk.qkv, qkv, rows, nullptr, nullptr, stream));
      p2v_window_attention wa = blk.attn;
      wa.dump_a1 = wa.dump_a2 = nullptr;
      wa.dump_softmax = nullptr;
      P2V_TRY(p2v_window_attention_int(qkv, att, b, &wa, stream));
      P2V_TRY(swin_gemm(att, blk.proj, x1, rows, xs, nullptr, stream));             // + shortcut, qact2
      P2V_TRY(p2v_layernorm_int(x1, C, ln, nullptr, (int)rows, C, &blk.norm2, stream));   // LN2 .. mlp.qact0
      P2V_TRY(swin_gemm(ln, blk.fc1, hidden, rows, nullptr, nullptr, stream));
      P2V_TRY(swin_gemm(hidden, blk.fc2, xn, rows, x1, nullptr, stream));           // + residual, qact4
      cur = (cur + 2) % 3;
      xs = xn;
    }
    if (st.has_merge) {
      int8_t* cat = ws + L.cat;
      int8_t* cat_ln = ws + L.cat_ln;
      int8_t* xn = ws + L.stream[(cur + 1) % 3];
      const int tokens = st.height * st.width;
      P2V_TRY(p2v_gather_row_segments(xs, cat, st.merge_idx, b, tokens, tokens / 4, 4, C, stream));
      P2V_TRY(p2v_layernorm_int(cat, 4 * C, cat_ln, nullptr, (int)(rows / 4), 4 * C, &st.merge_norm, stream));
      P2V_TRY(swin_gemm(cat_ln, st.reduction, xn, rows / 4, nullptr, nullptr, stream));
      cur = (cur + 1) % 3;
      xs = xn;
    }
  }
  const p2v_swin_stage_desc& last = d->stages[d->num_stages - 1];
  const int C = last.has_merge ? last.reduction.n : last.dim;
  const int tokens = last.has_merge ? last.height * last.width / 4 : last.height * last.width;
  rows = (int64_t)b * tokens;
  int8_t* pooled = ws + L.pooled;
  P2V_TRY(p2v_layernorm_int(xs, C, ln, nullptr, (int)rows, C, &d->norm, stream));
  P2V_TRY(p2v_avgpool_requant(ln, pooled, b, tokens, C, d->pool_in_scale, d->pool_out_scale, 0.f, stream));
  P2V_TRY(swin_gemm(pooled, d->head, logit_codes, b, nullptr, logits, stream));
  return P2V_OK;
}

extern "C" int p2v_swin_forward(const p2v_swin_desc* d, const float* x, float* logits, int8_t* logit_codes, int b,
                                void* workspace, void* stream) {
  P2V_REQUIRE(x != nullptr, "p2v_swin_forward: null pointer");
  return swin_forward_impl(d, x, nullptr, nullptr, nullptr, logits, logit_codes, b, workspace, stream);
}

extern "C" int p2v_swin_forward_u8(const p2v_swin_desc* d, const uint8_t* x, const float* mean, const float* stdv,
                                   float* logits, int8_t* logit_codes, int b, void* workspace, void* stream) {
  P2V_REQUIRE(x && mean && stdv, "p2v_swin_forward_u8: null pointer");
  return swin_forward_impl(d, nullptr, x, mean, stdv, logits, logit_codes, b, workspace, stream);
}
