// p2v_swin.cu - the kernels the quantized Swin forward (BASELINE config 5) needs beside the GEMM / LayerNorm family:
//
//   p2v_window_attention_int   windowed / shifted-window integer attention      (models/swin_quant.py:177-221)
//   p2v_quant_patchify (p < 16) input QAct + im2col for the 4 x 4 patch conv     (models/layers_quant.py:211-224)
//   p2v_gather_row_segments    the 2 x 2 neighbourhood concat of PatchMerging   (models/swin_quant.py:445-457)
//   p2v_avgpool_requant        token average + qact3                            (models/swin_quant.py:810-813)
//
// Window attention, per (window, head): one thread owns one of the n = ws^2 <= 64 query rows.
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

namespace p2v {

constexpr int kWaMaxN = 64;     // tokens per window
constexpr int kWaHeadDim = 32;  // every Swin variant of the reference: C / heads = 32

__device__ __forceinline__ uint32_t pack4_sat_s8(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}
__device__ __forceinline__ int sx8(uint32_t w, int j) { return (int)(int8_t)((w >> (8 * j)) & 0xffu); }

__global__ void __launch_bounds__(kWaMaxN)
window_attention_kernel(const int8_t* __restrict__ qkv, int8_t* __restrict__ out, const p2v_window_attention a) {
  __shared__ int ks[kWaMaxN][kWaHeadDim];
  __shared__ int vs[kWaMaxN][kWaHeadDim];
  __shared__ int8_t xs[kWaMaxN][kWaMaxN];   // [key][row]: thread `row` only ever touches its own column
  __shared__ uint8_t rid[kWaMaxN];
  const int n = a.n, C = a.channels;
  const int head = blockIdx.x % a.heads;
  const int wg = blockIdx.x / a.heads;            // window over the whole batch
  const int img = wg / a.windows, w = wg % a.windows;
  const int i = threadIdx.x;
  const bool valid = i < n;
  int64_t src_row = 0;
  int M[kWaHeadDim];
  if (valid) {
    src_row = (int64_t)img * a.tokens + a.perm[w * n + i];
    const int8_t* base = qkv + src_row * (3 * (int64_t)C) + head * kWaHeadDim;
    const float unit = __uint_as_float((uint32_t)(127 + a.qshift) << 23);   // 2^qshift
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 q4 = *reinterpret_cast<const uint4*>(base + 16 * h);
      const uint4 k4 = *reinterpret_cast<const uint4*>(base + C + 16 * h);
      const uint4 v4 = *reinterpret_cast<const uint4*>(base + 2 * C + 16 * h);
      const uint32_t qw[4] = {q4.x, q4.y, q4.z, q4.w}, kw[4] = {k4.x, k4.y, k4.z, k4.w}, vw[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 16 * h + 4 * x + e;
          // fl32(code * c32) * 2^qshift: an integer below 2^31 (see the file header), so the conversion is exact
          M[c] = __float2int_rn(fmul(fmul((float)sx8(qw[x], e), a.qscale), unit));
          ks[i][c] = sx8(kw[x], e);
          vs[i][c] = sx8(vw[x], e);
        }
      }
    }
    rid[i] = a.region != nullptr ? a.region[w * n + i] : (uint8_t)0;
  }
  __syncthreads();
  if (!valid) return;

  const int64_t dump_row = ((int64_t)(wg * a.heads + head) * n + i) * n;
  const float* bias_h = a.bias + (int64_t)head * n * n;
  const int my_rid = rid[i];
  // ---- scores: exact product, the two re-quantisations, row maximum of the masked codes ----
  int mx = INT_MIN;
  for (int j = 0; j < n; ++j) {
    long long acc = 0;
#pragma unroll
    for (int c = 0; c < kWaHeadDim; c += 4) {
      const int4 kk = *reinterpret_cast<const int4*>(&ks[j][c]);
      acc += (long long)M[c] * kk.x;
      acc += (long long)M[c + 1] * kk.y;
      acc += (long long)M[c + 2] * kk.z;
      acc += (long long)M[c + 3] * kk.w;
    }
    const float s = __double2float_rn(__dmul_rn(__ll2double_rn(acc), a.acc_scale));   // one rounding
    const float c1 = fminf(fmaxf(rintf(fmul(s, a.a1_rscale)), -128.f), 127.f);        // qact_attn1
    const float t = fadd(fmul(c1, a.a1_scale), __ldg(bias_h + j * n + i));              // + bias (dequantized table entry)
    const int x = (int)fminf(fmaxf(rintf(fmul(t, a.a2_rscale)), -128.f), 127.f);       // qact2
    xs[j][i] = (int8_t)x;
    if (a.dump_a1 != nullptr) a.dump_a1[dump_row + j] = (int8_t)c1;
    if (a.dump_a2 != nullptr) a.dump_a2[dump_row + j] = (int8_t)x;
    const int xm = x - (rid[j] != my_rid ? a.mask_int : 0);
    mx = max(mx, xm);
  }
  // ---- row sum of the integer exp ----
  const int dmax = a.lut_n - 1;
  double sum = 0.0;
  for (int j = 0; j < n; ++j) {
    const int xm = (int)xs[j][i] - (rid[j] != my_rid ? a.mask_int : 0);
    sum += (double)__ldg(a.exp_lut + min(mx - xm, dmax));
  }
  const float fsum = __double2float_rn(sum);
  // ---- log2 codes, P V ----
  int O[kWaHeadDim];
#pragma unroll
  for (int c = 0; c < kWaHeadDim; ++c) O[c] = 0;
  for (int j = 0; j < n; ++j) {
    const int xm = (int)xs[j][i] - (rid[j] != my_rid ? a.mask_int : 0);
    const float e = __ldg(a.exp_lut + min(mx - xm, dmax));
    const int k = softmax_log_code(fsum, e, a.softmax_levels);
    if (a.dump_softmax != nullptr) a.dump_softmax[dump_row + j] = (uint8_t)k;
    const int p = k >= a.softmax_levels ? 0 : (0x8000 >> k);
#pragma unroll
    for (int c = 0; c < kWaHeadDim; c += 4) {
      const int4 vv = *reinterpret_cast<const int4*>(&vs[j][c]);
      O[c] += p * vv.x;
      O[c + 1] += p * vv.y;
      O[c + 2] += p * vv.z;
      O[c + 3] += p * vv.w;
    }
  }
  // ---- qact3 and the store, back through the permutation ----
  uint32_t w8[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float val = fmul(__int2float_rn(O[4 * x + e]), a.out_unit);    // the fp32 value of (attn @ v): one rounding
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
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_words; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int kw4 = (int)(t % wp); t /= wp;
    const int kh = (int)(t % p); t /= p;
    const int ch = (int)(t % c); t /= c;
    const int px = (int)(t % gw); t /= gw;
    const int py = (int)(t % gh);
    const int64_t img = t / gh;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((img * c + ch) * h + (py * p + kh)) * (int64_t)w + px * p + kw4 * 4));
    reinterpret_cast<uint32_t*>(codes)[i] =
        pack4_sat_s8(quant_div(v.x, scale, zp, -128, 127), quant_div(v.y, scale, zp, -128, 127),
                     quant_div(v.z, scale, zp, -128, 127), quant_div(v.w, scale, zp, -128, 127));
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

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_window_attention_int(const int8_t* qkv, int8_t* out, int images, const p2v_window_attention* p,
                                        void* stream) {
  P2V_REQUIRE(qkv && out && p && p->perm && p->bias && p->exp_lut, "p2v_window_attention_int: null pointer");
  P2V_REQUIRE(images > 0 && p->n > 0 && p->n <= kWaMaxN && p->heads > 0 && p->windows > 0,
              "p2v_window_attention_int: bad shape images=%d n=%d heads=%d windows=%d", images, p->n, p->heads, p->windows);
  P2V_REQUIRE(p->channels == p->heads * kWaHeadDim, "p2v_window_attention_int: channels=%d is not heads * 32", p->channels);
  P2V_REQUIRE(p->tokens == p->windows * p->n, "p2v_window_attention_int: tokens=%d is not windows * n", p->tokens);
  P2V_REQUIRE(p->lut_n >= 1 && p->qshift >= 0 && p->qshift <= 60 && p->softmax_levels >= 1 && p->softmax_levels <= 16,
              "p2v_window_attention_int: bad table / shift / levels");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "p2v_window_attention_int: qkv and out must be 16-byte aligned");
  const int64_t items = (int64_t)images * p->windows * p->heads;
  P2V_REQUIRE(items < (1ll << 31), "p2v_window_attention_int: too many (window, head) items");
  window_attention_kernel<<<(unsigned)items, kWaMaxN, 0, (cudaStream_t)stream>>>(qkv, out, *p);
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
