// p2v_attention.cu - fused integer attention for one layer:
//   S = Q K^T (int32) -> qact_attn1 int8 codes -> log-int-softmax 4-bit log2 codes -> P V -> qact2 codes
// with the [n, n] score matrix never leaving the SM.
//
// Replaces models/vit_fquant.py:308-326 and QIntSoftmax.forward (models/ptq/layers.py:323-376).
//
// Integer formulation (probed bit-exact against the reference):
//   score code   sc = clamp(RNE(acc * 2^(2 e_qkv - 3 - e_score)))          acc = sum_d q*k   (|acc| < 2^20)
//   integer exp  e(d) for d = rowmax - sc in [0, 255]: a 256-entry table built on the host with the
//                reference's own expressions; row sum exact in 64-bit integers
//   log2 code    k = log_round(RNE(sum / e)) in [0, 15], 16 = probability 0
//   AV           acc2 = sum_j v_j * 2^(15 - k_j)  (int32), out = clamp(RNE(acc2 * 2^(e_qkv - 15 - e_out)))
// The probabilities 2^(15-k) span 16 bits, so P is split into two u8 planes (k <= 7 -> 2^(7-k) in units of
// 2^8; k >= 8 -> 2^(15-k)) and the AV product is two u8 x s8 tensor-core MMAs recombined as (hi << 8) + lo.
//
// One CTA = one (image, head) x 112 query rows; one warp = 16 query rows, all keys.  Both products use
// mma.sync m16n8k32 (IMMA): the score fragment layout (row g, cols 2t,2t+1 of each 8-key tile) is re-used
// directly as the A operand of the AV product by permuting the key order of V when it is transposed into
// shared memory, so the 4-bit codes never round-trip through memory.
#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

constexpr int kHd = 64;            // head dim (all DeiT/ViT configs of the reference)
constexpr int kAttWarps = 7;
constexpr int kAttRows = kAttWarps * 16;  // 112 query rows per CTA
constexpr int kMaxKeys = 224;      // keys padded to a multiple of 32
constexpr int kMaxTiles = kMaxKeys / 8;
constexpr int kQKStride = 80;      // bytes per Q/K row in smem (64 + pad, conflict-free fragment loads)
constexpr int kVtStride = 240;     // bytes per V^T row in smem (224 + pad)

__device__ __forceinline__ void mma_s8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// position of key j inside the permuted K dimension of the AV product (see file header)
__device__ __forceinline__ int av_perm(int j) {
  const int s = j >> 5, jj = j & 31;
  const int ti = jj >> 3, col = jj & 7;
  const int t = col >> 1, i0 = col & 1;
  const int h = ti >> 1, i = ((ti & 1) << 1) | i0;
  return (s << 5) + (h << 4) + (t << 2) + i;
}

__global__ void __launch_bounds__(kAttWarps * 32)
attention_int_kernel(const int8_t* __restrict__ qkv, int8_t* __restrict__ out, int n, int heads,
                     const p2v_attention p) {
  __shared__ __align__(16) uint8_t Ks[kMaxKeys * kQKStride];
  __shared__ __align__(16) uint8_t Qs[kAttRows * kQKStride];
  __shared__ __align__(16) uint8_t Vt[kHd * kVtStride];
  __shared__ float lut_f[256];
  __shared__ unsigned long long lut_i[256];

  const int bh = blockIdx.x;
  const int img = bh / heads, head = bh % heads;
  const int row_base = blockIdx.y * kAttRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nkp = (n + 31) & ~31;   // padded key count
  const int ntiles = nkp >> 3;
  const int64_t row_stride = (int64_t)3 * heads * kHd;
  const int8_t* base = qkv + (int64_t)img * n * row_stride;

  // ---- stage Q, K (row-major, padded stride) and V (transposed + permuted) in shared memory ----------
  for (int i = tid; i < 256; i += blockDim.x) {
    const float e = p.exp_lut[i];
    lut_f[i] = e;
    lut_i[i] = (unsigned long long)e;
  }
  for (int i = tid; i < nkp * 4; i += blockDim.x) {
    const int j = i >> 2, part = i & 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < n) v = __ldg(reinterpret_cast<const uint4*>(base + j * row_stride + (heads + head) * kHd) + part);
    *reinterpret_cast<uint4*>(Ks + j * kQKStride + part * 16) = v;
  }
  for (int i = tid; i < kAttRows * 4; i += blockDim.x) {
    const int r = i >> 2, part = i & 3;
    const int row = row_base + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < n) v = __ldg(reinterpret_cast<const uint4*>(base + row * row_stride + head * kHd) + part);
    *reinterpret_cast<uint4*>(Qs + r * kQKStride + part * 16) = v;
  }
  for (int i = tid; i < nkp * 4; i += blockDim.x) {
    const int j = i >> 2, part = i & 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < n) v = __ldg(reinterpret_cast<const uint4*>(base + j * row_stride + (2 * heads + head) * kHd) + part);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const int pj = av_perm(j);
#pragma unroll
    for (int b = 0; b < 16; ++b) Vt[(part * 16 + b) * kVtStride + pj] = (uint8_t)((w[b >> 2] >> (8 * (b & 3))) & 0xff);
  }
  __syncthreads();

  const int r0 = row_base + warp * 16;
  if (r0 >= n) return;  // warp-uniform; no block-level sync follows

  // ---- S = Q K^T, re-quantized to int8 score codes, packed 4 per register ----------------------------
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint8_t* q0 = Qs + (warp * 16 + g) * kQKStride + ks * 32 + t * 4;
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(q0);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(q0 + 8 * kQKStride);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(q0 + 16);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(q0 + 8 * kQKStride + 16);
  }
  uint32_t codeA[kMaxTiles / 2], codeB[kMaxTiles / 2];  // rows g and g+8
  int maxA = -128, maxB = -128;
#pragma unroll
  for (int j = 0; j < kMaxTiles; ++j) {
    if ((j & 1) == 0) { codeA[j >> 1] = 0; codeB[j >> 1] = 0; }
    if (j < ntiles) {
      int c[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint8_t* kp = Ks + (j * 8 + g) * kQKStride + ks * 32 + t * 4;
        mma_s8s8(c, qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 16));
      }
      int sc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        sc[e] = clamp_i(rne(fadd(fmul((float)c[e], p.score_mul), p.score_zp)), -128, 127);
      const int col = j * 8 + t * 2;
      if (col < n) { maxA = max(maxA, sc[0]); maxB = max(maxB, sc[2]); }
      if (col + 1 < n) { maxA = max(maxA, sc[1]); maxB = max(maxB, sc[3]); }
      const int sh = (j & 1) * 16;
      codeA[j >> 1] |= ((uint32_t)(sc[0] & 0xff) | ((uint32_t)(sc[1] & 0xff) << 8)) << sh;
      codeB[j >> 1] |= ((uint32_t)(sc[2] & 0xff) | ((uint32_t)(sc[3] & 0xff) << 8)) << sh;
    }
  }
  maxA = max(maxA, __shfl_xor_sync(0xffffffffu, maxA, 1));
  maxA = max(maxA, __shfl_xor_sync(0xffffffffu, maxA, 2));
  maxB = max(maxB, __shfl_xor_sync(0xffffffffu, maxB, 1));
  maxB = max(maxB, __shfl_xor_sync(0xffffffffu, maxB, 2));

  // ---- exact integer row sums of the integer exp --------------------------------------------------------
  unsigned long long sumA = 0, sumB = 0;
#pragma unroll
  for (int j = 0; j < kMaxTiles; ++j) {
    if (j < ntiles) {
      const int sh = (j & 1) * 16;
      const int col = j * 8 + t * 2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (col + e < n) {
          const int a = (int)(int8_t)((codeA[j >> 1] >> (sh + 8 * e)) & 0xff);
          const int b = (int)(int8_t)((codeB[j >> 1] >> (sh + 8 * e)) & 0xff);
          sumA += lut_i[maxA - a];
          sumB += lut_i[maxB - b];
        }
      }
    }
  }
  sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
  sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
  sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
  sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
  const float fsumA = __ull2float_rn(sumA), fsumB = __ull2float_rn(sumB);

  // ---- log2 codes -> two u8 probability planes -> P V ------------------------------------------------------
  const int rowA = r0 + g, rowB = r0 + g + 8;
  int8_t* dsc = p.dump_scores ? p.dump_scores + ((int64_t)bh * n) * n : nullptr;
  uint8_t* dsm = p.dump_softmax ? p.dump_softmax + ((int64_t)bh * n) * n : nullptr;
  int hi[8][4], lo[8][4];
#pragma unroll
  for (int jn = 0; jn < 8; ++jn)
#pragma unroll
    for (int e = 0; e < 4; ++e) { hi[jn][e] = 0; lo[jn][e] = 0; }

#pragma unroll
  for (int s = 0; s < kMaxTiles / 4; ++s) {
    if (s * 4 < ntiles) {
      uint32_t pa_hi[4] = {0, 0, 0, 0}, pa_lo[4] = {0, 0, 0, 0};  // a0..a3 of the two planes
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {        // hh = 0: tiles 4s, 4s+1 (a0/a1); hh = 1: tiles 4s+2, 4s+3 (a2/a3)
        const uint32_t wa = codeA[2 * s + hh], wb = codeB[2 * s + hh];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = (4 * s + 2 * hh + (i >> 1)) * 8 + t * 2 + (i & 1);
          if (col < n) {
            const int a = (int)(int8_t)((wa >> (8 * i)) & 0xff);
            const int b = (int)(int8_t)((wb >> (8 * i)) & 0xff);
            const int ka = softmax_log_code(fsumA, lut_f[maxA - a], p.softmax_levels);
            const int kb = softmax_log_code(fsumB, lut_f[maxB - b], p.softmax_levels);
            if (ka <= 7) pa_hi[2 * hh] |= (uint32_t)(1u << (7 - ka)) << (8 * i);
            else if (ka <= 15) pa_lo[2 * hh] |= (uint32_t)(1u << (15 - ka)) << (8 * i);
            if (kb <= 7) pa_hi[2 * hh + 1] |= (uint32_t)(1u << (7 - kb)) << (8 * i);
            else if (kb <= 15) pa_lo[2 * hh + 1] |= (uint32_t)(1u << (15 - kb)) << (8 * i);
            if (dsc != nullptr) {
              if (rowA < n) { dsc[(int64_t)rowA * n + col] = (int8_t)a; dsm[(int64_t)rowA * n + col] = (uint8_t)ka; }
              if (rowB < n) { dsc[(int64_t)rowB * n + col] = (int8_t)b; dsm[(int64_t)rowB * n + col] = (uint8_t)kb; }
            }
          }
        }
      }
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const uint8_t* vp = Vt + (jn * 8 + g) * kVtStride + s * 32 + t * 4;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vp), b1 = *reinterpret_cast<const uint32_t*>(vp + 16);
        mma_u8s8(hi[jn], pa_hi[0], pa_hi[1], pa_hi[2], pa_hi[3], b0, b1);
        mma_u8s8(lo[jn], pa_lo[0], pa_lo[1], pa_lo[2], pa_lo[3], b0, b1);
      }
    }
  }

  // ---- re-quantize and store ----------------------------------------------------------------------------
  const int64_t out_stride = (int64_t)heads * kHd;
#pragma unroll
  for (int jn = 0; jn < 8; ++jn) {
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int acc = hi[jn][e] * 256 + lo[jn][e];
      const double v = rint((double)acc * p.out_mul) + (double)p.out_zp;
      q[e] = (int)fmin(fmax(v, -128.0), 127.0);
    }
    const int col = head * kHd + jn * 8 + t * 2;
    if (rowA < n)
      *reinterpret_cast<uint16_t*>(out + ((int64_t)img * n + rowA) * out_stride + col) =
          (uint16_t)((q[0] & 0xff) | ((q[1] & 0xff) << 8));
    if (rowB < n)
      *reinterpret_cast<uint16_t*>(out + ((int64_t)img * n + rowB) * out_stride + col) =
          (uint16_t)((q[2] & 0xff) | ((q[3] & 0xff) << 8));
  }
}

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_attention_int(const int8_t* qkv, int8_t* out, int b, int n, int heads, const p2v_attention* p,
                                 void* stream) {
  P2V_REQUIRE(qkv && out && p && p->exp_lut, "p2v_attention_int: null pointer");
  P2V_REQUIRE(b > 0 && heads > 0 && n > 0, "p2v_attention_int: bad shape b=%d n=%d heads=%d", b, n, heads);
  P2V_REQUIRE(n <= kMaxKeys, "p2v_attention_int: n=%d tokens exceeds the %d-key tile of this kernel", n, kMaxKeys);
  P2V_REQUIRE((p->dump_scores == nullptr) == (p->dump_softmax == nullptr),
              "p2v_attention_int: dump_scores and dump_softmax must be given together");
  dim3 grid(b * heads, (n + kAttRows - 1) / kAttRows);
  attention_int_kernel<<<grid, kAttWarps * 32, 0, (cudaStream_t)stream>>>(qkv, out, n, heads, *p);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
