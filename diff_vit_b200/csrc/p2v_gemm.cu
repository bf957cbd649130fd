// p2v_gemm.cu - int8 x int8 -> int32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8) with the
// P2-ViT re-quantization epilogues fused in.
//
//   out[m, n] = epilogue( sum_k a[m, k] * w[n, k] )
//
// Replaces QLinear.forward / QConv2d.forward in quantized mode followed by the QAct that consumes the
// result (reference: models/ptq/layers.py:82-88,171-178,207-220).  Both operands are K-major int8, moved
// by TMA into 128-byte-swizzled shared-memory tiles; the accumulator lives in TMEM; one elected thread
// issues the MMAs; four epilogue warps read TMEM back with tcgen05.ld and apply
// scale/bias/GELU/re-quantize/residual in the reference's fp32 operation order (p2v_math.cuh).
//
// Kernel organisation (persistent, one CTA per SM):
//   warp 0      TMA producer      : A/B k-blocks into a 4-stage smem ring (full/empty mbarriers)
//   warp 1      MMA issuer        : 4 x (M128 x N128 x K32) per k-block into one of 2 TMEM accumulators
//   warp 2      TMEM allocator
//   warps 4-19  epilogue          : 16 warps, each one TMEM lane quarter x 32 columns: TMEM -> registers ->
//                                   int8 codes -> global, overlapped with the next tile's MMAs through the
//                                   second accumulator (tmem full/empty mbarriers)
#include <cuda.h>

#include "p2v_common.cuh"
#include "p2v_math.cuh"
#include "p2v_requant.cuh"

namespace p2v {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kBlockK = 128;  // bytes == int8 elements: one 128B swizzle row
constexpr int kUmmaK = 32;    // K per tcgen05.mma for 8-bit operands
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTileBytes = kBlockM * kBlockK;  // 16 KiB per operand per stage
constexpr int kEpiWarps = 16;                  // 4 TMEM lane quarters x 4 column groups of 32
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = 128 + kEpiThreads;
constexpr uint32_t kTmemCols = kAccStages * kBlockN;  // 256 columns of 32-bit accumulators
constexpr int kChanCols = 512;                        // per-channel constants kept in shared memory

// per-channel epilogue constants staged in shared memory, one array per field (broadcast float4 reads)
enum { CH_A = 0, CH_B, CH_RSO, CH_SO, CH_SR, CH_RSO2, CH_SO2, CH_FIELDS };

struct GemmSmem {
  alignas(1024) uint8_t a[kStages][kTileBytes];
  alignas(1024) uint8_t b[kStages][kTileBytes];
  alignas(16) float chan[CH_FIELDS][kChanCols];   // whole n when n <= kChanCols, else two 128-column slots
  alignas(16) uint8_t stage[4][2][32 * 144];   // per TMEM lane quarter, double-buffered (see epilogue_block)
  alignas(8) uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
};

struct GemmArgs {
  int8_t* out;
  int64_t ld_out;
  int m, n, k;
  p2v_epilogue epi;
  int32_t* raw_acc;  // test hook: dump accumulators instead of codes
};

__device__ __forceinline__ float clamp_code(float r) { return fminf(fmaxf(r, -128.f), 127.f); }

// Internal template bit (not part of the C-ABI flags): the launch has no dump target (raw_acc, aux_codes, out_f32),
// n and ld_out are multiples of 16 and out / residual are 16-byte aligned - what every GEMM of the fused forward
// looks like.  The host checks this once per launch; the kernels then carry none of the per-chunk pointer tests,
// ragged-edge masks and 64-bit dump addressing (ncu: 60 % of the light epilogue's instructions were such overhead).
constexpr uint32_t EPI_PLAIN = 16u;

// ---- erf-GELU straight to the output grid -----------------------------------------------------------------------
// libdevice's erff selects between two polynomial sets per element (9 FSEL + 10 FMA-pipe ops + MUFU): with the
// surrounding arithmetic 32 instructions per output of the fc1 epilogue, which is issue-bound (ncu: 80 %).
// Here erfc(t) = 2^(t Q(t)), t = min(|x|, 4), Q a degree-6 minimax fit of log2(erfc(t)) / t weighted by
// erfc(t): |error| < 5e-7 absolute including MUFU.EX2, one polynomial, no selects.  The result on the output grid
//   tq = (0.5 y rso) (1 + erf(y / sqrt 2)) = hr + |hr| (1 - erfc(|y| / sqrt 2)),  hr = 0.5 y rso
// is accepted only if it lies further than 1.5e-6 |0.5 y rso| from a rounding boundary - more than the difference
// to the reference expression gelu_erf(y) * rso (approximation error plus a few ulp of association) - so accepted
// elements round to the same int8 code as the reference expression; the others (~2e-4) are redone with gelu_erf.
// p2v_test_gelu_fast sweeps all 2^32 inputs on the device (as the pairs the epilogue forms) and counts accepted
// elements whose codes differ: zero.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two adjacent outputs: the operations whose operands are registers anyway go through the packed instructions, the
// polynomial keeps the scalar immediate-operand FFMA form.  tq: the value on the output grid; lo / hi: the same moved
// down / up by the error bound 1.5e-6 |0.5 y rso|.  The caller packs lo and hi to int8 codes: where they agree no
// rounding boundary (and no saturation edge) lies inside the bound, so the reference expression rounds to that code
// too; where they differ the element is redone with gelu_erf.  (The first version tested |tq - RNE(tq)| against the
// bound per element: five instructions per output against 1.75 for the two extra FFMA2 halves, the second pack and
// the word compare.)
__device__ __forceinline__ void gelu_code_fast2(const float (&y)[2], const float (&half_rso)[2], float (&lo)[2], float (&hi)[2]) {
  // Q's coefficients with the 1 / sqrt 2 of x = y / sqrt 2 folded in: t Q(t) = u Q'(u), u = min(|y|, 4 sqrt 2)
  constexpr double c = 0.70710678118654752440;
  constexpr float k0 = (float)(-1.6279137134552002 * c), k1 = (float)(-9.183286428451538e-1 * c * c),
                  k2 = (float)(-1.4896366000175476e-1 * c * c * c), k3 = (float)(2.9452499002218246e-2 * c * c * c * c),
                  k4 = (float)(-2.3022270761430264e-3 * c * c * c * c * c),
                  k5 = (float)(-4.6157639008015394e-4 * c * c * c * c * c * c),
                  k6 = (float)(1.0022142669185996e-4 * c * c * c * c * c * c * c);
  // the polynomial on both elements at once: six packed FMAs with immediate coefficients instead of twelve scalar ones
  // (a packed FP32 instruction issues every other cycle, a scalar one at 0.69 per cycle: tools/ubench/pipes.cu)
  const float2 u = make_float2(fminf(fabsf(y[0]), 5.6568542494923802f), fminf(fabsf(y[1]), 5.6568542494923802f));
  float2 q = ffma2(make_float2(k6, k6), u, make_float2(k5, k5));
  q = ffma2(q, u, make_float2(k4, k4));
  q = ffma2(q, u, make_float2(k3, k3));
  q = ffma2(q, u, make_float2(k2, k2));
  q = ffma2(q, u, make_float2(k1, k1));
  q = ffma2(q, u, make_float2(k0, k0));
  const float2 uq = fmul2(u, q);
  const float E[2] = {ex2_approx(uq.x), ex2_approx(uq.y)};   // erfc(|y| / sqrt 2)
  // hr (1 + erf(x)) = hr + |hr| (1 - E): sign(hr) = sign(x) because rso > 0.  With nah = -|hr| (one LOP3 per element)
  // everything else is packed FMAs: base = hr + |hr|, tq = base - |hr| E (one rounding), lo / hi = tq -+ 1.5e-6 |hr|.
  const float2 hr = fmul2(make_float2(y[0], y[1]), make_float2(half_rso[0], half_rso[1]));
  const float2 nah = make_float2(u2f(f2u(hr.x) | 0x80000000u), u2f(f2u(hr.y) | 0x80000000u));
  const float2 base = ffma2(nah, make_float2(-1.0f, -1.0f), hr);
  const float2 t2 = ffma2(nah, make_float2(E[0], E[1]), base);
  const float2 l2 = ffma2(nah, make_float2(1.5e-6f, 1.5e-6f), t2);
  const float2 h2 = ffma2(nah, make_float2(-1.5e-6f, -1.5e-6f), t2);
  lo[0] = l2.x; lo[1] = l2.y;
  hi[0] = h2.x; hi[1] = h2.y;
}

// Epilogue of 16 consecutive columns of one output row (one thread).  ch: this accumulator stage's
// channel constants, c: column offset inside the tile.
// Staged mode (res_staged / out_staged non-null): the residual codes of these 16 columns were brought in, and the
// packed result leaves, through a shared-memory tile so that global traffic is fully coalesced.
template <uint32_t FLAGS, int CW>
__device__ __forceinline__ void epilogue16(const uint32_t (&acc)[16], const float (*ch)[CW], int c,
                                           const GemmArgs& g, int row, int col0, const uint4* res_staged = nullptr,
                                           uint4* out_staged = nullptr) {
  constexpr bool kPlain = (FLAGS & EPI_PLAIN) != 0;
  const int ncols = kPlain ? 16 : min(16, g.n - col0);
  if (kPlain) {
    if (col0 >= g.n) return;   // n is a multiple of 16: a chunk is entirely inside or outside (its staged bytes are never stored)
  } else {
    if (out_staged != nullptr) *out_staged = make_uint4(0, 0, 0, 0);
    if (ncols <= 0) return;
  }
  const int64_t off = (int64_t)row * g.ld_out + col0;
  if (!kPlain && g.raw_acc != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < ncols) g.raw_acc[(int64_t)row * g.n + col0 + j] = (int32_t)acc[j];
    return;
  }
  const bool vec = kPlain || ((ncols == 16) && ((g.ld_out & 15) == 0) && ((col0 & 15) == 0));
  constexpr bool kFold = (FLAGS & EPI_OUT_POT) && !(FLAGS & EPI_GELU);  // scale and bias pre-multiplied by 1/s_out
  float code[16];   // first-stage codes (clamped, integral)
  uint32_t resw[4] = {0, 0, 0, 0};
  if (FLAGS & EPI_RESIDUAL) {
    if (res_staged != nullptr) {
      resw[0] = res_staged->x; resw[1] = res_staged->y; resw[2] = res_staged->z; resw[3] = res_staged->w;
    } else if (vec) {
      const uint4 r4 = ld_act_v4(g.epi.residual + off);
      resw[0] = r4.x; resw[1] = r4.y; resw[2] = r4.z; resw[3] = r4.w;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < ncols) resw[j >> 2] |= (uint32_t)(uint8_t)g.epi.residual[off + j] << (8 * (j & 3));
    }
  }
#pragma unroll
  for (int j4 = 0; j4 < 16; j4 += 4) {
    const float4 A = *reinterpret_cast<const float4*>(&ch[CH_A][c + j4]);
    const float4 B = *reinterpret_cast<const float4*>(&ch[CH_B][c + j4]);
    const float a4[4] = {A.x, A.y, A.z, A.w}, b4[4] = {B.x, B.y, B.z, B.w};
    float rso[4] = {1.f, 1.f, 1.f, 1.f}, so[4] = {1.f, 1.f, 1.f, 1.f};
    if (!kFold) {
      const float4 R = *reinterpret_cast<const float4*>(&ch[CH_RSO][c + j4]);
      rso[0] = R.x; rso[1] = R.y; rso[2] = R.z; rso[3] = R.w;
    }
    if (!(FLAGS & EPI_OUT_POT) || (FLAGS & EPI_RESIDUAL)) {
      const float4 S = *reinterpret_cast<const float4*>(&ch[CH_SO][c + j4]);
      so[0] = S.x; so[1] = S.y; so[2] = S.z; so[3] = S.w;
    }
    float y4[4], r4[4];
    constexpr bool kGeluPot = (FLAGS & EPI_GELU) && (FLAGS & EPI_OUT_POT);
#pragma unroll
    for (int e = 0; e < 4; e += 2) {   // acc * scale + bias, two columns per fma.rn.f32x2 (kFold: already on the output grid)
      const float2 yy = ffma2(make_float2((float)(int)acc[j4 + e], (float)(int)acc[j4 + e + 1]),
                              make_float2(a4[e], a4[e + 1]), make_float2(b4[e], b4[e + 1]));
      y4[e] = yy.x;
      y4[e + 1] = yy.y;
    }
    if ((FLAGS & EPI_GELU) && !kGeluPot) {
#pragma unroll
      for (int e = 0; e < 4; ++e) y4[e] = gelu_erf(y4[e]);
    }
    if (kGeluPot) {
      float lo4[4], hi4[4];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {   // rso[] holds rso / 2 here
        const float yy[2] = {y4[e], y4[e + 1]}, hh[2] = {rso[e], rso[e + 1]};
        float l2[2], h2[2];
        gelu_code_fast2(yy, hh, l2, h2);
        lo4[e] = l2[0]; lo4[e + 1] = l2[1];
        hi4[e] = h2[0]; hi4[e + 1] = h2[1];
        r4[e] = h2[0];
        r4[e + 1] = h2[1];
      }
      if (pack_sat4(lo4[0], lo4[1], lo4[2], lo4[3]) != pack_sat4(hi4[0], hi4[1], hi4[2], hi4[3])) {   // rare (~1e-3 of the groups)
#pragma unroll
        for (int e = 0; e < 4; ++e) r4[e] = fmul(gelu_erf(y4[e]), fmul(rso[e], 2.0f));
      }
    } else if (kFold) {
#pragma unroll
      for (int e = 0; e < 4; ++e) r4[e] = y4[e];                       // the saturating pack rounds half-even
    } else if (FLAGS & EPI_OUT_POT) {
#pragma unroll
      for (int e = 0; e < 4; ++e) r4[e] = fmul(y4[e], rso[e]);         // idem
    } else {
      div_round4(y4, so, rso, g.epi.out_zp, r4);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // residual variants need the integral, clamped branch code now; otherwise the saturating pack rounds
      if ((FLAGS & EPI_RESIDUAL) && (FLAGS & EPI_OUT_POT)) r4[e] = rne_small(r4[e]);
      code[j4 + e] = (FLAGS & EPI_RESIDUAL) ? clamp_code(r4[e]) : r4[e];
    }
    if (FLAGS & EPI_RESIDUAL) {
      const float4 SR = *reinterpret_cast<const float4*>(&ch[CH_SR][c + j4]);
      const float4 R2 = *reinterpret_cast<const float4*>(&ch[CH_RSO2][c + j4]);
      const float4 S2 = *reinterpret_cast<const float4*>(&ch[CH_SO2][c + j4]);
      const float sr[4] = {SR.x, SR.y, SR.z, SR.w}, r2[4] = {R2.x, R2.y, R2.z, R2.w}, s2[4] = {S2.x, S2.y, S2.z, S2.w};
      const uint32_t w = resw[j4 >> 2];
      const float res[4] = {(float)(int8_t)(w & 0xff), (float)(int8_t)((w >> 8) & 0xff),
                            (float)(int8_t)((w >> 16) & 0xff), (float)(int8_t)(w >> 24)};
      if (!kPlain && (g.epi.aux_codes != nullptr || g.epi.out_f32 != nullptr)) {   // dump of the branch codes (qact3 / mlp.qact2)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j4 + e;
          if (j < ncols) {
            if (g.epi.aux_codes != nullptr) g.epi.aux_codes[off + j] = (int8_t)code[j];
            if (g.epi.out_f32 != nullptr)
              g.epi.out_f32[(int64_t)row * g.n + col0 + j] = fmul(fsub(code[j], g.epi.out_zp), so[e]);
          }
        }
      }
      // fp32: dequantize residual stream and branch, add, re-quantize on the block-level grid
      float sum[4], q2[4];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        // packed products, scalar adds: ptxas contracts a packed mul feeding a packed add into one FFMA2 even with
        // explicit .rn (seen in the SASS), which would round once where the reference rounds twice
        const float2 pr = fmul2(make_float2(res[e], res[e + 1]), make_float2(sr[e], sr[e + 1]));
        const float2 pc = fmul2(make_float2(code[j4 + e], code[j4 + e + 1]), make_float2(so[e], so[e + 1]));
        sum[e] = fadd(pr.x, pc.x);
        sum[e + 1] = fadd(pr.y, pc.y);
      }
      div_round4(sum, s2, r2, 0.f, q2);
#pragma unroll
      for (int e = 0; e < 4; ++e) code[j4 + e] = q2[e];
    }
  }
  if (!kPlain && !(FLAGS & EPI_RESIDUAL) && g.epi.out_f32 != nullptr) {
    float* fo = g.epi.out_f32 + (int64_t)row * g.n + col0;
    // n % 4 == 0 (the 1000-class head): whole float4 stores, 64 contiguous bytes per row instead of 16 scalar stores
    // that each touch 32 rows (the head GEMM took 30 us of which 10 were these stores)
    const bool f4 = ((g.n & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.epi.out_f32) & 15) == 0);
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      float f[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        f[e] = fmul(fsub(clamp_code(rintf(code[j4 + e])), g.epi.out_zp), ch[CH_SO][c + j4 + e]);
      if (f4) {
        if (j4 < ncols) *reinterpret_cast<float4*>(fo + j4) = make_float4(f[0], f[1], f[2], f[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j4 + e < ncols) fo[j4 + e] = f[e];
      }
    }
  }
  if (out_staged != nullptr || vec) {
    const uint4 packed =
        make_uint4(pack_sat4(code[0], code[1], code[2], code[3]), pack_sat4(code[4], code[5], code[6], code[7]),
                   pack_sat4(code[8], code[9], code[10], code[11]), pack_sat4(code[12], code[13], code[14], code[15]));
    if (out_staged != nullptr) *out_staged = packed;
    else *reinterpret_cast<uint4*>(g.out + off) = packed;
  } else if (((g.ld_out | g.n) & 3) == 0 && (reinterpret_cast<uintptr_t>(g.out) & 3) == 0) {
    // rows are only 4-byte aligned (n = 1000): packed words, ncols is a multiple of 4
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4)
      if (j4 < ncols)
        *reinterpret_cast<uint32_t*>(g.out + off + j4) = pack_sat4(code[j4], code[j4 + 1], code[j4 + 2], code[j4 + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < ncols) g.out[off + j] = (int8_t)clamp_code(rintf(code[j]));
  }
}

// Stage the per-channel constants of one 128-column tile (epilogue warps only).
template <uint32_t FLAGS, int CW>
__device__ __forceinline__ void load_channels(float (*ch)[CW], const p2v_epilogue& e, int n0, int n, int tid,
                                              int dst0 = 0, int count = CW) {
  constexpr bool kFold = (FLAGS & EPI_OUT_POT) && !(FLAGS & EPI_GELU);
  for (int jj = tid; jj < count; jj += kEpiThreads) {
    const int col = n0 + jj;
    const int j = dst0 + jj;
    float A = 0.f, B = 0.f, RSO = 1.f, SO = 1.f, SR = 0.f, RSO2 = 1.f, SO2 = 1.f;
    if (col < n && e.acc_scale != nullptr) {
      A = e.acc_scale[col];
      B = e.bias ? e.bias[col] : 0.f;
      SO = e.out_scale[col];
      RSO = e.out_rscale ? e.out_rscale[col] : __frcp_rn(SO);
      if (kFold) {  // exact: scaling by a power of two commutes with the single rounding of acc*A + B
        A = fmul(A, RSO);
        B = fmul(B, RSO);
      }
      if ((FLAGS & EPI_GELU) && (FLAGS & EPI_OUT_POT)) RSO = fmul(RSO, 0.5f);   // gelu_code_fast2 takes rso / 2 (exact)
      if (FLAGS & EPI_RESIDUAL) {
        SR = e.res_scale[col];
        SO2 = e.out2_scale[col];
        RSO2 = __frcp_rn(SO2);
      }
    }
    ch[CH_A][j] = A; ch[CH_B][j] = B; ch[CH_RSO][j] = RSO; ch[CH_SO][j] = SO;
    ch[CH_SR][j] = SR; ch[CH_RSO2][j] = RSO2; ch[CH_SO2][j] = SO2;
  }
}

// ---- epilogue of one 32-row x 128-column block (one TMEM lane quarter, four warps) -----------------------------
// A thread owns one output row, so writing its codes straight to global memory touches 32 different cache
// lines per store instruction (ncu: L1TEX was the busiest unit of the first version).  Instead the four
// warps of a quarter exchange through a padded shared-memory tile: residual codes come in, and result codes
// go out, as 128-byte row segments (8 lanes x 16 B per row).
constexpr int kStagePitch = 144;                    // 128 B of codes + 16 B pad: conflict-free 16-byte accesses
constexpr int kStageBytes = 32 * kStagePitch;       // per quarter

__device__ __forceinline__ void quarter_barrier(int quarter) {
  asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
}

template <uint32_t FLAGS, int CW>
__device__ __forceinline__ void epilogue_block(const uint32_t (&v0)[16], const uint32_t (&v1)[16], const float (*ch)[CW],
                                               int c, const GemmArgs& g, int row0, int col0, int quarter, int cgroup,
                                               int lane, uint8_t* stage, bool staged, const uint4* pre0 = nullptr,
                                               const uint4* pre1 = nullptr) {
  const int row = row0 + lane;            // this thread's output row
  const int mycol = col0 + cgroup * 32;   // first of this thread's 32 columns
  if (!staged) {
    if (row < g.m) {
      epilogue16<FLAGS, CW>(v0, ch, c + cgroup * 32, g, row, mycol, pre0);
      epilogue16<FLAGS, CW>(v1, ch, c + cgroup * 32 + 16, g, row, mycol + 16, pre1);
    }
    return;
  }
  const int tq = cgroup * 32 + lane;      // 0..127 inside the quarter
  uint8_t* mine = stage + lane * kStagePitch + cgroup * 32;
  uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
  if (FLAGS & EPI_RESIDUAL) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tq + 128 * i, rr = idx >> 3, ck = idx & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + rr < g.m && col0 + ck * 16 < g.n)
        v = ld_act_v4(g.epi.residual + (int64_t)(row0 + rr) * g.ld_out + col0 + ck * 16);
      *reinterpret_cast<uint4*>(stage + rr * kStagePitch + ck * 16) = v;
    }
    quarter_barrier(quarter);
    r0 = *reinterpret_cast<const uint4*>(mine);
    r1 = *reinterpret_cast<const uint4*>(mine + 16);
    quarter_barrier(quarter);
  }
  uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
  if (row < g.m) {
    epilogue16<FLAGS, CW>(v0, ch, c + cgroup * 32, g, row, mycol, &r0, &o0);
    epilogue16<FLAGS, CW>(v1, ch, c + cgroup * 32 + 16, g, row, mycol + 16, &r1, &o1);
  }
  *reinterpret_cast<uint4*>(mine) = o0;
  *reinterpret_cast<uint4*>(mine + 16) = o1;
  quarter_barrier(quarter);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tq + 128 * i, rr = idx >> 3, ck = idx & 7;
    if (row0 + rr < g.m && col0 + ck * 16 < g.n)
      *reinterpret_cast<uint4*>(g.out + (int64_t)(row0 + rr) * g.ld_out + col0 + ck * 16) =
          *reinterpret_cast<const uint4*>(stage + rr * kStagePitch + ck * 16);
  }
  // no second barrier: the next block of this quarter writes the OTHER tile, and a warp reaches the barrier of that
  // block only after it has finished reading this one, so the tile is free again when its turn comes (block + 2)
}

// Coalesced staging needs 16-byte aligned rows and whole 16-column chunks.  It pays off for the light
// (folded power-of-two) epilogue, whose stores were the busiest unit; the GELU and residual epilogues are
// instruction-issue bound and the extra barriers cost more than the stores (measured), so they store directly.
template <uint32_t FLAGS>
__device__ __forceinline__ bool can_stage(const GemmArgs& g) {
  if (FLAGS & EPI_PLAIN) return !(FLAGS & (EPI_GELU | EPI_RESIDUAL));
  return !(FLAGS & (EPI_GELU | EPI_RESIDUAL)) && g.raw_acc == nullptr && (g.ld_out & 15) == 0 && (g.n & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(g.out) & 15) == 0;
}

// One 16-column chunk of residual codes of one row, issued early.
template <uint32_t FLAGS>
__device__ __forceinline__ bool prefetch_residual16(const GemmArgs& g, int row, int col, uint4& r) {
  if ((FLAGS & EPI_PLAIN) && (FLAGS & EPI_RESIDUAL)) {
    if (row >= g.m || col >= g.n) return false;   // n % 16 == 0: a chunk lies entirely inside or outside the row
    r = ld_act_v4(g.epi.residual + (int64_t)row * g.ld_out + col);
    return true;
  }
  if (!(FLAGS & EPI_RESIDUAL) || row >= g.m || col + 16 > g.n || (g.ld_out & 15) != 0 || (col & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(g.epi.residual) & 15) != 0)
    return false;
  r = ld_act_v4(g.epi.residual + (int64_t)row * g.ld_out + col);
  return true;
}

// Early issue of the residual loads of one thread's 32 columns (two 16-byte words), so that their latency
// overlaps the wait for the accumulator.  Returns false when the vector path does not apply.
template <uint32_t FLAGS>
__device__ __forceinline__ bool prefetch_residual(const GemmArgs& g, int row, int col, uint4& r0, uint4& r1) {
  if (!(FLAGS & EPI_RESIDUAL) || row >= g.m || col + 32 > g.n || (g.ld_out & 15) != 0 || (col & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(g.epi.residual) & 15) != 0)
    return false;
  const uint4* rp = reinterpret_cast<const uint4*>(g.epi.residual + (int64_t)row * g.ld_out + col);
  r0 = ld_act_v4(rp);
  r1 = ld_act_v4(rp + 1);
  return true;
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <uint32_t FLAGS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // align by pointer arithmetic, not through an integer: the compiler then still knows these are shared-memory
  // addresses (LDS / STS instead of generic LD / ST plus window-base arithmetic on every epilogue constant load)
  GemmSmem& s = *reinterpret_cast<GemmSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (g.m + kBlockM - 1) / kBlockM;
  const int tiles_n = (g.n + kBlockN - 1) / kBlockN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(&s.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  // Tiles are walked n-fastest so that CTAs running at the same time share A tiles through L2.
  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();   // A is the previous kernel's output
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(&s.empty[stage], phase ^ 1);
          mbar_expect_tx(&s.full[stage], 2 * kTileBytes);
          tma_load_2d(s.a[stage], &tmap_a, &s.full[stage], kb * kBlockK, m0);
          tma_load_2d(s.b[stage], &tmap_b, &s.full[stage], kb * kBlockK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_i8(kBlockM, kBlockN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait_relaxed(&s.acc_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_sw128_kmajor(smem_u32(s.a[stage]));
          const uint64_t db = umma_desc_sw128_kmajor(smem_u32(s.b[stage]));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 bytes along K inside the swizzle row: +2 in the 16-byte address field
            tc_mma_i8(tmem_d, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc,
                      (uint32_t)((kb | k) != 0));
          }
          tc_commit(&s.empty[stage]);  // smem slot is free once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&s.acc_full[acc]);  // accumulator complete
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quarter = ew & 3;        // TMEM lane quarter this warp may access (= warp id % 4)
    const int cgroup = ew >> 2;        // 32-column group of the tile
    const int etid = threadIdx.x - 128;
    const bool staged = can_stage<FLAGS>(g);
    constexpr bool kHeavy = (FLAGS & (EPI_GELU | EPI_RESIDUAL)) != 0;
    // per-channel constants: resident for the whole kernel when n fits, else reloaded per tile into one of two slots
    const bool chan_resident = g.n <= kChanCols;
    if (chan_resident) {
      load_channels<FLAGS, kChanCols>(s.chan, g.epi, 0, g.n, etid);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    pdl_wait();   // residual codes are read, and out may still be read by the previous kernel
    uint32_t acc = 0, acc_phase = 0;
    uint32_t stage_turn = 0;   // which of the quarter's two staging tiles the next block uses
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
      int cbase = n0;   // column of this tile's constants inside s.chan
      if (!chan_resident) {
        // slot `acc` was last read two tiles ago; every epilogue warp has passed the barrier of the tile in between
        cbase = acc * kBlockN;
        load_channels<FLAGS, kChanCols>(s.chan, g.epi, n0, g.n, etid, cbase, kBlockN);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const int row = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + cgroup * 32;
      if (kHeavy && !staged) {
        // GELU / residual epilogues: one 16-column chunk at a time keeps the hot loop small (instruction cache);
        // the next chunk's residual codes are in flight while the current one is processed
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
        bool cur_ok = prefetch_residual16<FLAGS>(g, row, n0 + cgroup * 32, cur), nxt_ok = false;
        mbar_wait(&s.acc_full[acc], acc_phase);
        tc_fence_after_sync();
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          if (ch == 0) nxt_ok = prefetch_residual16<FLAGS>(g, row, n0 + cgroup * 32 + 16, nxt);
          uint32_t v[16];
          tmem_ld_32x16(taddr + ch * 16, v);
          tmem_ld_wait();
          if (ch == 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_empty[acc]);
          }
          if (row < g.m)
            epilogue16<FLAGS, kChanCols>(v, s.chan, cbase + cgroup * 32 + ch * 16, g, row, n0 + cgroup * 32 + ch * 16,
                                         cur_ok ? &cur : nullptr);
          cur = nxt;
          cur_ok = nxt_ok;
        }
      } else {
        mbar_wait(&s.acc_full[acc], acc_phase);
        tc_fence_after_sync();
        uint32_t v0[16], v1[16];
        tmem_ld_32x16(taddr, v0);
        tmem_ld_32x16(taddr + 16, v1);
        tmem_ld_wait();
        // the accumulator is in registers: release it to the MMA warp before the arithmetic
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.acc_empty[acc]);
        epilogue_block<FLAGS, kChanCols>(v0, v1, s.chan, cbase, g, m0 + quarter * 32, n0, quarter, cgroup, lane,
                                         s.stage[quarter][stage_turn & 1], staged);
        ++stage_turn;
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---- weight-stationary variant for k <= 384 ---------------------------------------------------------------
// With D = 384 three of the four GEMMs of a block have a K of only three k-blocks, and streaming both
// operands per 128x128 tile re-reads them from L2 so often that L2->SM bandwidth, not the tensor pipe,
// sets the pace (ncu: 340 MB of operand traffic for the 19 MB qkv input).  Here a CTA owns a slab of up to
// 256 output columns for its whole lifetime: the slab's weights ([256, k] <= 96 KiB) are loaded once and
// stay in shared memory, only A tiles stream through a 6-stage ring, each k-block feeds ONE
// M128 x N256 x K32 MMA chain into a 256-column accumulator, and two accumulators ping-pong so the 16
// epilogue warps drain one while the tensor pipe fills the other.  Per-channel constants are staged once.
constexpr int kBsMaxKb = 3;        // k <= 384
constexpr int kBsStagesA = 5;   // A ring (the double-buffered epilogue staging tiles took the sixth stage's shared memory)
constexpr int kBsSlabCols = 256;
constexpr int kBsMaxSlabs = 16;

struct BsSlab {
  int n0, nw;            // first column, width (<= 256)
  int cta_begin, cta_count;
};
struct BsPlan {
  int nslabs;
  BsSlab slab[kBsMaxSlabs];
};

struct BsSmem {
  alignas(1024) uint8_t b[kBsMaxKb][2][kTileBytes];   // [k-block][128-column half] = rows of the slab
  alignas(1024) uint8_t a[kBsStagesA][kTileBytes];
  alignas(16) float chan[CH_FIELDS][kBsSlabCols];
  alignas(16) uint8_t stage[4][2][32 * 144];   // per TMEM lane quarter, double-buffered (see epilogue_block)
  alignas(8) uint64_t full[kBsStagesA];
  uint64_t empty[kBsStagesA];
  uint64_t b_full;
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

template <uint32_t FLAGS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_bs_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmArgs g, const BsPlan plan) {
  extern __shared__ uint8_t smem_raw[];
  BsSmem& s = *reinterpret_cast<BsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (g.m + kBlockM - 1) / kBlockM;
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;

  // which slab does this CTA own, and which of the slab's CTAs is it
  int n0 = 0, nw = 0, local = 0, cnt = 0;
  for (int i = 0; i < plan.nslabs; ++i) {
    const BsSlab sl = plan.slab[i];
    if ((int)blockIdx.x >= sl.cta_begin && (int)blockIdx.x < sl.cta_begin + sl.cta_count) {
      n0 = sl.n0; nw = sl.nw; local = (int)blockIdx.x - sl.cta_begin; cnt = sl.cta_count;
    }
  }
  const int nsub = (nw + kBlockN - 1) / kBlockN;   // 128-column halves in use (0 for an idle CTA)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kBsStagesA; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (nsub > 0) {
    if (warp == 0) {
      if (elect_one()) {
        // the slab's weights, once
        mbar_expect_tx(&s.b_full, (uint32_t)(num_kb * nsub * kTileBytes));
        for (int kb = 0; kb < num_kb; ++kb)
          for (int h = 0; h < nsub; ++h) tma_load_2d(s.b[kb][h], &tmap_b, &s.b_full, kb * kBlockK, n0 + h * kBlockN);
        pdl_wait();   // the weights above are static; A is the previous kernel's output
        uint32_t stage = 0, phase = 0;
        for (int tile = local; tile < tiles_m; tile += cnt) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_relaxed(&s.empty[stage], phase ^ 1);
            mbar_expect_tx(&s.full[stage], kTileBytes);
            tma_load_2d(s.a[stage], &tmap_a, &s.full[stage], kb * kBlockK, tile * kBlockM);
            if (++stage == kBsStagesA) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_i8(kBlockM, nsub * kBlockN);
        mbar_wait(&s.b_full, 0);
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = local; tile < tiles_m; tile += cnt, ++it) {
          const uint32_t p = it & 1;
          mbar_wait_relaxed(&s.acc_empty[p], ((it >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + p * kBsSlabCols;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&s.full[stage], phase);
            tc_fence_after_sync();
            const uint64_t da = umma_desc_sw128_kmajor(smem_u32(s.a[stage]));
            const uint64_t db = umma_desc_sw128_kmajor(smem_u32(s.b[kb][0]));   // 256 rows: both halves
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tc_mma_i8(tmem_d, da + (uint64_t)(k * (kUmmaK >> 4)), db + (uint64_t)(k * (kUmmaK >> 4)), idesc,
                        (uint32_t)((kb | k) != 0));
            tc_commit(&s.empty[stage]);
            if (++stage == kBsStagesA) { stage = 0; phase ^= 1; }
          }
          tc_commit(&s.acc_full[p]);
        }
      }
    } else if (warp >= 4) {
      const int ew = warp - 4;
      const int quarter = ew & 3, cgroup = ew >> 2;
      load_channels<FLAGS, kBsSlabCols>(s.chan, g.epi, n0, g.n, (int)threadIdx.x - 128);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      pdl_wait();   // residual codes are read, and out may still be read by the previous kernel
      uint32_t it = 0;
      uint32_t stage_turn = 0;   // which of the quarter's two staging tiles the next block uses
      const bool staged = can_stage<FLAGS>(g);
      constexpr bool kHeavy = (FLAGS & (EPI_GELU | EPI_RESIDUAL)) != 0;
      for (int tile = local; tile < tiles_m; tile += cnt, ++it) {
        const uint32_t p = it & 1;
        const int row0 = tile * kBlockM + quarter * 32;
        const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16) + p * kBsSlabCols;
        if (kHeavy && !staged) {
          // 16-column chunks, rolled: small hot loop, next chunk's residual codes prefetched
          const int nchunks = nsub * 2;
          uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
          bool cur_ok = prefetch_residual16<FLAGS>(g, row0 + lane, n0 + cgroup * 32, cur), nxt_ok = false;
          mbar_wait(&s.acc_full[p], (it >> 1) & 1);
          tc_fence_after_sync();
#pragma unroll 1
          for (int ch = 0; ch < nchunks; ++ch) {
            const int c = (ch >> 1) * kBlockN + cgroup * 32 + (ch & 1) * 16;   // column inside the slab
            if (ch + 1 < nchunks) {
              const int cn = ((ch + 1) >> 1) * kBlockN + cgroup * 32 + ((ch + 1) & 1) * 16;
              nxt_ok = prefetch_residual16<FLAGS>(g, row0 + lane, n0 + cn, nxt);
            }
            uint32_t v[16];
            tmem_ld_32x16(tq + c, v);
            tmem_ld_wait();
            if (ch == nchunks - 1) {   // everything is in registers: hand the accumulator back
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s.acc_empty[p]);
            }
            if (row0 + lane < g.m)
              epilogue16<FLAGS, kBsSlabCols>(v, s.chan, c, g, row0 + lane, n0 + c, cur_ok ? &cur : nullptr);
            cur = nxt;
            cur_ok = nxt_ok;
          }
        } else {
          mbar_wait(&s.acc_full[p], (it >> 1) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h >= nsub) break;
            const int c = h * kBlockN + cgroup * 32;
            uint32_t v0[16], v1[16];
            tmem_ld_32x16(tq + c, v0);
            tmem_ld_32x16(tq + c + 16, v1);
            tmem_ld_wait();
            if (h == nsub - 1) {
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s.acc_empty[p]);
            }
            epilogue_block<FLAGS, kBsSlabCols>(v0, v1, s.chan, h * kBlockN, g, row0, n0 + h * kBlockN, quarter, cgroup,
                                               lane, s.stage[quarter][stage_turn & 1], staged);
            ++stage_turn;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---- exhaustive check of gelu_code_fast2 (test hook) ---------------------------------------------------------------
// All 2^32 fp32 bit patterns: counts[0] = accepted elements whose int8 code differs from the reference expression,
// counts[1] = rejected (guard) among |y| < 8, counts[2] = finite inputs with |y| < 8.
__global__ void gelu_fast_sweep_kernel(float rso, unsigned long long* __restrict__ counts) {
  unsigned long long bad = 0, rejected = 0, total = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  // the epilogue evaluates pairs (gelu_code_fast2): sweep consecutive bit patterns two at a time
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 31); i += stride) {
    const float yy[2] = {__uint_as_float((uint32_t)(2 * i)), __uint_as_float((uint32_t)(2 * i + 1))};
    const float hh[2] = {0.5f * rso, 0.5f * rso};
    float lo[2], tq[2];
    gelu_code_fast2(yy, hh, lo, tq);
    // accepted as the epilogue accepts: the low and the high end of the error interval pack to the same codes (the
    // epilogue compares groups of four; a pair here is at least as strict per element)
    const bool ok = pack_sat4(lo[0], lo[1], 0.f, 0.f) == pack_sat4(tq[0], tq[1], 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // |y| < 2^64: far beyond any |acc| < 2^31 times a scale below 1 plus a bias; above it 0.5 y rso can overflow to
      // infinity, whose error interval is not a number
      if (!isfinite(yy[k]) || fabsf(yy[k]) >= 1.8446744e19f) continue;
      const float ref = fmul(gelu_erf(yy[k]), rso);
      const int cf = max(-128, min(127, __float2int_rn(fminf(fmaxf(tq[k], -1e6f), 1e6f))));
      const int cr = max(-128, min(127, __float2int_rn(fminf(fmaxf(ref, -1e6f), 1e6f))));
      if (ok && cf != cr) ++bad;
      if (fabsf(yy[k]) < 8.f) {
        ++total;
        if (!ok) ++rejected;
      }
    }
  }
  atomicAdd(&counts[0], bad);
  atomicAdd(&counts[1], rejected);
  atomicAdd(&counts[2], total);
}

// ---- CUDA-core cross-check (dp4a) -----------------------------------------------------------------------
template <uint32_t FLAGS>
__global__ void gemm_i8_simt_kernel(const int8_t* __restrict__ a, int64_t lda, const int8_t* __restrict__ w,
                                    const GemmArgs g) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= g.n || row >= g.m) return;
  const int* ap = reinterpret_cast<const int*>(a + (int64_t)row * lda);
  const int* wp = reinterpret_cast<const int*>(w + (int64_t)col * g.k);
  int acc = 0;
  for (int kk = 0; kk < g.k / 4; ++kk) acc = __dp4a(ap[kk], wp[kk], acc);
  const p2v_epilogue& e = g.epi;
  EpiChannel c = {e.acc_scale[col], e.bias ? e.bias[col] : 0.f, e.out_scale[col],
                  e.out_rscale ? e.out_rscale[col] : 0.f, 0.f, 1.f};
  int q = epilogue_code<FLAGS>(acc, c, e.out_zp);
  const int64_t off = (int64_t)row * g.ld_out + col;
  if (e.out_f32 != nullptr) e.out_f32[(int64_t)row * g.n + col] = fmul(fsub((float)q, e.out_zp), c.out_scale);
  if (FLAGS & EPI_RESIDUAL) {
    c.res_scale = e.res_scale[col];
    c.out2_scale = e.out2_scale[col];
    if (e.aux_codes != nullptr) e.aux_codes[off] = (int8_t)q;
    q = residual_code(q, (int)e.residual[off], c);
  }
  g.out[off] = (int8_t)q;
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// K-major int8 matrix [rows, k] with `ld` bytes between rows -> box {128 B of K, 128 rows}, 128B swizzle.
int make_tmap_kmajor(CUtensorMap* map, const void* ptr, int64_t rows, int64_t k, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return P2V_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & 15) != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row stride (ptr=%p ld=%lld)", ptr,
              (long long)ld);
    return P2V_ERR_INVALID;
  }
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld k=%lld ld=%lld)", (int)r, (long long)rows,
              (long long)k, (long long)ld);
    return P2V_ERR_CUDA;
  }
  return P2V_OK;
}

constexpr int kGemmSmemBytes = (int)sizeof(GemmSmem) + 1024;
constexpr int kBsSmemBytes = (int)sizeof(BsSmem) + 1024;

template <uint32_t FLAGS>
static int configure_one() {
  P2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_tc_kernel<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kGemmSmemBytes));
  P2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_bs_kernel<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kBsSmemBytes));
  return P2V_OK;
}

static int g_gemm_mode = 0;   // 0 = auto, 1 = streaming kernel only, 2 = weight-stationary whenever legal

// Split n into slabs of <= 256 columns and hand each slab CTAs in proportion to its width.
static bool make_bs_plan(int n, int k, int grid, BsPlan* plan) {
  if (k > kBsMaxKb * kBlockK) return false;
  const int nslabs = (n + kBsSlabCols - 1) / kBsSlabCols;
  if (nslabs > kBsMaxSlabs || nslabs > grid) return false;
  int units_total = 0, units[kBsMaxSlabs];
  for (int i = 0; i < nslabs; ++i) {
    const int nw = (n - i * kBsSlabCols) < kBsSlabCols ? (n - i * kBsSlabCols) : kBsSlabCols;
    plan->slab[i].n0 = i * kBsSlabCols;
    plan->slab[i].nw = nw;
    units[i] = (nw + kBlockN - 1) / kBlockN;
    units_total += units[i];
  }
  int given = 0;
  for (int i = 0; i < nslabs; ++i) {
    int c = grid * units[i] / units_total;
    plan->slab[i].cta_count = c < 1 ? 1 : c;
    given += plan->slab[i].cta_count;
  }
  for (int i = 0; given < grid; i = (i + 1) % nslabs) {   // leftovers to the widest slabs first
    if (units[i] == 2 || nslabs == 1 || i >= nslabs - 1) { plan->slab[i].cta_count++; given++; }
  }
  if (given > grid) return false;
  int begin = 0;
  for (int i = 0; i < nslabs; ++i) {
    plan->slab[i].cta_begin = begin;
    begin += plan->slab[i].cta_count;
  }
  plan->nslabs = nslabs;
  return true;
}

// Opt every instantiation into > 48 KiB of dynamic shared memory.  Done once per process, outside any
// stream capture.
int gemm_configure() {
  static unsigned long long done = 0;   // one bit per device
  int dev = 0;
  if (needs_configure(done, &dev)) {
    int rc = configure_one<0>();
    if (!rc) rc = configure_one<1>();
    if (!rc) rc = configure_one<2>();
    if (!rc) rc = configure_one<3>();
    if (!rc) rc = configure_one<4>();
    if (!rc) rc = configure_one<5>();
    if (!rc) rc = configure_one<6>();
    if (!rc) rc = configure_one<7>();
    if (!rc) rc = configure_one<EPI_PLAIN | 2>();
    if (!rc) rc = configure_one<EPI_PLAIN | 4>();
    if (!rc) rc = configure_one<EPI_PLAIN | 5>();
    if (!rc) rc = configure_one<EPI_PLAIN | 6>();
    if (rc) return rc;
    mark_configured(done, dev);
  }
  return P2V_OK;
}

template <uint32_t FLAGS>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, cudaStream_t st) {
  const int smem = kGemmSmemBytes;
  int rc = gemm_configure();
  if (rc) return rc;
  const int tiles_m = (g.m + kBlockM - 1) / kBlockM;
  const int tiles = tiles_m * ((g.n + kBlockN - 1) / kBlockN);
  BsPlan plan;
  // weight-stationary pays off once every CTA re-uses its slab for several row tiles
  if (g_gemm_mode != 1 && (g_gemm_mode == 2 || tiles >= 2 * kNumSMs) && make_bs_plan(g.n, g.k, kNumSMs, &plan)) {
    P2V_CHECK_CUDA(launch_pdl(1, gemm_i8_bs_kernel<FLAGS>, dim3(kNumSMs), dim3(kGemmThreads), kBsSmemBytes, st, ta, tb, g, plan));
    return P2V_OK;
  }
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  P2V_CHECK_CUDA(launch_pdl(1, gemm_i8_tc_kernel<FLAGS>, dim3(grid), dim3(kGemmThreads), smem, st, ta, tb, g));
  return P2V_OK;
}

static bool plain_launch(const GemmArgs& g) {
  const p2v_epilogue& e = g.epi;
  return g.raw_acc == nullptr && e.aux_codes == nullptr && e.out_f32 == nullptr && (g.n & 15) == 0 &&
         (g.ld_out & 15) == 0 && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 &&
         (!(e.flags & P2V_EPI_RESIDUAL) || (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0);
}

static int dispatch_tc(uint32_t flags, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g,
                       cudaStream_t st) {
  if (plain_launch(g)) {   // the shapes of the fused forward: kernels without dump / ragged-edge handling
    switch (flags & 7u) {
      case 2: return launch_tc<EPI_PLAIN | 2>(ta, tb, g, st);
      case 4: return launch_tc<EPI_PLAIN | 4>(ta, tb, g, st);
      case 5: return launch_tc<EPI_PLAIN | 5>(ta, tb, g, st);
      case 6: return launch_tc<EPI_PLAIN | 6>(ta, tb, g, st);
      default: break;
    }
  }
  switch (flags & 7u) {
    case 0: return launch_tc<0>(ta, tb, g, st);
    case 1: return launch_tc<1>(ta, tb, g, st);
    case 2: return launch_tc<2>(ta, tb, g, st);
    case 3: return launch_tc<3>(ta, tb, g, st);
    case 4: return launch_tc<4>(ta, tb, g, st);
    case 5: return launch_tc<5>(ta, tb, g, st);
    case 6: return launch_tc<6>(ta, tb, g, st);
    default: return launch_tc<7>(ta, tb, g, st);
  }
}

static int check_gemm_args(const int8_t* a, int64_t lda, const int8_t* w, const void* out, int m, int n, int k,
                           const p2v_epilogue* epi, bool need_epi) {
  P2V_REQUIRE(a && w && out, "p2v_gemm_i8: null operand");
  P2V_REQUIRE(m > 0 && n > 0 && k > 0, "p2v_gemm_i8: empty problem m=%d n=%d k=%d", m, n, k);
  P2V_REQUIRE((k & 15) == 0, "p2v_gemm_i8: k=%d must be a multiple of 16", k);
  P2V_REQUIRE(lda >= k, "p2v_gemm_i8: lda=%lld < k=%d", (long long)lda, k);
  if (need_epi) {
    P2V_REQUIRE(epi && epi->acc_scale && epi->out_scale, "p2v_gemm_i8: epilogue needs acc_scale and out_scale");
    P2V_REQUIRE(!(epi->flags & P2V_EPI_OUT_POT) || epi->out_rscale, "p2v_gemm_i8: OUT_POT needs out_rscale");
    P2V_REQUIRE(!(epi->flags & P2V_EPI_RESIDUAL) || (epi->residual && epi->res_scale && epi->out2_scale),
                "p2v_gemm_i8: RESIDUAL needs residual, res_scale and out2_scale");
    P2V_REQUIRE(!(epi->flags & P2V_EPI_OUT_F32) || epi->out_f32, "p2v_gemm_i8: OUT_F32 needs out_f32");
  }
  return P2V_OK;
}

int gemm_i8_tc(const CUtensorMap& ta, const CUtensorMap& tb, int8_t* out, int64_t ld_out, int m, int n, int k,
               const p2v_epilogue& epi, cudaStream_t st) {
  GemmArgs g = {out, ld_out, m, n, k, epi, nullptr};
  if (!(epi.flags & P2V_EPI_OUT_F32)) g.epi.out_f32 = nullptr;
  return dispatch_tc(epi.flags, ta, tb, g, st);
}

}  // namespace p2v

using namespace p2v;

extern "C" int p2v_gemm_i8(const int8_t* a, int64_t lda, const int8_t* w, int8_t* out, int64_t ld_out, int m,
                           int n, int k, const p2v_epilogue* epi, void* stream) {
  int rc = check_gemm_args(a, lda, w, out, m, n, k, epi, true);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_kmajor(&ta, a, m, k, lda))) return rc;
  if ((rc = make_tmap_kmajor(&tb, w, n, k, k))) return rc;
  return gemm_i8_tc(ta, tb, out, ld_out, m, n, k, *epi, (cudaStream_t)stream);
}

extern "C" int p2v_test_gelu_fast(float out_rscale, unsigned long long* counts, void* stream) {
  P2V_REQUIRE(counts != nullptr && out_rscale > 0.f, "p2v_test_gelu_fast: bad arguments");
  gelu_fast_sweep_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(out_rscale, counts);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

extern "C" int p2v_gemm_set_mode(int mode) {
  P2V_REQUIRE(mode >= 0 && mode <= 2, "p2v_gemm_set_mode: mode must be 0 (auto), 1 (streaming) or 2 (weight-stationary)");
  g_gemm_mode = mode;
  return P2V_OK;
}

extern "C" int p2v_gemm_i8_acc(const int8_t* a, int64_t lda, const int8_t* w, int32_t* acc, int m, int n, int k,
                               void* stream) {
  int rc = check_gemm_args(a, lda, w, acc, m, n, k, nullptr, false);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_kmajor(&ta, a, m, k, lda))) return rc;
  if ((rc = make_tmap_kmajor(&tb, w, n, k, k))) return rc;
  GemmArgs g = {};
  g.m = m; g.n = n; g.k = k; g.ld_out = n; g.raw_acc = acc;
  return dispatch_tc(0, ta, tb, g, (cudaStream_t)stream);
}

extern "C" int p2v_gemm_i8_simt(const int8_t* a, int64_t lda, const int8_t* w, int8_t* out, int64_t ld_out,
                                int m, int n, int k, const p2v_epilogue* epi, void* stream) {
  int rc = check_gemm_args(a, lda, w, out, m, n, k, epi, true);
  if (rc) return rc;
  P2V_REQUIRE((lda & 3) == 0, "p2v_gemm_i8_simt: lda must be a multiple of 4");
  GemmArgs g = {out, ld_out, m, n, k, *epi, nullptr};
  if (!(epi->flags & P2V_EPI_OUT_F32)) g.epi.out_f32 = nullptr;
  dim3 grid((n + 127) / 128, m), block(128);
  cudaStream_t st = (cudaStream_t)stream;
  switch (epi->flags & 7u) {
    case 0: gemm_i8_simt_kernel<0><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 1: gemm_i8_simt_kernel<1><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 2: gemm_i8_simt_kernel<2><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 3: gemm_i8_simt_kernel<3><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 4: gemm_i8_simt_kernel<4><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 5: gemm_i8_simt_kernel<5><<<grid, block, 0, st>>>(a, lda, w, g); break;
    case 6: gemm_i8_simt_kernel<6><<<grid, block, 0, st>>>(a, lda, w, g); break;
    default: gemm_i8_simt_kernel<7><<<grid, block, 0, st>>>(a, lda, w, g); break;
  }
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}
