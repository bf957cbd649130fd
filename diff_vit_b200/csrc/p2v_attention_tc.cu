// p2v_attention_tc.cu - fused integer attention on the 5th-gen tensor cores: QK^T and PV as tcgen05.mma kind::i8
// with TMEM accumulators, operands moved by TMA, and a thread-per-row log-int-softmax between them.
//
//   S = Q K^T (int32, TMEM) -> qact_attn1 int8 codes -> log-int-softmax 4-bit log2 codes -> P V (TMEM) -> qact2 codes
//
// Replaces models/vit_fquant.py:308-326 and QIntSoftmax.forward (models/ptq/layers.py:323-376) for the power-of-two
// grids of the minmax observer (every BASELINE config but 3); other grids keep attention_int_kernel
// (p2v_attention.cu), and both produce the same codes (tests/test_gpu_kernels.py runs the two against the host
// arithmetic and against each other).
//
// One work item = one (image, head): Q [n, 64], K [n, 64], V [n, 64] int8, n <= 208.  A persistent CTA per SM walks
// its items.  Per item the 197 query rows form two M = 128 row tiles; each tile owns a 240-column TMEM region:
//
//   columns   0..207  S = Q K^T for this tile (M128 x N208 x K64: two kind::i8 MMAs), pre-biased (see below)
//   then, as the softmax consumes S left to right, the same columns are re-used:
//   columns   0.. 55  P high plane, 4 keys per 32-bit column (A operand of PV, read by the MMA straight from TMEM)
//   columns  56.. 79  P low plane, keys 128..223          columns 208..239  P low plane, keys 0..127
//   columns  80..143  O high = P_hi V    columns 144..207  O low = P_lo V   (M128 x N64 x K32 per 32 keys, V MN-major)
//
// Warp roles (19 warps): 0-7 softmax + epilogue of tile 0, 8-15 of tile 1, 16 / 17 MMA issuers of tile 0 / 1 (one elected
// lane each; the two tile pipelines run independently of each other), 18 TMA producer (Q, K, V of the next item into a
// 2-stage ring).  Control warps only ever sleep on mbarriers (try_wait with a suspend hint): no polling.  A softmax warp w owns TMEM lane quarter w % 4 of its tile (tcgen05.ld lets a warp touch no other lanes),
// i.e. 32 rows, one per thread (32x32b loads): row maximum, exact row sum and per-row constants need no shuffles.  Two
// warps share each 32 rows and walk alternate 32-key chunks of them (w / 4 even: chunks 0, 2, 4, 6; odd: 1, 3, 5),
// exchanging the row maximum / minimum and the partial row sums through shared memory: a lone warp spends most of its
// time on the latencies of TMEM loads, table lookups and the fp64 sum chain (timeline: 0.4 instructions per cycle),
// four resident warps per scheduler hide them.
//
// Arithmetic per score element (the kernel is bound by CUDA-core issue, not by the tensor pipe, so this is what counts):
//   * the TMEM accumulator starts from 0x4B400000 (tcgen05.st by the softmax warps, MMAs accumulate on top), so the
//     int32 re-read as fp32 IS 1.5 * 2^23 + acc and one FFMA with the power-of-two multiplier gives
//     1.5 * 2^23 + RNE(acc * mul) + zp + 128: the biased score code sits in the low mantissa bits, rounded half-even by
//     the FMA itself (two elements per fma.rn.f32x2);
//   * those bits, shifted, address per-lane replicated tables directly (one LEA): every lane reads its own bank, so the
//     256-entry lookups of e(d) (pass 2: exact row sum in fp64) and of 1 / (3 e(d)) (pass 3) are conflict-free;
//   * the log2 code is the exponent of fma(S, 1/(3e), 1/6) (see p2v_attention.cu); it is evaluated for a low and a
//     high bracket of 1/(3e) in one fma.rn.f32x2, and only if some element's two exponents differ (1e-6 of the
//     elements) the chunk is redone with the exact IEEE-division formula;  2^(15-k) = 0x100000 >> exponent;
//   * rows whose clamps to [-128, 127] can bite, and rows whose maximum holds more than 2/3 of the mass (the one
//     irregular step of the code function), take warp-uniform variants of the loops.
#include <cuda.h>
#include <math.h>

#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

constexpr int kTcMaxN = 208;           // keys per item: N of the S MMA (multiple of 16)
constexpr int kTcVRows = 224;          // V rows staged (multiple of 32: seven K = 32 steps of PV)
constexpr int kTcSoftWarps = 16;       // 2 row tiles x 4 lane quarters x 2 column halves
constexpr int kTcThreads = (kTcSoftWarps + 3) * 32;   // 608 threads: up to 104 registers each (the kernel needs 96)
constexpr int kTcMmaWarp0 = 16;        // MMA issuer of tile 0 (one elected lane), tile 1: the next warp
constexpr int kTcProducerWarp = 18;
// (setmaxnreg was tried to move registers from the control warps to the softmax warps: ptxas 12.9 then either fails to
// allocate or spills more than without it, so the softmax passes work on 16-column pieces that fit the plain budget.)
constexpr int kTcStages = 2;
constexpr int kTcTileCols = 240;       // TMEM columns per row tile
constexpr int kColPhi = 0, kColPlo1 = 56, kColOhi = 80, kColOlo = 144, kColPlo0 = 208;
constexpr uint32_t kMagic = 0x4B400000u;   // 1.5 * 2^23

struct TcSmem {
  alignas(1024) uint8_t q[kTcStages][2][128 * 64];
  alignas(1024) uint8_t k[kTcStages][kTcMaxN * 64];
  alignas(1024) uint8_t v[kTcStages][kTcVRows * 64];
  alignas(1024) uint8_t ostage[8][32 * 64];   // per warp pair: its 32 output rows, 64-byte swizzle, TMA-stored
  alignas(16) uint32_t tab_e[256 * 32];   // [255 - d][lane]: high word of (double)e(d)
  alignas(16) float2 tab_r[256 * 32];     // [255 - d][lane]: low / high bracket of 1 / (3 e(d))
  float lut[256];                         // e(d), exact path
  // exchange between the two warps that share 32 rows (each walks every other 32-key chunk of them)
  int2 x_minmax[8][2][32];
  double x_sum[8][2][32];
  alignas(8) uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t s_full[2], p_ready[2], o_full[2], s_free[2];
  uint32_t tmem_base;
};

struct TcArgs {
  int n, heads, items;
  float score_mul, c0;       // f = fma(t, score_mul, c0) = 1.5 * 2^23 + RNE(acc * mul) + zp + 128
  int out_shift, out_zp;
  const float* exp_lut;
  int8_t* dump_scores;
  uint8_t* dump_softmax;
  long long* timeline;       // test hook (p2v_attention_tc_set_timeline): per-phase clock64 stamps of CTA 0
};

// ---- PTX wrappers this kernel adds to p2v_common.cuh -------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 lanes x 8 columns = 128 rows x 32 bytes of K) is read from TMEM
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// the same value into 16 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_fill_32x16(uint32_t taddr, uint32_t c) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(c)
      : "memory");
}
__device__ __forceinline__ void tmem_fill_32x8(uint32_t taddr, uint32_t c) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(c) : "memory");
}
__device__ __forceinline__ void pair_barrier(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); }
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32(taddr, v); }

// Shared-memory matrix descriptors for tiles whose rows are 64 bytes with the 64-byte swizzle (what a TMA box of
// {64 B, rows} with CU_TENSOR_MAP_SWIZZLE_64B writes): 8-row groups are 512 B apart (SBO).
//   K-major  (Q, K: a row = one token's 64 head channels = the contraction dimension)
//   MN-major (V for P V: a row = one key = one step of the contraction; its 64 bytes are the N dimension)
// The layout-type / stride fields are the same; which dimension the 64 contiguous bytes are is told by the
// instruction descriptor's major bits.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;                     // leading byte offset (one swizzle atom in that direction: unused)
  d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                              // SWIZZLE_64B
  return d;
}
// kind::i8 instruction descriptor with explicit signedness and B major-ness
__host__ __device__ constexpr uint32_t umma_idesc_i8x(uint32_t m, uint32_t n, bool a_signed, bool b_signed, bool b_mn_major) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((b_mn_major ? 1u : 0u) << 16) |
         ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t s) {
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));   // PTX: shift amounts > 31 give 0
  return r;
}
__device__ __forceinline__ uint32_t pack4_s8(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}

__device__ __noinline__ uint32_t tc_exact_prob16(float fsum, float e) {
  const int k = softmax_log_code(fsum, e, 16);
  return k >= 16 ? 0u : (0x8000u >> k);
}

// ---- one row's three passes over its score row --------------------------------------------------------------------
struct RowConst {
  float mul, c0;          // score re-quantisation (see TcArgs)
  float flo, fhi;         // clamp bounds: biased codes 0 and 255 in the magic representation
};

template <bool kClamp>
__device__ __forceinline__ float2 score_pair(uint32_t r0, uint32_t r1, const RowConst& rc) {
  float2 f = ffma2(make_float2(__uint_as_float(r0), __uint_as_float(r1)), make_float2(rc.mul, rc.mul), make_float2(rc.c0, rc.c0));
  if (kClamp) {
    f.x = fminf(fmaxf(f.x, rc.flo), rc.fhi);
    f.y = fminf(fmaxf(f.y, rc.flo), rc.fhi);
  }
  return f;
}

// CN (8 or 16) consecutive columns of this thread's row.  (Pieces of 16 columns, not 32: the passes keep a piece and
// its per-element temporaries in registers, and sixteen softmax warps leave 120 registers per thread.)
template <int CN>
__device__ __forceinline__ void ld_piece(uint32_t taddr, uint32_t (&v)[16]) {
  if constexpr (CN == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
  }
  tmem_ld_wait();
}

// pass 1 over one piece: row maximum / minimum of the raw accumulators (monotone in the score code)
template <int CN, bool kMask>
__device__ __forceinline__ void minmax_piece(const uint32_t (&v)[16], int cnt, int& mx, int& mn) {
#pragma unroll
  for (int j = 0; j < CN; ++j) {
    if (!kMask || j < cnt) {
      mx = max(mx, (int)v[j]);
      mn = min(mn, (int)v[j]);
    }
  }
}

// pass 2 over one piece: exact sum of e(d) over its first `cnt` columns (all CN without kMask)
template <bool kClamp, int CN, bool kMask>
__device__ __forceinline__ void sum_piece(const uint32_t (&v)[16], int cnt, const RowConst& rc, uint32_t ke, double (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < CN; j += 2) {
    const float2 f = score_pair<kClamp>(v[j], v[j + 1], rc);
    const uint32_t e0 = lds32((__float_as_uint(f.x) << 7) + ke);
    const uint32_t e1 = lds32((__float_as_uint(f.y) << 7) + ke);
    if (!kMask || j < cnt) acc[(j >> 1) & 3] += __hiloint2double((int)e0, 0);
    if (!kMask || j + 1 < cnt) acc[((j >> 1) + 2) & 3] += __hiloint2double((int)e1, 0);
  }
}

// pass 3 over one piece: 16-bit probabilities 2^(15-k) of its first `cnt` columns into v (0 beyond, up to column 16),
// returns the OR of (low-bracket bits ^ high-bracket bits): a set exponent bit means some element sits next to a step
template <bool kClamp, bool kPeak, int CN, bool kMask>
__device__ __forceinline__ uint32_t prob_piece(uint32_t (&v)[16], int cnt, const RowConst& rc, uint32_t kr, float fsum,
                                               float sixth, uint32_t fmax_bits, uint32_t p_top) {
  uint32_t guard = 0;
#pragma unroll
  for (int j = 0; j < CN; j += 2) {
    const float2 f = score_pair<kClamp>(v[j], v[j + 1], rc);
    const float2 ra = lds64f((__float_as_uint(f.x) << 8) + kr);
    const float2 rb = lds64f((__float_as_uint(f.y) << 8) + kr);
    const float2 ua = ffma2(make_float2(fsum, fsum), ra, make_float2(sixth, sixth));
    const float2 ub = ffma2(make_float2(fsum, fsum), rb, make_float2(sixth, sixth));
    uint32_t ga = __float_as_uint(ua.x) ^ __float_as_uint(ua.y), gb = __float_as_uint(ub.x) ^ __float_as_uint(ub.y);
    uint32_t pa = shr_clamp(0x100000u, __float_as_uint(ua.x) >> 23);
    uint32_t pb = shr_clamp(0x100000u, __float_as_uint(ub.x) >> 23);
    if (kPeak) {   // the row maximum itself: its exact probability (the irregular first steps of the code function)
      pa = __float_as_uint(f.x) == fmax_bits ? p_top : pa;
      pb = __float_as_uint(f.y) == fmax_bits ? p_top : pb;
    }
    if (kMask) {
      if (j >= cnt) { pa = 0u; ga = 0u; }
      if (j + 1 >= cnt) { pb = 0u; gb = 0u; }
    }
    guard |= ga | gb;
    v[j] = pa;
    v[j + 1] = pb;
  }
#pragma unroll
  for (int j = CN; j < 16; ++j) v[j] = 0u;
  return guard;
}

template <bool kDump>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_q32,
                    const __grid_constant__ CUtensorMap tm_q16, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_out, const TcArgs a) {
  extern __shared__ uint8_t tc_smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n;
  const int ntiles = n > 128 ? 2 : 1;
  const int nmma = max(16, (n + 15) & ~15);          // N of the S MMA
  const int nchunks = (n + 31) >> 5;                 // 32-key chunks = K steps of P V
  const int my_items = ((int)a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == kTcProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_q128); tma_prefetch_desc(&tm_q32); tma_prefetch_desc(&tm_q16);
    tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_out);
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], (uint32_t)ntiles);   // one commit per tile pipeline
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s.s_full[t], 1);
      mbar_init(&s.p_ready[t], 8);
      mbar_init(&s.o_full[t], 1);
      mbar_init(&s.s_free[t], 8);
    }
    fence_mbar_init();
  }
  if (warp == kTcMmaWarp0) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (warp < kTcSoftWarps) {
    // ---- tables (warps 0-7, 32 entries each) and the initial accumulator bias of this warp's half of its lanes' columns --
    if (warp < 8) {
      const int d = warp * 32 + lane;                 // this lane evaluates entry d ...
      const float e = a.exp_lut[d];
      s.lut[d] = e;
      const float r3 = __fdiv_rn(1.0f, 3.0f * e);       // 3e is exact (<= 24 significant bits)
      const float rlo = __fmul_rn(r3, 1.0f - 9.5367431640625e-07f);   // 1 -+ 2^-20
      const float rhi = __fmul_rn(r3, 1.0f + 9.5367431640625e-07f);
      const uint32_t ehi = (uint32_t)__double2hiint((double)e);
      for (int j = 0; j < 32; ++j) {                  // ... and the warp writes each of its entries once per lane replica
        const int dj = warp * 32 + j;
        const uint32_t ej = __shfl_sync(0xffffffffu, ehi, j);
        const float lj = __shfl_sync(0xffffffffu, rlo, j), hj = __shfl_sync(0xffffffffu, rhi, j);
        s.tab_e[(255 - dj) * 32 + lane] = ej;
        s.tab_r[(255 - dj) * 32 + lane] = make_float2(lj, hj);
      }
    }
    const int t = warp >> 3, hf = (warp >> 2) & 1, q = warp & 3;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + t * kTcTileCols;
    for (int c = hf * (kTcMaxN / 2); c < (hf + 1) * (kTcMaxN / 2); c += 8) tmem_fill_32x8(trow + c, kMagic);
    tmem_ld_wait_st();
    tc_fence_before_sync();
  }
  __syncthreads();
  tc_fence_after_sync();

  if (warp == kTcProducerWarp) {
    // ---- TMA producer -----------------------------------------------------------------------------------------------
    if (elect_one()) {
      const uint32_t bytes = 8192u + (ntiles == 2 ? 5120u : 0u) + (uint32_t)(kTcMaxN * 64) + (uint32_t)(kTcVRows * 64);
      for (int i = 0; i < my_items; ++i) {
        const int item = blockIdx.x + i * gridDim.x, img = item / a.heads, head = item % a.heads;
        const int st = i & 1;
        mbar_wait_relaxed(&s.empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&s.full[st], bytes);
        tma_load_3d(s.q[st][0], &tm_q128, &s.full[st], head * 64, 0, img);
        tma_load_3d(s.k[st], &tm_k, &s.full[st], (a.heads + head) * 64, 0, img);
        if (ntiles == 2) {
          // rows 128.. of the second tile, 32 at a time, rotated over the four lane quarters from item to item so
          // that the short tile (69 rows for n = 197) loads every scheduler's warps equally in the long run
          const int rot = i & 3;
          tma_load_3d(s.q[st][1] + ((0 + rot) & 3) * 2048, &tm_q32, &s.full[st], head * 64, 128, img);
          tma_load_3d(s.q[st][1] + ((1 + rot) & 3) * 2048, &tm_q32, &s.full[st], head * 64, 160, img);
          tma_load_3d(s.q[st][1] + ((2 + rot) & 3) * 2048, &tm_q16, &s.full[st], head * 64, 192, img);
        }
        tma_load_3d(s.v[st], &tm_v, &s.full[st], (2 * a.heads + head) * 64, 0, img);
      }
    }
  } else if (warp == kTcMmaWarp0 || warp == kTcMmaWarp0 + 1) {
    // ---- MMA issuers: one warp (one elected lane) per tile pipeline ----------------------------------------------------
    const int t = warp - kTcMmaWarp0;
    if (t < ntiles && elect_one()) {
      const uint32_t idesc_s = umma_idesc_i8x(128, (uint32_t)nmma, true, true, false);
      const uint32_t idesc_pv = umma_idesc_i8x(128, 64, false, true, true);
      const uint32_t tile = tmem_base + t * kTcTileCols;
      for (int i = 0; i < my_items; ++i) {
        const int st = i & 1;
        // S of item i: its operands have landed, and the tile's TMEM region was handed back (bias restored)
        mbar_wait_parked(&s.full[st], (i >> 1) & 1);
        if (i > 0) mbar_wait_parked(&s.s_free[t], (i - 1) & 1);
        tc_fence_after_sync();
        const uint64_t dq = umma_desc_sw64(smem_u32(s.q[st][t])), dk = umma_desc_sw64(smem_u32(s.k[st]));
        tc_mma_i8(tile, dq, dk, idesc_s, 1u);                 // accumulates onto the 1.5 * 2^23 bias
        tc_mma_i8(tile, dq + 2, dk + 2, idesc_s, 1u);         // second half of the head dimension: +32 bytes
        tc_commit(&s.s_full[t]);
        // P V of item i: all eight warps of the tile have written their probability planes
        mbar_wait_parked(&s.p_ready[t], i & 1);
        tc_fence_after_sync();
        const uint64_t dv = umma_desc_sw64(smem_u32(s.v[st]));
        for (int ks = 0; ks < nchunks; ++ks) {
          const uint64_t dvk = dv + (uint64_t)(ks * (2048 >> 4));
          tc_mma_i8_ts(tile + kColOhi, tile + kColPhi + 8 * ks, dvk, idesc_pv, (uint32_t)(ks != 0));
          tc_mma_i8_ts(tile + kColOlo, tile + (ks < 4 ? kColPlo0 + 8 * ks : kColPlo1 + 8 * (ks - 4)), dvk, idesc_pv,
                       (uint32_t)(ks != 0));
        }
        tc_commit(&s.o_full[t]);
        tc_commit(&s.empty[st]);      // this pipeline is done with the stage once these MMAs retire
      }
    }
  } else if (warp < kTcSoftWarps) {
    // ---- softmax + epilogue warps -------------------------------------------------------------------------------------
    const int t = warp >> 3, hf = (warp >> 2) & 1, q = warp & 3;
    const int pair = t * 4 + q;                       // the two warps (hf = 0, 1) that share 32 rows
    if (t < ntiles) {
      const uint32_t tile = tmem_base + ((uint32_t)(q * 32) << 16) + t * kTcTileCols;
      const float sixth = __fmul_rn(0.16666667f, __uint_as_float((127u - 120u) << 23));
      const uint32_t tab_e = smem_u32(s.tab_e) + lane * 4, tab_r = smem_u32(s.tab_r) + lane * 8;
      const int sh = a.out_shift;
      const int half_m1 = (1 << (sh - 1)) - 1 + (a.out_zp << sh);   // RNE shift with the zero point folded in
      uint8_t* const ost = s.ostage[pair];

      for (int i = 0; i < my_items; ++i) {
        const int item = blockIdx.x + i * gridDim.x, img = item / a.heads, head = item % a.heads;
        const int seg = t == 0 ? q : ((q - (i & 3)) & 3);        // which 32 rows of the tile this warp's lanes hold
        const int row0 = t * 128 + seg * 32;
        const int row = row0 + lane;
        const bool warp_on = row0 < n && (t == 0 || seg < 3);
        const bool valid = warp_on && row < n;
        // lanes without a row (beyond n, or staging rows no load wrote) see a constant score row: their table
        // addresses stay inside the tables whatever the tensor core left in their TMEM lanes
        RowConst rc;
        rc.mul = valid ? a.score_mul : 0.f;
        rc.c0 = valid ? a.c0 : __uint_as_float(kMagic + 128u);
        rc.flo = __uint_as_float(kMagic);
        rc.fhi = __uint_as_float(kMagic + 255u);
        auto stamp = [&](int phase) {
          if (a.timeline != nullptr && blockIdx.x == 0 && lane == 0 && i < 12)
            a.timeline[(i * 16 + warp) * 8 + phase] = clock64();
        };
        stamp(0);
        mbar_wait_parked(&s.s_full[t], i & 1);
        tc_fence_after_sync();
        stamp(1);
        if (warp_on) {
          uint32_t v[16];
          // ---- pass 1: row maximum / minimum of the raw accumulators, this warp's chunks, then both halves ----
          int mx = (int)0x80000000, mn = 0x7fffffff;
          for (int c = hf; c < nchunks; c += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int col = 32 * c + 16 * h, cnt = n - col;
              if (cnt >= 16) { ld_piece<16>(tile + col, v); minmax_piece<16, false>(v, 16, mx, mn); }
              else if (cnt > 8) { ld_piece<16>(tile + col, v); minmax_piece<16, true>(v, cnt, mx, mn); }
              else if (cnt > 0) { ld_piece<8>(tile + col, v); minmax_piece<8, true>(v, cnt, mx, mn); }
            }
          }
          s.x_minmax[pair][hf][lane] = make_int2(mx, mn);
          pair_barrier(pair);
          {
            const int2 o = s.x_minmax[pair][hf ^ 1][lane];
            mx = max(mx, o.x);
            mn = min(mn, o.y);
          }
          stamp(2);
          const int gmax = (int)(__float_as_uint(__fmaf_rn(__int_as_float(mx), rc.mul, rc.c0)) - kMagic);
          const int gmin = (int)(__float_as_uint(__fmaf_rn(__int_as_float(mn), rc.mul, rc.c0)) - kMagic);
          const bool clampw = __any_sync(0xffffffffu, gmax > 255 || gmin < 0);
          const int cmaxb = min(max(gmax, 0), 255);                 // biased code of the row maximum
          const uint32_t fmax_bits = kMagic + (uint32_t)cmaxb;
          const uint32_t ke = tab_e + (uint32_t)(255 - cmaxb) * 128u - (kMagic << 7);
          const uint32_t kr = tab_r + (uint32_t)(255 - cmaxb) * 256u - (kMagic << 8);

          // ---- pass 2: exact row sum of the integer exp ----
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          for (int c = hf; c < nchunks; c += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int col = 32 * c + 16 * h, cnt = n - col;
              if (cnt >= 16) {
                ld_piece<16>(tile + col, v);
                if (clampw) sum_piece<true, 16, false>(v, 16, rc, ke, acc);
                else sum_piece<false, 16, false>(v, 16, rc, ke, acc);
              } else if (cnt > 8) { ld_piece<16>(tile + col, v); sum_piece<true, 16, true>(v, cnt, rc, ke, acc); }
              else if (cnt > 0) { ld_piece<8>(tile + col, v); sum_piece<true, 8, true>(v, cnt, rc, ke, acc); }
            }
          }
          const double part = (acc[0] + acc[1]) + (acc[2] + acc[3]);     // integers < 2^53: exact in any order
          s.x_sum[pair][hf][lane] = part;
          pair_barrier(pair);
          const float fsum = __double2float_rn(part + s.x_sum[pair][hf ^ 1][lane]);   // exact integer -> RNE, as u64 -> f32
          stamp(3);
          // u = S / (3e) + 1/6 is evaluated scaled by 2^-120 (exact), which puts its exponent field into 5 .. 31:
          // a shift count, 2^(15-k) = 0x100000 >> field
          const float fsum_s = __fmul_rn(fsum, __uint_as_float((127u - 120u) << 23));
          // the maximum itself: exact code; a row whose maximum holds most of the mass leaves the regular step pattern
          const int k_top = softmax_log_code(fsum, s.lut[0], 16);
          const uint32_t p_top = k_top >= 16 ? 0u : (0x8000u >> k_top);
          const bool peakw = __any_sync(0xffffffffu, k_top < 2);

          // ---- pass 3: probabilities 2^(15-k) as two byte planes, written back to TMEM as the A operand of P V ----
          // The probability planes of chunk c go to columns 8c.. (high) and 208 + 8c.. / 56 + 8(c - 4).. (low): score
          // columns of chunks 0, 1 and 2.  Chunks 0 / 2 are read by the even warp of the pair, chunk 1 by the odd one, so
          // a pair barrier after each warp's first and second chunk load keeps every write behind the reads of BOTH warps
          // (P(1), P(3) -> chunk 0 and P(4), P(6) -> chunk 1 after barrier one; P(5) -> chunk 2 after barrier two).
#pragma unroll 1
          for (int it = 0; it < 4; ++it) {
            const int c = hf + 2 * it;
            if (c >= nchunks) {
              if (it < 2) pair_barrier(pair);
              continue;
            }
            uint32_t hi[8], lo[8];                    // the chunk's two planes: words 0-3 keys 0-15, words 4-7 keys 16-31
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int col = 32 * c + 16 * h, cnt = min(n - col, 16);
              if (cnt >= 16 || cnt > 8) ld_piece<16>(tile + col, v);
              else if (cnt > 0) ld_piece<8>(tile + col, v);
              if (h == 1 && it < 2) pair_barrier(pair);      // both warps hold their whole chunk: P may now be written
              if (cnt <= 0) {
#pragma unroll
                for (int w = 0; w < 4; ++w) hi[4 * h + w] = lo[4 * h + w] = 0u;
                continue;
              }
              uint32_t guard;
              if (cnt >= 16) {
                if (clampw) guard = peakw ? prob_piece<true, true, 16, false>(v, 16, rc, kr, fsum_s, sixth, fmax_bits, p_top)
                                          : prob_piece<true, false, 16, false>(v, 16, rc, kr, fsum_s, sixth, fmax_bits, p_top);
                else guard = peakw ? prob_piece<false, true, 16, false>(v, 16, rc, kr, fsum_s, sixth, fmax_bits, p_top)
                                   : prob_piece<false, false, 16, false>(v, 16, rc, kr, fsum_s, sixth, fmax_bits, p_top);
              } else if (cnt > 8) {
                guard = prob_piece<true, true, 16, true>(v, cnt, rc, kr, fsum_s, sixth, fmax_bits, p_top);
              } else {
                guard = prob_piece<true, true, 8, true>(v, cnt, rc, kr, fsum_s, sixth, fmax_bits, p_top);
              }
              const bool redo = __any_sync(0xffffffffu, (guard & 0x7f800000u) != 0u);
              if (redo || kDump) {
                // redo (rare): some element within 2^-20 of a step of the code function: the whole piece again with the
                // exact IEEE-division formula.  The raw scores are still in TMEM (P is written behind the read position).
                uint32_t r2[16];
                ld_piece<16>(tile + col, r2);
                int8_t* dsc = kDump ? a.dump_scores + ((int64_t)item * n + row) * n + col : nullptr;
                uint8_t* dsm = kDump ? a.dump_softmax + ((int64_t)item * n + row) * n + col : nullptr;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float f = fminf(fmaxf(__fmaf_rn(__uint_as_float(r2[j]), rc.mul, rc.c0), rc.flo), rc.fhi);
                  const int g = (int)(__float_as_uint(f) - kMagic);
                  if (redo) v[j] = j < cnt ? tc_exact_prob16(fsum, s.lut[cmaxb - g]) : 0u;
                  if (kDump && valid && j < cnt) {
                    dsc[j] = (int8_t)(g - 128);
                    dsm[j] = (uint8_t)(v[j] ? __clz(v[j]) - 16 : 16);
                  }
                }
              }
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const uint32_t a01 = v[4 * w] | (v[4 * w + 1] << 16), a23 = v[4 * w + 2] | (v[4 * w + 3] << 16);
                lo[4 * h + w] = __byte_perm(a01, a23, 0x6420);
                hi[4 * h + w] = __byte_perm(a01, a23, 0x7531);
              }
            }
            tmem_st_32x8(tile + kColPhi + 8 * c, hi);
            tmem_st_32x8(tile + (c < 4 ? kColPlo0 + 8 * c : kColPlo1 + 8 * (c - 4)), lo);
          }
          tmem_ld_wait_st();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.p_ready[t]);
        stamp(4);

        // ---- epilogue: O = (O_hi << 8) + O_lo -> RNE shift to the qact2 grid -> int8, out through a TMA store ----
        mbar_wait_parked(&s.o_full[t], i & 1);
        tc_fence_after_sync();
        stamp(5);
        if (warp_on) {
          if (hf == 0 && lane == 0) tma_store_wait_read();   // the previous item's store has finished reading the staging tile
          pair_barrier(pair);
          {
            uint32_t oh[32], ol[32];                  // this warp's 32 of the 64 head channels
            tmem_ld_32x32(tile + kColOhi + 32 * hf, oh);
            tmem_ld_32x32(tile + kColOlo + 32 * hf, ol);
            tmem_ld_wait();
            uint32_t w8[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              int qv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int acc2 = ((int)oh[4 * w + e] << 8) + (int)ol[4 * w + e];
                qv[e] = (acc2 + half_m1 + ((acc2 >> sh) & 1)) >> sh;
              }
              w8[w] = pack4_s8(qv[0], qv[1], qv[2], qv[3]);
            }
            // row `lane` of the 32 x 64-byte staging tile, 16-byte chunks XOR-swizzled by (row >> 1) & 3 (SWIZZLE_64B)
#pragma unroll
            for (int ck = 0; ck < 2; ++ck) {
              const int chunk = (2 * hf + ck) ^ ((lane >> 1) & 3);
              *reinterpret_cast<uint4*>(ost + lane * 64 + chunk * 16) = make_uint4(w8[4 * ck], w8[4 * ck + 1], w8[4 * ck + 2], w8[4 * ck + 3]);
            }
          }
          fence_proxy_async_smem();
          pair_barrier(pair);
          if (hf == 0 && lane == 0) {
            tma_store_3d(&tm_out, ost, head * 64, row0, img);   // rows >= n are clipped by the tensor map
            tma_store_commit();
          }
        }
        stamp(6);
        // hand the region back with the accumulator bias in place for the next item's S (every warp: with the rotation
        // of the second tile a quarter that idles now holds rows next time)
        for (int c = hf * (kTcMaxN / 2); c < (hf + 1) * (kTcMaxN / 2); c += 8) tmem_fill_32x8(tile + c, kMagic);
        tmem_ld_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        stamp(7);
        if (lane == 0) mbar_arrive(&s.s_free[t]);
      }
      if (hf == 0 && lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kTcMmaWarp0) tmem_dealloc<512>(tmem_base);
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [images][tokens][row_bytes] int8 view, box {64 bytes, box_rows tokens, 1 image}, 64-byte swizzle; tokens >= n are
// out of bounds: zero-filled on loads, clipped on stores
static int make_tmap_tokens(CUtensorMap* map, const void* ptr, int b, int n, int64_t row_bytes, int box_rows) {
  EncodeTiledFn fn = tc_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return P2V_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)row_bytes, (cuuint64_t)n, (cuuint64_t)b};
  cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)row_bytes * (cuuint64_t)n};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (attention) failed with CUresult %d (b=%d n=%d row=%lld box=%d)", (int)r, b, n,
              (long long)row_bytes, box_rows);
    return P2V_ERR_CUDA;
  }
  return P2V_OK;
}

constexpr int kTcSmemBytes = (int)sizeof(TcSmem) + 1024;
static long long* g_tc_timeline = nullptr;

int attention_tc_configure() {
  P2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  P2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  return P2V_OK;
}

// Whether the tensor-core kernel covers this call (power-of-two score and output grids, no zero point on q/k/v,
// 16 softmax levels, table entries of <= 21 significant bits, n <= 208, tensor-map alignment).
bool attention_tc_applicable(const int8_t* qkv, const int8_t* out, int b, int n, int heads, const p2v_attention* p) {
  if (n > kTcMaxN || p->in_zp != 0.f || p->softmax_levels != 16) return false;
  if (p->lut_sig_bits <= 0 || p->lut_sig_bits > 21) return false;
  int ex = 0;
  if (!(p->score_mul > 0.f) || frexpf(p->score_mul, &ex) != 0.5f || ex > 1 || ex < -20) return false;
  if (p->score_zp != (float)(int)p->score_zp || fabsf(p->score_zp) > 128.f) return false;
  int ex2 = 0;
  if (!(p->out_mul > 0) || frexp(p->out_mul, &ex2) != 0.5 || ex2 > 0 || ex2 < -29) return false;
  if (p->out_zp != (float)(int)p->out_zp || fabsf(p->out_zp) > 128.f) return false;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  (void)b; (void)heads;
  return true;
}

int attention_tc_launch(const int8_t* qkv, int8_t* out, int b, int n, int heads, const p2v_attention* p, cudaStream_t st) {
  const int64_t row = (int64_t)3 * heads * 64, orow = (int64_t)heads * 64;
  CUtensorMap tq128, tq32, tq16, tk, tv, to;
  int rc;
  if ((rc = make_tmap_tokens(&tq128, qkv, b, n, row, 128))) return rc;
  if ((rc = make_tmap_tokens(&tq32, qkv, b, n, row, 32))) return rc;
  if ((rc = make_tmap_tokens(&tq16, qkv, b, n, row, 16))) return rc;
  if ((rc = make_tmap_tokens(&tk, qkv, b, n, row, kTcMaxN))) return rc;
  if ((rc = make_tmap_tokens(&tv, qkv, b, n, row, kTcVRows))) return rc;
  if ((rc = make_tmap_tokens(&to, out, b, n, orow, 32))) return rc;
  TcArgs a;
  a.n = n; a.heads = heads; a.items = b * heads;
  a.score_mul = p->score_mul;
  // 1.5 * 2^23 * (1 - mul) + zp + 128: exact in fp32 for mul = 2^-s, s <= 20 (checked by attention_tc_applicable)
  a.c0 = (float)(12582912.0 * (1.0 - (double)p->score_mul) + (double)p->score_zp + 128.0);
  int ex = 0;
  frexp(p->out_mul, &ex);
  a.out_shift = 1 - ex;
  a.out_zp = (int)p->out_zp;
  a.exp_lut = p->exp_lut;
  a.dump_scores = p->dump_scores;
  a.dump_softmax = p->dump_softmax;
  a.timeline = g_tc_timeline;
  const int grid = a.items < kNumSMs ? a.items : kNumSMs;
  if (p->dump_scores != nullptr)
    attention_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, st>>>(tq128, tq32, tq16, tk, tv, to, a);
  else
    attention_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, st>>>(tq128, tq32, tq16, tk, tv, to, a);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

}  // namespace p2v

// Test hook: device buffer of 12 x 16 x 8 int64 that receives clock64 stamps of CTA 0's softmax warps
// ([item][warp][phase]: 0 before / 1 after the S wait, 2 / 3 / 4 after passes 1 / 2 / 3, 5 after the O wait, 6 after the
// epilogue arithmetic, 7 after the accumulator bias was restored); NULL switches it off.
extern "C" int p2v_attention_tc_set_timeline(long long* buf) {
  p2v::g_tc_timeline = buf;
  return P2V_OK;
}
