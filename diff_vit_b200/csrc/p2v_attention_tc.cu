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
// Warp roles (10 warps): 0-3 softmax + epilogue of tile 0 (warp q owns TMEM lane quarter q = its 32 rows), 4-7 the
// same for tile 1, 8 TMA producer (Q, K, V of the next item into a 2-stage ring), 9 MMA issuer (one elected lane; it
// polls both tile pipelines, which run independently of each other).  With tcgen05.ld 32x32b a thread owns a whole
// score row: row maximum, exact row sum and the per-row constants need no shuffles.
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
constexpr int kTcSoftWarps = 8;
constexpr int kTcThreads = (kTcSoftWarps + 4) * 32;   // warps 8, 9 unused: the two control warps sit on the
                                                         // schedulers (warp id % 4 = 2, 3) whose softmax warps have the least to do
constexpr int kTcProducerWarp = 10, kTcMmaWarp = 11;
constexpr int kTcStages = 2;
constexpr int kTcTileCols = 240;       // TMEM columns per row tile
constexpr int kColPhi = 0, kColPlo1 = 56, kColOhi = 80, kColOlo = 144, kColPlo0 = 208;
constexpr uint32_t kMagic = 0x4B400000u;   // 1.5 * 2^23

struct TcSmem {
  alignas(1024) uint8_t q[kTcStages][2][128 * 64];
  alignas(1024) uint8_t k[kTcStages][kTcMaxN * 64];
  alignas(1024) uint8_t v[kTcStages][kTcVRows * 64];
  alignas(1024) uint8_t ostage[kTcSoftWarps][32 * 64];   // per warp: its 32 output rows, 64-byte swizzle, TMA-stored
  alignas(16) uint32_t tab_e[256 * 32];   // [255 - d][lane]: high word of (double)e(d)
  alignas(16) float2 tab_r[256 * 32];     // [255 - d][lane]: low / high bracket of 1 / (3 e(d))
  float lut[256];                         // e(d), exact path
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

// CN (8, 16 or 32) consecutive columns of this thread's row
template <int CN>
__device__ __forceinline__ void ld_chunk(uint32_t taddr, uint32_t (&v)[32]) {
  if constexpr (CN == 32) {
    tmem_ld_32x32(taddr, v);
  } else if constexpr (CN == 16) {
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

// pass 1 over one chunk: row maximum / minimum of the raw accumulators (monotone in the score code)
template <int CN, bool kMask>
__device__ __forceinline__ void minmax_chunk(const uint32_t (&v)[32], int cnt, int& mx, int& mn) {
#pragma unroll
  for (int j = 0; j < CN; ++j) {
    if (!kMask || j < cnt) {
      mx = max(mx, (int)v[j]);
      mn = min(mn, (int)v[j]);
    }
  }
}

// pass 2 over one chunk: exact sum of e(d) over its first `cnt` columns (all CN without kMask)
template <bool kClamp, int CN, bool kMask>
__device__ __forceinline__ void sum_chunk(const uint32_t (&v)[32], int cnt, const RowConst& rc, uint32_t ke, double (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < CN; j += 2) {
    const float2 f = score_pair<kClamp>(v[j], v[j + 1], rc);
    const uint32_t e0 = lds32((__float_as_uint(f.x) << 7) + ke);
    const uint32_t e1 = lds32((__float_as_uint(f.y) << 7) + ke);
    if (!kMask || j < cnt) acc[(j >> 1) & 3] += __hiloint2double((int)e0, 0);
    if (!kMask || j + 1 < cnt) acc[((j >> 1) + 2) & 3] += __hiloint2double((int)e1, 0);
  }
}

// pass 3 over one chunk: 16-bit probabilities 2^(15-k) of its first `cnt` columns into v (0 beyond, up to column 32),
// returns the OR of (low-bracket bits ^ high-bracket bits): a set exponent bit means some element sits next to a step
template <bool kClamp, bool kPeak, int CN, bool kMask>
__device__ __forceinline__ uint32_t prob_chunk(uint32_t (&v)[32], int cnt, const RowConst& rc, uint32_t kr, float fsum,
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
  for (int j = CN; j < 32; ++j) v[j] = 0u;
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
      mbar_init(&s.empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s.s_full[t], 1);
      mbar_init(&s.p_ready[t], 4);
      mbar_init(&s.o_full[t], 1);
      mbar_init(&s.s_free[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == kTcMmaWarp) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (warp < kTcSoftWarps) {
    // ---- tables (all eight softmax warps) and the initial accumulator bias of this warp's lanes ----------------------
    {
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
    const int t = warp >> 2, q = warp & 3;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + t * kTcTileCols;
#pragma unroll
    for (int c = 0; c < kTcMaxN; c += 16) tmem_fill_32x16(trow + c, kMagic);
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
  } else if (warp == kTcMmaWarp) {
    // ---- MMA issuer: two independent tile pipelines, polled -----------------------------------------------------------
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_i8x(128, (uint32_t)nmma, true, true, false);
      const uint32_t idesc_pv = umma_idesc_i8x(128, 64, false, true, true);
      int s_item[2] = {0, 0};    // next item whose S this tile needs
      int pv_item[2] = {0, 0};   // next item whose P V this tile needs
      int released = 0;          // items whose smem stage went back to the producer
      const int last_tile = ntiles - 1;
      while (pv_item[0] < my_items || (ntiles == 2 && pv_item[1] < my_items)) {
        bool progressed = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t > last_tile) continue;
          const uint32_t tile = tmem_base + t * kTcTileCols;
          // S of item s_item[t]: its operands have landed, and the tile's TMEM region was handed back
          if (s_item[t] < my_items && s_item[t] == pv_item[t]) {
            const int i = s_item[t], st = i & 1;
            if (mbar_test(&s.full[st], (i >> 1) & 1) && (i == 0 || mbar_test(&s.s_free[t], (i - 1) & 1))) {
              tc_fence_after_sync();
              const uint64_t dq = umma_desc_sw64(smem_u32(s.q[st][t])), dk = umma_desc_sw64(smem_u32(s.k[st]));
              tc_mma_i8(tile, dq, dk, idesc_s, 1u);                 // accumulates onto the 1.5 * 2^23 bias
              tc_mma_i8(tile, dq + 2, dk + 2, idesc_s, 1u);         // second half of the head dimension: +32 bytes
              tc_commit(&s.s_full[t]);
              ++s_item[t];
              progressed = true;
            }
          }
          // P V of item pv_item[t]: all four warps of the tile have written their probability planes
          if (pv_item[t] < s_item[t]) {
            const int i = pv_item[t], st = i & 1;
            if (mbar_test(&s.p_ready[t], i & 1)) {
              tc_fence_after_sync();
              const uint64_t dv = umma_desc_sw64(smem_u32(s.v[st]));
              for (int ks = 0; ks < nchunks; ++ks) {
                const uint64_t dvk = dv + (uint64_t)(ks * (2048 >> 4));
                tc_mma_i8_ts(tile + kColOhi, tile + kColPhi + 8 * ks, dvk, idesc_pv, (uint32_t)(ks != 0));
                tc_mma_i8_ts(tile + kColOlo, tile + (ks < 4 ? kColPlo0 + 8 * ks : kColPlo1 + 8 * (ks - 4)), dvk, idesc_pv,
                             (uint32_t)(ks != 0));
              }
              tc_commit(&s.o_full[t]);
              ++pv_item[t];
              progressed = true;
              // the stage is free once both tiles' P V of its item have been issued (the commit covers all earlier MMAs)
              const int done = ntiles == 2 ? min(pv_item[0], pv_item[1]) : pv_item[0];
              while (released < done) {
                tc_commit(&s.empty[released & 1]);
                ++released;
              }
            }
          }
        }
        if (!progressed) __nanosleep(64);
      }
    }
  } else if (warp < kTcSoftWarps) {
    // ---- softmax + epilogue warps -------------------------------------------------------------------------------------
    const int t = warp >> 2, q = warp & 3;
    if (t < ntiles) {
      const uint32_t tile = tmem_base + ((uint32_t)(q * 32) << 16) + t * kTcTileCols;
      const float sixth = __fmul_rn(0.16666667f, __uint_as_float((127u - 120u) << 23));
      const uint32_t tab_e = smem_u32(s.tab_e) + lane * 4, tab_r = smem_u32(s.tab_r) + lane * 8;
      const int sh = a.out_shift;
      const int half_m1 = (1 << (sh - 1)) - 1 + (a.out_zp << sh);   // RNE shift with the zero point folded in
      uint8_t* const ost = s.ostage[warp];
      const int nfull = n >> 5;                       // chunks of 32 valid columns
      const int last_cnt = n - 32 * nfull;            // valid columns of the ragged last chunk (0: none)
      const int last_cn = last_cnt <= 8 ? 8 : (last_cnt <= 16 ? 16 : 32);

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
            a.timeline[(i * 8 + warp) * 8 + phase] = clock64();
        };
        stamp(0);
        mbar_wait_parked(&s.s_full[t], i & 1);
        tc_fence_after_sync();
        stamp(1);
        if (warp_on) {
          uint32_t v[32];
          // ---- pass 1: row maximum / minimum of the raw accumulators ----
          int mx = (int)0x80000000, mn = 0x7fffffff;
          for (int c = 0; c < nfull; ++c) {
            ld_chunk<32>(tile + 32 * c, v);
            minmax_chunk<32, false>(v, 32, mx, mn);
          }
          if (last_cnt > 0) {
            if (last_cn == 8) { ld_chunk<8>(tile + 32 * nfull, v); minmax_chunk<8, true>(v, last_cnt, mx, mn); }
            else if (last_cn == 16) { ld_chunk<16>(tile + 32 * nfull, v); minmax_chunk<16, true>(v, last_cnt, mx, mn); }
            else { ld_chunk<32>(tile + 32 * nfull, v); minmax_chunk<32, true>(v, last_cnt, mx, mn); }
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
          for (int c = 0; c < nfull; ++c) {
            ld_chunk<32>(tile + 32 * c, v);
            if (clampw) sum_chunk<true, 32, false>(v, 32, rc, ke, acc);
            else sum_chunk<false, 32, false>(v, 32, rc, ke, acc);
          }
          if (last_cnt > 0) {
            if (last_cn == 8) { ld_chunk<8>(tile + 32 * nfull, v); sum_chunk<true, 8, true>(v, last_cnt, rc, ke, acc); }
            else if (last_cn == 16) { ld_chunk<16>(tile + 32 * nfull, v); sum_chunk<true, 16, true>(v, last_cnt, rc, ke, acc); }
            else { ld_chunk<32>(tile + 32 * nfull, v); sum_chunk<true, 32, true>(v, last_cnt, rc, ke, acc); }
          }
          stamp(3);
          const float fsum = __double2float_rn((acc[0] + acc[1]) + (acc[2] + acc[3]));   // exact integer -> RNE, as u64 -> f32
          // u = S / (3e) + 1/6 is evaluated scaled by 2^-120 (exact), which puts its exponent field into 5 .. 31:
          // a shift count, 2^(15-k) = 0x100000 >> field
          const float fsum_s = __fmul_rn(fsum, __uint_as_float((127u - 120u) << 23));
          // the maximum itself: exact code; a row whose maximum holds most of the mass leaves the regular step pattern
          const int k_top = softmax_log_code(fsum, s.lut[0], 16);
          const uint32_t p_top = k_top >= 16 ? 0u : (0x8000u >> k_top);
          const bool peakw = __any_sync(0xffffffffu, k_top < 2);

          // ---- pass 3: probabilities 2^(15-k) as two byte planes, written back to TMEM as the A operand of P V ----
          for (int c = 0; c < nchunks; ++c) {
            const bool ragged = c == nfull;
            const int cnt = ragged ? last_cnt : 32;
            uint32_t guard;
            if (!ragged) {
              ld_chunk<32>(tile + 32 * c, v);
              if (clampw) guard = peakw ? prob_chunk<true, true, 32, false>(v, 32, rc, kr, fsum_s, sixth, fmax_bits, p_top)
                                        : prob_chunk<true, false, 32, false>(v, 32, rc, kr, fsum_s, sixth, fmax_bits, p_top);
              else guard = peakw ? prob_chunk<false, true, 32, false>(v, 32, rc, kr, fsum_s, sixth, fmax_bits, p_top)
                                 : prob_chunk<false, false, 32, false>(v, 32, rc, kr, fsum_s, sixth, fmax_bits, p_top);
            } else if (last_cn == 8) {
              ld_chunk<8>(tile + 32 * c, v);
              guard = prob_chunk<true, true, 8, true>(v, cnt, rc, kr, fsum_s, sixth, fmax_bits, p_top);
            } else if (last_cn == 16) {
              ld_chunk<16>(tile + 32 * c, v);
              guard = prob_chunk<true, true, 16, true>(v, cnt, rc, kr, fsum_s, sixth, fmax_bits, p_top);
            } else {
              ld_chunk<32>(tile + 32 * c, v);
              guard = prob_chunk<true, true, 32, true>(v, cnt, rc, kr, fsum_s, sixth, fmax_bits, p_top);
            }
            const bool redo = __any_sync(0xffffffffu, (guard & 0x7f800000u) != 0u);
            if (redo || kDump) {
              // redo (rare): some element within 2^-20 of a step of the code function: the whole chunk again with the
              // exact IEEE-division formula.  The raw scores are still in TMEM (P is written behind the read position).
              uint32_t r2[32];
              ld_chunk<32>(tile + 32 * c, r2);
              int8_t* dsc = kDump ? a.dump_scores + ((int64_t)item * n + row) * n + 32 * c : nullptr;
              uint8_t* dsm = kDump ? a.dump_softmax + ((int64_t)item * n + row) * n + 32 * c : nullptr;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float f = fminf(fmaxf(__fmaf_rn(__uint_as_float(r2[j]), rc.mul, rc.c0), rc.flo), rc.fhi);
                const int g = (int)(__float_as_uint(f) - kMagic);
                if (redo) v[j] = j < cnt ? tc_exact_prob16(fsum, s.lut[cmaxb - g]) : 0u;
                if (kDump && valid && j < cnt) {
                  dsc[j] = (int8_t)(g - 128);
                  dsm[j] = (uint8_t)(v[j] ? __clz(v[j]) - 16 : 16);
                }
              }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              const uint32_t a01 = v[4 * w] | (v[4 * w + 1] << 16), a23 = v[4 * w + 2] | (v[4 * w + 3] << 16);
              lo[w] = __byte_perm(a01, a23, 0x6420);
              hi[w] = __byte_perm(a01, a23, 0x7531);
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
          if (lane == 0) tma_store_wait_read();      // the previous item's store has finished reading the staging tile
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t oh[32], ol[32];
            tmem_ld_32x32(tile + kColOhi + 32 * h, oh);
            tmem_ld_32x32(tile + kColOlo + 32 * h, ol);
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
              const int chunk = (2 * h + ck) ^ ((lane >> 1) & 3);
              *reinterpret_cast<uint4*>(ost + lane * 64 + chunk * 16) = make_uint4(w8[4 * ck], w8[4 * ck + 1], w8[4 * ck + 2], w8[4 * ck + 3]);
            }
          }
          fence_proxy_async_smem();
        }
        stamp(6);
        // hand the region back with the accumulator bias in place for the next item's S (every warp: with the rotation
        // of the second tile a quarter that idles now holds rows next time)
#pragma unroll
        for (int c = 0; c < kTcMaxN; c += 16) tmem_fill_32x16(tile + c, kMagic);
        tmem_ld_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        stamp(7);
        if (lane == 0) {
          mbar_arrive(&s.s_free[t]);
          if (warp_on) {
            tma_store_3d(&tm_out, ost, head * 64, row0, img);   // rows >= n are clipped by the tensor map
            tma_store_commit();
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kTcMmaWarp) tmem_dealloc<512>(tmem_base);
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

// Test hook: device buffer of 12 x 8 x 8 int64 that receives clock64 stamps of CTA 0's softmax warps
// ([item][warp][phase]: 0 before / 1 after the S wait, 2 / 3 / 4 after passes 1 / 2 / 3, 5 after the O wait, 6 after the
// epilogue arithmetic, 7 after the accumulator bias was restored); NULL switches it off.
extern "C" int p2v_attention_tc_set_timeline(long long* buf) {
  p2v::g_tc_timeline = buf;
  return P2V_OK;
}
