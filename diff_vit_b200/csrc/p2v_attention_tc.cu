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

// Shared memory, by shared-window address (the lookup table must sit on a 64 KiB boundary, see below):
//   [A0, A0 + 59 KiB)      operand stage 0 (Q tile 0 / 1, K, V) and the eight 2 KiB output staging tiles; A0 = the
//                          1 KiB-aligned start of dynamic shared memory (must be <= 5 KiB)
//   [64 KiB, 128 KiB)      table, one 256-byte row per reversed distance 255 - d:
//                            bytes   0..127  high word of (double)e(d), replicated per lane (lane * 4)
//                            bytes 128..255  (low, high) bracket of 1 / (3 e(d)), 16 replicas ((lane & 15) * 8): a
//                                            64-bit load is served per half-warp, so lanes l and l + 16 never collide
//                          A lookup address is base | (index << 8): one PRMT that drops the index byte into byte 1 of
//                          the per-lane base - which is why the table needs the alignment.
//   [128 KiB, ...)         operand stage 1, e(d) for the exact path, the pair exchange slots, mbarriers
constexpr uint32_t kOffQ = 0, kOffK = 16384, kOffV = kOffK + kTcMaxN * 64, kStageBytes = kOffV + kTcVRows * 64;   // 44032
constexpr uint32_t kOffOstage = kStageBytes;                     // relative to A0
constexpr uint32_t kRegionABytes = kOffOstage + 8 * 2048;        // 60416
constexpr uint32_t kTabAddr = 65536, kRegionCAddr = 131072;
struct TcTail {                                                  // at kRegionCAddr + kStageBytes
  float lut[256];
  int2 x_minmax[8][2][32];     // exchange between the two warps that share 32 rows
  double x_sum[8][2][32];
  alignas(8) uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t s_full[2], o_full[2];
  uint32_t tmem_base;
};
constexpr uint32_t kTcSmemEnd = kRegionCAddr + kStageBytes + (uint32_t)sizeof(TcTail);

struct TcArgs {
  int n, heads, items;
  float score_mul, c0;       // f = fma(t, score_mul, c0) = acc * mul + zp + 128 for t = 1.5 * 2^23 + acc (exact)
  int out_shift, out_zp;
  const float* exp_lut;
  int8_t* dump_scores;
  uint8_t* dump_softmax;
  long long* timeline;       // test hook (p2v_attention_tc_set_timeline): per-phase clock64 stamps of CTA 0
  int skew;                  // cycles by which tile pipeline 1 starts late: its TMEM-bound pass 1 then overlaps the
                             // ALU-bound pass 3 of pipeline 0 instead of competing with the other pipeline's same pass
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
// Hand-overs between a tile's MMA warp and its eight softmax warps go through named barriers, on which the waiting
// side blocks in hardware.  (Waiting on mbarriers, the waiting warps came back from try_wait every few hundred
// cycles or less: ncu counted a fifth of all issued instructions in the softmax warps' loops in the first version,
// and later still 11 % - 3.3 M polls per launch - in the MMA warps' wait for the probabilities, on two of the four
// schedulers.)  Per item and tile four hand-overs follow each other strictly: S ready (MMA -> softmax), P written
// (softmax -> MMA), O ready (MMA -> softmax), region handed back (softmax -> MMA).  Two barrier ids per tile are enough:
// kNbToSoft for the two the MMA warp arrives on and the softmax warps wait on, kNbToMma for the two the other way
// round - a thread can only reach the second use of an id after the first one has completed.
constexpr int kNbToSoft = 9, kNbToMma = 11;      // + tile; ids 1..8 are the pair barriers
__device__ __forceinline__ void tile_barrier_wait(int id) { asm volatile("bar.sync %0, 288;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void tile_barrier_release(int id) { asm volatile("bar.arrive %0, 288;" ::"r"(id) : "memory"); }
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

// ---- the softmax passes of one thread (= one score row) ---------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
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
// RNE + unsigned saturation of two floats into bytes 0 and 1, `upper`'s low half into bytes 2 and 3 (one F2IP)
__device__ __forceinline__ uint32_t pack2_u8f(float lo, float hi, uint32_t upper) {
  uint32_t r;
  const int a = __float2int_rn(hi), b = __float2int_rn(lo);
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(upper));
  return r;
}
// base | (byte `J` of w) << 8: the table address of one packed index (base has a zero byte 1)
template <int J>
__device__ __forceinline__ uint32_t tab_addr(uint32_t w, uint32_t base) {
  return __byte_perm(w, base, 0x7604 | (J << 4));
}

// Pass 1 over one 16-column piece: raw accumulators -> biased score codes clamp(RNE(acc * mul) + zp, -128, 127) + 128,
// four to a word, and the row extrema of the raw accumulators (monotone in the code).  cnt < 16: the tail columns
// are beyond n; their codes are never used.
template <bool kMask>
__device__ __forceinline__ void code_piece(const uint32_t (&v)[16], int cnt, float mul, float c0, uint32_t (&w)[4], int& mx, int& mn) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float2 f0 = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(mul, mul), make_float2(c0, c0));
    const float2 f1 = ffma2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), make_float2(mul, mul), make_float2(c0, c0));
    w[j >> 2] = pack2_u8f(f0.x, f0.y, pack2_u8f(f1.x, f1.y, 0u));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (!kMask || j + e < cnt) {
        mx = max(mx, (int)v[j + e]);
        mn = min(mn, (int)v[j + e]);
      }
    }
  }
}

// Pass 2 over one piece: exact sum of e(d) over its first `cnt` codes.  bias: (255 - rowmax code) in every byte.
template <bool kMask>
__device__ __forceinline__ void sum_piece(const uint32_t (&w)[4], int cnt, uint32_t bias, uint32_t base_e, double (&acc)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i] + bias;            // per byte: 255 - d, no carries (code <= rowmax)
    const uint32_t e0 = lds32(tab_addr<0>(x, base_e)), e1 = lds32(tab_addr<1>(x, base_e));
    const uint32_t e2 = lds32(tab_addr<2>(x, base_e)), e3 = lds32(tab_addr<3>(x, base_e));
    if (!kMask || 4 * i < cnt) acc[0] += __hiloint2double((int)e0, 0);
    if (!kMask || 4 * i + 1 < cnt) acc[1] += __hiloint2double((int)e1, 0);
    if (!kMask || 4 * i + 2 < cnt) acc[2] += __hiloint2double((int)e2, 0);
    if (!kMask || 4 * i + 3 < cnt) acc[3] += __hiloint2double((int)e3, 0);
  }
}

// One element of pass 3: 2^(15-k) as a 16-bit integer from the bracketed 1 / (3e) (see file header).  The exponent
// field comes out of a multiply-high (FMA pipe): the ALU pipe issues a warp instruction only every other cycle.
__device__ __forceinline__ uint32_t prob_one(uint32_t addr, float fsum, float sixth, uint32_t& guard) {
  const float2 r = lds64f(addr);
  const float2 u = ffma2(make_float2(fsum, fsum), r, make_float2(sixth, sixth));
  guard |= __float_as_uint(u.x) ^ __float_as_uint(u.y);
  return shr_clamp(0x100000u, __umulhi(__float_as_uint(u.x), 512u));   // 0x100000 >> (bits >> 23)
}
// Pass 3 over one piece of 16 keys: its two probability byte planes.  ONE instantiation serves every piece (the three
// passes are kept small enough for the instruction caches: a first version with per-case variants of this loop,
// unrolled over the warp's chunks, was 158 KB of code and starved on instruction fetch); the two special cases are
// patched afterwards by the caller: keys beyond n (ragged last piece) and the maximum of a peaked row.
// Returns the OR of (low-bracket bits ^ high-bracket bits): a set exponent bit = some element sits next to a step.
__device__ __forceinline__ uint32_t prob_piece(const uint32_t (&w)[4], uint32_t bias, uint32_t base_r, float fsum, float sixth,
                                               uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  uint32_t guard = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i] + bias;
    const uint32_t p0 = prob_one(tab_addr<0>(x, base_r), fsum, sixth, guard);
    const uint32_t p1 = prob_one(tab_addr<1>(x, base_r), fsum, sixth, guard);
    const uint32_t p2 = prob_one(tab_addr<2>(x, base_r), fsum, sixth, guard);
    const uint32_t p3 = prob_one(tab_addr<3>(x, base_r), fsum, sixth, guard);
    const uint32_t a01 = p0 | (p1 << 16), a23 = p2 | (p3 << 16);
    lo[i] = __byte_perm(a01, a23, 0x6420);
    hi[i] = __byte_perm(a01, a23, 0x7531);
  }
  return guard;
}
// 0x80 in every byte of v that is 0xff (exact per byte)
__device__ __forceinline__ uint32_t bytes_ff(uint32_t v) {
  const uint32_t y = ~v;
  return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}

template <bool kDump>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_q32,
                    const __grid_constant__ CUtensorMap tm_q16, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_out, const TcArgs a) {
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw_addr = smem_u32(tc_smem_raw);
  const uint32_t a0 = (raw_addr + 1023u) & ~1023u;
  uint8_t* const win = tc_smem_raw - raw_addr;          // generic pointer of shared-window address 0
  uint8_t* const stage_ptr[2] = {win + a0, win + kRegionCAddr};
  uint8_t* const ostage = win + a0 + kOffOstage;
  TcTail& s = *reinterpret_cast<TcTail*>(win + kRegionCAddr + kStageBytes);
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n;
  const int ntiles = n > 128 ? 2 : 1;
  const int nmma = max(16, (n + 15) & ~15);          // N of the S MMA
  const int nchunks = (n + 31) >> 5;                 // 32-key chunks = K steps of P V
  const int my_items = ((int)a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (a0 + kRegionABytes > kTabAddr) __trap();       // static shared memory crept in: the layout above no longer holds

  if (warp == kTcProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_q128); tma_prefetch_desc(&tm_q32); tma_prefetch_desc(&tm_q16);
    tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_out);
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], (uint32_t)ntiles);   // one commit per tile pipeline
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s.s_full[t], 1);
      mbar_init(&s.o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == kTcMmaWarp0) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (warp < kTcSoftWarps) {
    // ---- table (warps 0-7, 32 entries each) and the initial accumulator bias of this warp's half of its lanes' columns --
    if (warp < 8) {
      const int d = warp * 32 + lane;                 // this lane evaluates entry d ...
      const float e = a.exp_lut[d];
      s.lut[d] = e;
      const float r3 = __fdiv_rn(1.0f, 3.0f * e);       // 3e is exact (<= 24 significant bits)
      const float rlo = __fmul_rn(r3, 1.0f - 9.5367431640625e-07f);   // 1 -+ 2^-20
      const float rhi = __fmul_rn(r3, 1.0f + 9.5367431640625e-07f);
      const uint32_t ehi = (uint32_t)__double2hiint((double)e);
      for (int j = 0; j < 32; ++j) {                  // ... and the warp writes each of its entries once per lane replica
        uint8_t* rowp = win + kTabAddr + (uint32_t)(255 - (warp * 32 + j)) * 256u;
        const uint32_t ej = __shfl_sync(0xffffffffu, ehi, j);
        const float lj = __shfl_sync(0xffffffffu, rlo, j), hj = __shfl_sync(0xffffffffu, rhi, j);
        reinterpret_cast<uint32_t*>(rowp)[lane] = ej;
        if (lane < 16) reinterpret_cast<float2*>(rowp + 128)[lane] = make_float2(lj, hj);
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
  pdl_wait();   // tables, barriers and the TMEM bias above are static; q / k / v are the previous kernel's output

  if (warp == kTcProducerWarp) {
    // ---- TMA producer -----------------------------------------------------------------------------------------------
    if (elect_one()) {
      const uint32_t bytes = 8192u + (ntiles == 2 ? 5120u : 0u) + (uint32_t)(kTcMaxN * 64) + (uint32_t)(kTcVRows * 64);
      for (int i = 0; i < my_items; ++i) {
        const int item = blockIdx.x + i * gridDim.x, img = item / a.heads, head = item % a.heads;
        const int st = i & 1;
        uint8_t* const sp = stage_ptr[st];
        while (!mbar_test(&s.empty[st], ((i >> 1) & 1) ^ 1)) __nanosleep(1000);   // a whole item of slack
        mbar_expect_tx(&s.full[st], bytes);
        tma_load_3d(sp + kOffQ, &tm_q128, &s.full[st], head * 64, 0, img);
        tma_load_3d(sp + kOffK, &tm_k, &s.full[st], (a.heads + head) * 64, 0, img);
        if (ntiles == 2) {
          // rows 128.. of the second tile, 32 at a time, rotated over the four lane quarters from item to item so
          // that the short tile (69 rows for n = 197) loads every scheduler's warps equally in the long run
          const int rot = i & 3;
          tma_load_3d(sp + kOffQ + 8192 + ((0 + rot) & 3) * 2048, &tm_q32, &s.full[st], head * 64, 128, img);
          tma_load_3d(sp + kOffQ + 8192 + ((1 + rot) & 3) * 2048, &tm_q32, &s.full[st], head * 64, 160, img);
          tma_load_3d(sp + kOffQ + 8192 + ((2 + rot) & 3) * 2048, &tm_q16, &s.full[st], head * 64, 192, img);
        }
        tma_load_3d(sp + kOffV, &tm_v, &s.full[st], (2 * a.heads + head) * 64, 0, img);
      }
    }
  } else if (warp == kTcMmaWarp0 || warp == kTcMmaWarp0 + 1) {
    // ---- MMA issuers: one warp (one elected lane) per tile pipeline ----------------------------------------------------
    const int t = warp - kTcMmaWarp0;
    if (t < ntiles) {
      const uint32_t idesc_s = umma_idesc_i8x(128, (uint32_t)nmma, true, true, false);
      const uint32_t idesc_pv = umma_idesc_i8x(128, 64, false, true, true);
      const uint32_t tile = tmem_base + t * kTcTileCols;
      for (int i = 0; i < my_items; ++i) {
        const int st = i & 1;
        const uint32_t sa = (st ? kRegionCAddr : a0);
        if (lane == 0) {
          // S of item i: its operands have landed (the tile's TMEM region was handed back at the end of the last round)
          mbar_wait(&s.full[st], (i >> 1) & 1);
          tc_fence_after_sync();
          const uint64_t dq = umma_desc_sw64(sa + kOffQ + t * 8192), dk = umma_desc_sw64(sa + kOffK);
          tc_mma_i8(tile, dq, dk, idesc_s, 1u);                 // accumulates onto the 1.5 * 2^23 bias
          tc_mma_i8(tile, dq + 2, dk + 2, idesc_s, 1u);         // second half of the head dimension: +32 bytes
          tc_commit(&s.s_full[t]);
          mbar_wait(&s.s_full[t], i & 1);
          tc_fence_after_sync();
          tc_fence_before_sync();
        }
        __syncwarp();
        tile_barrier_release(kNbToSoft + t);
        tile_barrier_wait(kNbToMma + t);   // P V of item i: all eight warps of the tile have written their probability planes
        if (lane == 0) {
          tc_fence_after_sync();
          const uint64_t dv = umma_desc_sw64(sa + kOffV);
          for (int ks = 0; ks < nchunks; ++ks) {
            const uint64_t dvk = dv + (uint64_t)(ks * (2048 >> 4));
            tc_mma_i8_ts(tile + kColOhi, tile + kColPhi + 8 * ks, dvk, idesc_pv, (uint32_t)(ks != 0));
            tc_mma_i8_ts(tile + kColOlo, tile + (ks < 4 ? kColPlo0 + 8 * ks : kColPlo1 + 8 * (ks - 4)), dvk, idesc_pv,
                         (uint32_t)(ks != 0));
          }
          tc_commit(&s.o_full[t]);
          tc_commit(&s.empty[st]);      // this pipeline is done with the stage once these MMAs retire
          mbar_wait(&s.o_full[t], i & 1);
          tc_fence_after_sync();
          tc_fence_before_sync();
        }
        __syncwarp();
        tile_barrier_release(kNbToSoft + t);
        tile_barrier_wait(kNbToMma + t);   // epilogue done, accumulator bias restored: the region is free for the next S
        tc_fence_after_sync();
      }
    }
  } else if (warp < kTcSoftWarps) {
    // ---- softmax + epilogue warps -------------------------------------------------------------------------------------
    const int t = warp >> 3, hf = (warp >> 2) & 1, q = warp & 3;
    const int pair = t * 4 + q;                       // the two warps (hf = 0, 1) that share 32 rows
    if (t < ntiles) {
      const uint32_t tile = tmem_base + ((uint32_t)(q * 32) << 16) + t * kTcTileCols;
      const float sixth = __fmul_rn(0.16666667f, __uint_as_float((127u - 120u) << 23));
      const uint32_t base_e = kTabAddr + lane * 4, base_r = kTabAddr + 128 + (lane & 15) * 8;
      const int sh = a.out_shift;
      const int half_m1 = (1 << (sh - 1)) - 1 + (a.out_zp << sh);   // RNE shift with the zero point folded in
      const int inv_sh = sh >= 2 ? (1 << (32 - sh)) : 0;
      uint8_t* const ost = ostage + pair * 2048;

      for (int i = 0; i < my_items; ++i) {
        const int item = blockIdx.x + i * gridDim.x, img = item / a.heads, head = item % a.heads;
        const int seg = t == 0 ? q : ((q - (i & 3)) & 3);        // which 32 rows of the tile this warp's lanes hold
        const int row0 = t * 128 + seg * 32;
        const int row = row0 + lane;
        const bool warp_on = row0 < n && (t == 0 || seg < 3);
        const bool valid = warp_on && row < n;
        // lanes without a row (beyond n, or staging rows no load wrote) see a constant score row
        const float mul = valid ? a.score_mul : 0.f;
        const float c0 = valid ? a.c0 : 128.f;
        auto stamp = [&](int phase) {
          if (a.timeline != nullptr && blockIdx.x == 0 && lane == 0 && i < 12)
            a.timeline[(i * 16 + warp) * 8 + phase] = clock64();
        };
        stamp(0);
        if (i == 0 && t == 1 && a.skew > 0) {   // start the second tile pipeline out of phase with the first (see TcArgs)
          const long long t0 = clock64();
          while (clock64() - t0 < a.skew) __nanosleep(256);
        }
        tile_barrier_wait(kNbToSoft + t);
        tc_fence_after_sync();
        stamp(1);
        if (warp_on) {
          // ---- pass 1 (the only read of the raw scores): codes, four to a word, parked in the first four columns of their
          // own 16-column piece; row extrema of this warp's chunks, then of both halves ----
          int mx = (int)0x80000000, mn = 0x7fffffff;
          // whole 32-key chunks first (no masks, no branches in the loop), then the ragged last chunk in the warp that
          // owns it (n = 197: 5 keys; with the masked variants inside the loop n = 197 ran slower than n = 208)
          const int full_chunks = n >> 5;
#pragma unroll 1
          for (int c = hf; c < full_chunks; c += 2) {
            uint32_t v0[16], v1[16], w[4];
            tmem_ld_32x16(tile + 32 * c, v0);
            tmem_ld_32x16(tile + 32 * c + 16, v1);
            tmem_ld_wait();
            code_piece<false>(v0, 16, mul, c0, w, mx, mn);
            tmem_st_32x4(tile + 32 * c, w);
            code_piece<false>(v1, 16, mul, c0, w, mx, mn);
            tmem_st_32x4(tile + 32 * c + 16, w);
          }
          if ((n & 31) != 0 && hf == (full_chunks & 1)) {
            const int c = full_chunks, cnt0 = n & 31;
            uint32_t v0[16], v1[16], w[4];
            tmem_ld_32x16(tile + 32 * c, v0);
            if (cnt0 > 16) tmem_ld_32x16(tile + 32 * c + 16, v1);
            tmem_ld_wait();
            if (cnt0 >= 16) code_piece<false>(v0, 16, mul, c0, w, mx, mn);
            else code_piece<true>(v0, cnt0, mul, c0, w, mx, mn);
            tmem_st_32x4(tile + 32 * c, w);
            if (cnt0 > 16) {
              code_piece<true>(v1, cnt0 - 16, mul, c0, w, mx, mn);
              tmem_st_32x4(tile + 32 * c + 16, w);
            }
          }
          s.x_minmax[pair][hf][lane] = make_int2(mx, mn);
          tmem_ld_wait_st();
          pair_barrier(pair);
          {
            const int2 o = s.x_minmax[pair][hf ^ 1][lane];
            mx = max(mx, o.x);
            mn = min(mn, o.y);
          }
          stamp(2);
          // biased code of the row maximum (RNE and the clamp are monotone: it is the code of the largest accumulator)
          int cmaxb;
          {
            const float fm = __fmaf_rn(__int_as_float(mx), mul, c0);
            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(cmaxb) : "f"(fm));
          }
          const uint32_t bias = (uint32_t)(255 - cmaxb) * 0x01010101u;

          // ---- pass 2: exact row sum of the integer exp ----
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
          for (int c = hf; c < full_chunks; c += 2) {
            uint32_t w0[4], w1[4];
            tmem_ld_32x4(tile + 32 * c, w0);
            tmem_ld_32x4(tile + 32 * c + 16, w1);
            tmem_ld_wait();
            sum_piece<false>(w0, 16, bias, base_e, acc);
            sum_piece<false>(w1, 16, bias, base_e, acc);
          }
          if ((n & 31) != 0 && hf == (full_chunks & 1)) {
            const int c = full_chunks, cnt0 = n & 31;
            uint32_t w0[4], w1[4];
            tmem_ld_32x4(tile + 32 * c, w0);
            if (cnt0 > 16) tmem_ld_32x4(tile + 32 * c + 16, w1);
            tmem_ld_wait();
            if (cnt0 >= 16) sum_piece<false>(w0, 16, bias, base_e, acc);
            else sum_piece<true>(w0, cnt0, bias, base_e, acc);
            if (cnt0 > 16) sum_piece<true>(w1, cnt0 - 16, bias, base_e, acc);
          }
          const double part = (acc[0] + acc[1]) + (acc[2] + acc[3]);     // integers < 2^53: exact in any order
          s.x_sum[pair][hf][lane] = part;
          pair_barrier(pair);
          const float fsum = __double2float_rn(part + s.x_sum[pair][hf ^ 1][lane]);   // exact integer -> RNE, as u64 -> f32
          stamp(3);
          // u = S / (3e) + 1/6 is evaluated scaled by 2^-120 (exact), which puts its exponent field into 5 .. 31:
          // a shift count, 2^(15-k) = 0x100000 >> field
          const float fsum_s = __fmul_rn(fsum, __uint_as_float((127u - 120u) << 23));
          // A row whose maximum holds more than 2/3 of the mass (code 0) leaves the regular step pattern of the code
          // function at that one element: patched below, in the warps that have such a row
          const int k_top = softmax_log_code(fsum, s.lut[0], 16);
          const bool peakw = __any_sync(0xffffffffu, k_top == 0);

          // ---- pass 3: probabilities 2^(15-k) as two byte planes, written to TMEM as the A operand of P V ----
          // The planes of chunk c go to columns 8c.. (high) and 208 + 8c.. / 56 + 8(c - 4).. (low): columns of chunks
          // 0, 1 and 2, whose code words (columns 0-3, 16-19 of each chunk) the pair is still reading.  Chunks 0 / 2
          // belong to the even warp of the pair, chunk 1 to the odd one, so a pair barrier after each warp has loaded
          // its first and its second chunk keeps every write behind the reads of BOTH warps (P(1), P(3) -> chunk 0 and
          // P(4), P(6) -> chunk 1 after barrier one; P(5) -> chunk 2 after barrier two).
#pragma unroll 1
          for (int it = 0; it < 4; ++it) {
            const int c = hf + 2 * it;
            if (c >= nchunks) {
              if (it < 2) pair_barrier(pair);
              continue;
            }
            uint32_t wv[2][4];
            const int cnt0 = n - 32 * c;
            tmem_ld_32x4(tile + 32 * c, wv[0]);
            if (cnt0 > 16) tmem_ld_32x4(tile + 32 * c + 16, wv[1]);
            tmem_ld_wait();
            if (it < 2) pair_barrier(pair);
            uint32_t hi[8], lo[8];                    // the chunk's two planes: words 0-3 keys 0-15, words 4-7 keys 16-31
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cnt = min(cnt0 - 16 * h, 16);
              uint32_t (&w)[4] = wv[h];
              uint32_t (&ph)[4] = *reinterpret_cast<uint32_t(*)[4]>(&hi[4 * h]);
              uint32_t (&pl)[4] = *reinterpret_cast<uint32_t(*)[4]>(&lo[4 * h]);
              if (cnt <= 0) {
#pragma unroll
                for (int x = 0; x < 4; ++x) ph[x] = pl[x] = 0u;
                continue;
              }
              const uint32_t guard = prob_piece(w, bias, base_r, fsum_s, sixth, ph, pl);
              bool redo = (guard & 0x7f800000u) != 0u;
              if (cnt < 16) {     // ragged piece: keys beyond n have probability 0 and no say in the guard
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int left = cnt - 4 * x;
                  const uint32_t m = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : (1u << (8 * left)) - 1u);
                  ph[x] &= m;
                  pl[x] &= m;
                }
              }
              if (peakw && k_top == 0) {   // this row's maximum (index byte 0xff, unique): probability 2^15, not 2^14
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const uint32_t top = bytes_ff(w[x] + bias), fill = (top >> 7) * 255u;
                  ph[x] = (ph[x] & ~fill) | top;
                  pl[x] &= ~fill;
                }
              }
              redo = __any_sync(0xffffffffu, redo);
              if (redo || kDump) {
                // redo (rare): some element within 2^-20 of a step of the code function: all elements of the piece again
                // with the exact IEEE-division formula
                const int col = 32 * c + 16 * h;
                int8_t* dsc = kDump ? a.dump_scores + ((int64_t)item * n + row) * n + col : nullptr;
                uint8_t* dsm = kDump ? a.dump_softmax + ((int64_t)item * n + row) * n + col : nullptr;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  uint32_t p4[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int g = (int)((w[x] >> (8 * e)) & 0xffu);
                    const int j = 4 * x + e;
                    if (redo) p4[e] = j < cnt ? tc_exact_prob16(fsum, s.lut[cmaxb - min(g, cmaxb)]) : 0u;
                    else p4[e] = ((pl[x] >> (8 * e)) & 0xffu) | (((ph[x] >> (8 * e)) & 0xffu) << 8);
                    if (kDump && valid && j < cnt) {
                      dsc[j] = (int8_t)(g - 128);
                      dsm[j] = (uint8_t)(p4[e] ? __clz(p4[e]) - 16 : 16);
                    }
                  }
                  if (redo) {
                    const uint32_t a01 = p4[0] | (p4[1] << 16), a23 = p4[2] | (p4[3] << 16);
                    pl[x] = __byte_perm(a01, a23, 0x6420);
                    ph[x] = __byte_perm(a01, a23, 0x7531);
                  }
                }
              }
            }
            tmem_st_32x8(tile + kColPhi + 8 * c, hi);
            tmem_st_32x8(tile + (c < 4 ? kColPlo0 + 8 * c : kColPlo1 + 8 * (c - 4)), lo);
          }
          tmem_ld_wait_st();
        }
        tc_fence_before_sync();
        tile_barrier_release(kNbToMma + t);
        stamp(4);

        // ---- epilogue: O = (O_hi << 8) + O_lo -> RNE shift to the qact2 grid -> int8, out through a TMA store ----
        tile_barrier_wait(kNbToSoft + t);
        tc_fence_after_sync();
        stamp(5);
        if (warp_on) {
          if (hf == 0 && lane == 0) tma_store_wait_read();   // the previous item's store has finished reading the staging tile
          pair_barrier(pair);
#pragma unroll
          for (int ck = 0; ck < 2; ++ck) {            // this warp's 32 of the 64 head channels, 16 at a time
            uint32_t oh[16], ol[16];
            tmem_ld_32x16(tile + kColOhi + 32 * hf + 16 * ck, oh);
            tmem_ld_32x16(tile + kColOlo + 32 * hf + 16 * ck, ol);
            tmem_ld_wait();
            uint32_t w8[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              int qv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // (hi << 8) + lo, then RNE(. / 2^sh) + zp = (x + 2^(sh-1) - 1 + parity(x >> sh) + (zp << sh)) >> sh; the
                // two shifts are multiply-highs by 2^(32-sh) (FMA pipe; sh >= 2)
                const int acc2 = (int)oh[4 * w + e] * 256 + (int)ol[4 * w + e];
                qv[e] = sh >= 2 ? __mulhi(acc2 + half_m1 + (__mulhi(acc2, inv_sh) & 1), inv_sh)
                                : (acc2 + half_m1 + ((acc2 >> sh) & 1)) >> sh;
              }
              w8[w] = pack4_s8(qv[0], qv[1], qv[2], qv[3]);
            }
            // row `lane` of the 32 x 64-byte staging tile, 16-byte chunks XOR-swizzled by (row >> 1) & 3 (SWIZZLE_64B)
            const int chunk = (2 * hf + ck) ^ ((lane >> 1) & 3);
            *reinterpret_cast<uint4*>(ost + lane * 64 + chunk * 16) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
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
        stamp(7);
        tile_barrier_release(kNbToMma + t);
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

constexpr int kTcSmemBytes = (int)kTcSmemEnd;   // as if dynamic shared memory started at window address 0
static long long* g_tc_timeline = nullptr;
static int g_tc_skew = 0;

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
  // mul = 2^-s, 0 <= s <= 15: acc * mul + zp + 128 (|acc| < 2^20) is then exact in fp32, so the pack rounds once
  if (!(p->score_mul > 0.f) || frexpf(p->score_mul, &ex) != 0.5f || ex > 1 || ex < -14) return false;
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
  // zp + 128 - 1.5 * 2^23 * mul: exact in fp32 for mul = 2^-s, s <= 20 (checked by attention_tc_applicable)
  a.c0 = (float)((double)p->score_zp + 128.0 - 12582912.0 * (double)p->score_mul);
  int ex = 0;
  frexp(p->out_mul, &ex);
  a.out_shift = 1 - ex;
  a.out_zp = (int)p->out_zp;
  a.exp_lut = p->exp_lut;
  a.dump_scores = p->dump_scores;
  a.dump_softmax = p->dump_softmax;
  a.timeline = g_tc_timeline;
  a.skew = g_tc_skew;
  const int grid = a.items < kNumSMs ? a.items : kNumSMs;
  if (p->dump_scores != nullptr)
    P2V_CHECK_CUDA(launch_pdl(4, attention_tc_kernel<true>, dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, tq128, tq32, tq16, tk, tv, to, a));
  else
    P2V_CHECK_CUDA(launch_pdl(4, attention_tc_kernel<false>, dim3(grid), dim3(kTcThreads), kTcSmemBytes, st, tq128, tq32, tq16, tk, tv, to, a));
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
extern "C" int p2v_attention_tc_set_skew(int cycles) {
  p2v::g_tc_skew = cycles;
  return P2V_OK;
}
