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
//   warps 4-7   epilogue          : TMEM -> registers -> int8 codes -> global, overlapped with the next
//                                   tile's MMAs through the second accumulator (tmem full/empty mbarriers)
#include <cuda.h>

#include "p2v_common.cuh"
#include "p2v_math.cuh"

namespace p2v {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kBlockK = 128;  // bytes == int8 elements: one 128B swizzle row
constexpr int kUmmaK = 32;    // K per tcgen05.mma for 8-bit operands
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTileBytes = kBlockM * kBlockK;  // 16 KiB per operand per stage
constexpr int kGemmThreads = 256;
constexpr uint32_t kTmemCols = kAccStages * kBlockN;  // 256 columns of 32-bit accumulators

struct GemmSmem {
  alignas(1024) uint8_t a[kStages][kTileBytes];
  alignas(1024) uint8_t b[kStages][kTileBytes];
  EpiChannel chan[kAccStages][kBlockN];
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

__device__ __forceinline__ uint32_t pack4(int q0, int q1, int q2, int q3) {
  return (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) |
         ((uint32_t)(q3 & 0xff) << 24);
}

// Epilogue of one 32-column chunk held by one thread (= one output row).
template <uint32_t FLAGS>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], const EpiChannel* chan, const GemmArgs& g,
                                               int row, int col0) {
  const int ncols = min(32, g.n - col0);
  if (ncols <= 0) return;
  const int64_t off = (int64_t)row * g.ld_out + col0;
  if (g.raw_acc != nullptr) {
    for (int j = 0; j < ncols; ++j) g.raw_acc[(int64_t)row * g.n + col0 + j] = (int32_t)acc[j];
    return;
  }
  const bool vec = (ncols == 32) && ((g.ld_out & 15) == 0) && ((col0 & 15) == 0);
  int res[32];
  if (FLAGS & EPI_RESIDUAL) {
    if (vec) {
      const uint4* rp = reinterpret_cast<const uint4*>(g.epi.residual + off);
      uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int j = 0; j < 32; ++j) res[j] = (int)(int8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xff);
    } else {
      for (int j = 0; j < 32; ++j) res[j] = j < ncols ? (int)g.epi.residual[off + j] : 0;
    }
  }
  int q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const EpiChannel c = chan[j];
    int code = epilogue_code<FLAGS>((int)acc[j], c, g.epi.out_zp);
    if (FLAGS & EPI_RESIDUAL) {
      res[j] = residual_code(code, res[j], c);  // res[] now holds the block-level code
    }
    q[j] = code;
  }
  if (g.epi.out_f32 != nullptr) {
    for (int j = 0; j < ncols; ++j)
      g.epi.out_f32[(int64_t)row * g.n + col0 + j] = fmul(fsub((float)q[j], g.epi.out_zp), chan[j].out_scale);
  }
  const int* fin = (FLAGS & EPI_RESIDUAL) ? res : q;
  if (vec) {
    uint4 o0 = make_uint4(pack4(fin[0], fin[1], fin[2], fin[3]), pack4(fin[4], fin[5], fin[6], fin[7]),
                          pack4(fin[8], fin[9], fin[10], fin[11]), pack4(fin[12], fin[13], fin[14], fin[15]));
    uint4 o1 = make_uint4(pack4(fin[16], fin[17], fin[18], fin[19]), pack4(fin[20], fin[21], fin[22], fin[23]),
                          pack4(fin[24], fin[25], fin[26], fin[27]), pack4(fin[28], fin[29], fin[30], fin[31]));
    uint4* op = reinterpret_cast<uint4*>(g.out + off);
    op[0] = o0;
    op[1] = o1;
    if ((FLAGS & EPI_RESIDUAL) && g.epi.aux_codes != nullptr) {
      uint4* ap = reinterpret_cast<uint4*>(g.epi.aux_codes + off);
      ap[0] = make_uint4(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7]),
                         pack4(q[8], q[9], q[10], q[11]), pack4(q[12], q[13], q[14], q[15]));
      ap[1] = make_uint4(pack4(q[16], q[17], q[18], q[19]), pack4(q[20], q[21], q[22], q[23]),
                         pack4(q[24], q[25], q[26], q[27]), pack4(q[28], q[29], q[30], q[31]));
    }
  } else {
    for (int j = 0; j < ncols; ++j) g.out[off + j] = (int8_t)fin[j];
    if ((FLAGS & EPI_RESIDUAL) && g.epi.aux_codes != nullptr)
      for (int j = 0; j < ncols; ++j) g.epi.aux_codes[off + j] = (int8_t)q[j];
  }
}

__device__ __forceinline__ void load_channels(EpiChannel* chan, const p2v_epilogue& e, int n0, int n, int tid,
                                              int nthreads, uint32_t flags) {
  for (int j = tid; j < kBlockN; j += nthreads) {
    EpiChannel c = {0.f, 0.f, 1.f, 1.f, 0.f, 1.f};
    const int col = n0 + j;
    if (col < n && e.acc_scale != nullptr) {
      c.acc_scale = e.acc_scale[col];
      c.bias = e.bias ? e.bias[col] : 0.f;
      c.out_scale = e.out_scale[col];
      c.out_rscale = e.out_rscale ? e.out_rscale[col] : 0.f;
      if (flags & EPI_RESIDUAL) {
        c.res_scale = e.res_scale[col];
        c.out2_scale = e.out2_scale[col];
      }
    }
    chan[j] = c;
  }
}

template <uint32_t FLAGS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  GemmSmem& s = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
      mbar_init(&s.acc_empty[i], 4);  // one arrive per epilogue warp
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
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.empty[stage], phase ^ 1);
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
        mbar_wait(&s.acc_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
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
    const int ew = warp - 4;           // TMEM lane quarter this warp may access (warp id % 4)
    const int etid = threadIdx.x - 128;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
      // per-channel constants of this tile -> smem (only the epilogue warps touch chan[acc])
      load_channels(s.chan[acc], g.epi, n0, g.n, etid, 128, FLAGS);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&s.acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < kBlockN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * kBlockN + c * 32, v);
        tmem_ld_wait();
        if (row < g.m) epilogue_chunk<FLAGS>(v, &s.chan[acc][c * 32], g, row, n0 + c * 32);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
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

template <uint32_t FLAGS>
static int configure_one() {
  P2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_tc_kernel<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kGemmSmemBytes));
  return P2V_OK;
}

// Opt every instantiation into > 48 KiB of dynamic shared memory.  Done once per process, outside any
// stream capture.
int gemm_configure() {
  static int state = 1;  // 1 = not yet done
  if (state == 1) {
    int rc = configure_one<0>();
    if (!rc) rc = configure_one<1>();
    if (!rc) rc = configure_one<2>();
    if (!rc) rc = configure_one<3>();
    if (!rc) rc = configure_one<4>();
    if (!rc) rc = configure_one<5>();
    if (!rc) rc = configure_one<6>();
    if (!rc) rc = configure_one<7>();
    if (rc) return rc;
    state = 0;
  }
  return P2V_OK;
}

template <uint32_t FLAGS>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, cudaStream_t st) {
  const int smem = kGemmSmemBytes;
  int rc = gemm_configure();
  if (rc) return rc;
  const int tiles = ((g.m + kBlockM - 1) / kBlockM) * ((g.n + kBlockN - 1) / kBlockN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  gemm_i8_tc_kernel<FLAGS><<<grid, kGemmThreads, smem, st>>>(ta, tb, g);
  P2V_CHECK_CUDA(cudaGetLastError());
  return P2V_OK;
}

static int dispatch_tc(uint32_t flags, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g,
                       cudaStream_t st) {
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
