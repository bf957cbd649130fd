/* p2v.h - C ABI of libp2vit_b200.so: the B200 (sm_100a) integer inference path of P2-ViT.
 *
 * The reference (LeSN-Lab/diff-ViT) has no FFI: its boundary for this path is the Python nn.Module
 * surface of models/ptq (QConv2d, QLinear, QAct, QIntLayerNorm, QIntSoftmax; models/ptq/layers.py) and
 * VisionTransformer.forward (models/vit_fquant.py:780-799).  Each entry point below names the
 * reference interface whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative p2v_status; p2v_last_error() gives the message
 *     of the calling thread's last failure;
 *   - all pointers are DEVICE pointers unless the name says host; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - activation codes are int8 row-major [rows, channels]; weights are int8 [out, in] row-major
 *     (int4 layers store their [-8,7] codes in int8);
 *   - fp32 steps follow the reference's operation order (diff_vit_b200/csrc/p2v_math.cuh).
 */
#ifndef P2V_H_
#define P2V_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  P2V_OK = 0,
  P2V_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  P2V_ERR_CUDA = -2,        /* CUDA runtime or driver error */
  P2V_ERR_UNSUPPORTED = -3, /* device is not compute capability 10.0 (kind::i8 tensor cores) */
  P2V_ERR_STATE = -4        /* handle used before it was bound / wrong batch */
} p2v_status;

const char* p2v_last_error(void);
int p2v_version(void);
/* 0 when `device` is an sm_100 part (tcgen05 kind::i8 exists on B200, not on sm_103). */
int p2v_check_device(int device);

/* ---- epilogue flags (bit-or) ---------------------------------------------------------------- */
#define P2V_EPI_GELU 1u     /* erf-GELU before re-quantization            (models/layers_quant.py:331) */
#define P2V_EPI_RESIDUAL 2u /* + residual stream, block-level PTF re-quant (models/vit_fquant.py:431,468) */
#define P2V_EPI_OUT_POT 4u  /* out scale is a power of two: use out_rscale (exact reciprocal) */
#define P2V_EPI_OUT_F32 8u  /* also write dequantized fp32 (logits)        (models/vit_fquant.py:796) */

/* Per-output-channel epilogue of an int8 GEMM: replaces F.linear/F.conv2d + the following QAct
 * (models/ptq/layers.py:86-88,173-178 and :207-220).  All vectors have n entries. */
typedef struct p2v_epilogue {
  const float* acc_scale;   /* s_in * s_w[n]                                             */
  const float* bias;        /* fp32 bias (never quantized, layers.py:178)                */
  const float* out_scale;   /* s_out[n]                                                  */
  const float* out_rscale;  /* 1/s_out[n] (exact for power-of-two scales)                */
  const float* res_scale;   /* s_res[n], residual stream scale      (RESIDUAL only)      */
  const float* out2_scale;  /* s_out2[n], block-level output scale  (RESIDUAL only)      */
  const int8_t* residual;   /* [m, n] residual codes, row stride ld_out (RESIDUAL only)  */
  int8_t* aux_codes;        /* optional [m, n]: branch codes before the residual add     */
  float* out_f32;           /* optional [m, n] fp32 (OUT_F32)                            */
  float out_zp;             /* zero point of the output quantizer (0 for minmax / ptf)   */
  uint32_t flags;
} p2v_epilogue;

/* out[m, n] = epilogue( sum_k a[m, k] * w[n, k] ), int32 accumulation on tcgen05 kind::i8.
 * a: [m, k] int8, row stride lda bytes (multiple of 16); w: [n, k] int8 contiguous; out: [m, n] int8,
 * row stride ld_out.  k must be a multiple of 16.
 * Replaces QLinear.forward / QConv2d.forward in quantized mode (models/ptq/layers.py:82-88,171-178). */
int p2v_gemm_i8(const int8_t* a, int64_t lda, const int8_t* w, int8_t* out, int64_t ld_out, int m, int n,
                int k, const p2v_epilogue* epi, void* stream);
/* Same contract on CUDA cores (dp4a); a slow cross-check used by the tests. */
int p2v_gemm_i8_simt(const int8_t* a, int64_t lda, const int8_t* w, int8_t* out, int64_t ld_out, int m,
                     int n, int k, const p2v_epilogue* epi, void* stream);
/* Kernel selection for p2v_gemm_i8 (test hook): 0 = automatic, 1 = operand-streaming kernel only,
 * 2 = weight-stationary kernel whenever k <= 384. */
int p2v_gemm_set_mode(int mode);
/* Programmatic dependent launch of the GEMM, LayerNorm and attention kernels (on by default): a kernel's CTAs may be
 * placed, and run their prologue, while the previous kernel of the stream drains; they wait for its completion before
 * touching activations (csrc/p2v_common.cuh launch_pdl).  0 restores fully serialised launches (test / A-B hook;
 * graphs captured earlier keep the setting they were captured with). */
int p2v_set_pdl(int enabled);
/* Test hook: sweeps all 2^32 fp32 inputs through the fc1 epilogue's fast erf-GELU (csrc/p2v_gemm.cu gelu_code_fast2)
 * for one power-of-two 1/s_out and accumulates into counts[3] (device, pre-zeroed): accepted elements whose int8
 * code differs from RNE(gelu_erf(y) / s_out) - must stay 0 -, guard rejections and inputs among |y| < 8. */
int p2v_test_gelu_fast(float out_rscale, unsigned long long* counts, void* stream);
/* Raw int32 accumulators of the tensor-core kernel (test hook). */
int p2v_gemm_i8_acc(const int8_t* a, int64_t lda, const int8_t* w, int32_t* acc, int m, int n, int k,
                    void* stream);

/* Input QAct + im2col for a stride-P patch convolution: x fp32 [b, c, h, w] ->
 * codes int8 [b*(h/p)*(w/p), c*p*p], K order (c, kh, kw) = conv weight.reshape(out, -1).
 * Replaces qact_input (models/vit_fquant.py:705-706) and the unfold inside F.conv2d.  p: a multiple of 4
 * (16-pixel vector path when p and w are multiples of 16, 4-pixel path otherwise: Swin's 4 x 4 patches). */
int p2v_quant_patchify(const float* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                       float zero_point, void* stream);
/* The same from 8-bit pixels x [b, c, h, w] (device): the fp32 preprocessing of the reference's loaders,
 * (pixel / 255 - mean[c]) / std[c] (torchvision ToTensor + Normalize, test_quant.py:96-110), is evaluated on the
 * device op for op, so the codes equal those of the fp32 entry on the normalised tensor.  mean / std: HOST arrays [c].
 * p: a multiple of 4 (16-pixel vector path when p and w are multiples of 16, 4-pixel path otherwise). */
int p2v_quant_patchify_u8(const uint8_t* x, int8_t* codes, int b, int c, int h, int w, int p, float scale,
                          float zero_point, const float* mean, const float* stdv, void* stream);

/* Token assembly: cls concat, qact_embed, + qact_pos(pos_embed), qact1 (PTF)
 * (models/vit_fquant.py:718-733).  pe: patch-embed codes [b*np, d]; cls_value[d] / pos_value[(np+1)*d]
 * are the dequantized cls token (after qact_embed) and position embedding (after qact_pos);
 * out: [b*(np+1), d] codes on the per-channel grid out_scale[d]. */
int p2v_embed_assemble(const int8_t* pe, int8_t* out, int b, int np, int d, float pe_scale, float pe_zp,
                       float embed_scale, float embed_zp, const float* cls_value, const float* pos_value,
                       const float* out_scale, void* stream);

/* Integer LayerNorm fused with the QAct that consumes it.  Replaces QIntLayerNorm.forward mode 'int'
 * (models/ptq/layers.py:255-289) + x / channel_scale + qact0 (models/vit_fquant.py:284-289,
 * models/layers_quant.py:307-312).  Per channel c:
 *   xq = in[r, c] * in_mask[c];  code = LN_dyadic(xq; gamma, beta, ln_out_scale[c])
 *   out[r, c] = clamp(RNE(code * post_mul[c]))           post_mul = ln_out_scale / (cs_next * s_next)
 * All of ln_out_scale, post_mul are powers of two under the minmax observer (pot != 0); with pot == 0
 * the kernel divides by ln_out_scale and by post_div[c] = cs_next[c] * ... in fp32 instead.
 * in_row_stride: bytes between consecutive rows (lets the final norm read only the CLS rows).
 * ln_codes (optional, int32 [rows, d]): the unclamped LN codes. */
typedef struct p2v_layernorm {
  const float* in_mask;        /* [d] round(in_scale / min(in_scale)) in {1,2,4,8} */
  const float* gamma;          /* [d] */
  const float* beta;           /* [d] */
  const float* ln_out_scale;   /* [d] */
  const float* ln_out_rscale;  /* [d] exact reciprocal (pot) */
  const float* post_mul;       /* [d] (pot) */
  const float* post_div1;      /* [d] cs_next (non-pot) */
  float post_div2;             /* s_next (non-pot) */
  float post_zp;
  float in_scale1;             /* min over channels of the input scale */
  int pot;
  int pre_clamp;               /* non-zero: the LN code passes an int8 QAct of its own grid (clamp to [-128, 127]) before
                                  the re-gridding - Swin's block.qact3 in front of Mlp's qact0 (swin_quant.py:392-394) */
} p2v_layernorm;
int p2v_layernorm_int(const int8_t* in, int64_t in_row_stride, int8_t* out, int32_t* ln_codes, int rows,
                      int d, const p2v_layernorm* p, void* stream);

/* Fused integer attention for one layer: per (image, head)
 *   S = Q K^T (int32) -> qact_attn1 codes -> log-int-softmax 4-bit codes -> P V -> qact2 codes.
 * Replaces models/vit_fquant.py:308-326 and QIntSoftmax.forward (models/ptq/layers.py:323-376).
 * qkv: codes [b, n, 3, heads, 64]; out: [b, n, heads*64].  exp_lut[256]: integer exp of the row-max
 * distance d = max - code, built on the host from the score scale exactly as layers.py:334-358 does.
 * Optional dumps: score codes int8 [b, heads, n, n] and softmax codes uint8 [b, heads, n, n]. */
typedef struct p2v_attention {
  float score_mul;    /* s_qkv^2 * head_dim^-0.5 / s_score (power of two under minmax) */
  float score_zp;
  double out_mul;     /* 2^-15 * s_qkv / s_out */
  float out_zp;
  int softmax_levels; /* 2^bits = 16 */
  const float* exp_lut; /* [256] fp32 view of the integer exp (exact: < 2^24 significant bits) */
  int8_t* dump_scores;
  uint8_t* dump_softmax;
  float in_zp;        /* zero point of the q/k/v codes (qact1); non-zero only with asymmetric observers (omse):
                         S = sum (q - z)(k - z) and O = sum p (v - z) are formed from the raw int8 products plus row /
                         key sums, so the tensor-core operands stay int8 */
  int32_t lut_sig_bits; /* widest entry of exp_lut in significant bits (highest minus lowest set bit + 1 of the
                           integer); exp_lut is device memory, so the host-side kernel choice needs it from the caller.
                           1..21 admits the tcgen05 / TMEM kernel (its fp64 row sums take the table words as the high
                           halves of doubles); 0 = unknown: the mma.sync kernel is used */
  int32_t force_legacy; /* test hook: non-zero pins the mma.sync kernel (cross-checks of the two kernels) */
} p2v_attention;
int p2v_attention_int(const int8_t* qkv, int8_t* out, int b, int n, int heads, const p2v_attention* p,
                      void* stream);
/* Test hook of the tcgen05 attention kernel: buf (device, 12 x 8 x 8 int64, or NULL = off) receives clock64 stamps of
 * CTA 0's softmax warps, [item][warp][phase] (tools/att_timeline.py prints them). */
int p2v_attention_tc_set_timeline(long long* buf);
/* Tuning hook: SM cycles by which the second row-tile pipeline of the tcgen05 attention kernel starts late. */
int p2v_attention_tc_set_skew(int cycles);

/* Standalone QAct on fp32 data (module-level use): out = (clamp(RNE(x/s + zp)) - zp) * s with a scale
 * per channel of the innermost (inner == 1) or of an outer dimension.  models/ptq/layers.py:207-220. */
int p2v_fake_quant_f32(const float* x, float* out, int8_t* codes, int64_t outer, int channels, int64_t inner,
                       const float* scale, const float* zero_point, int qmin, int qmax, void* stream);

/* ---- Swin (BASELINE config 5): windowed attention and the data movement around it -------------------------- */
/* Fused integer window attention of one Swin layer (W-MSA / SW-MSA): per (window, head), n = ws^2 <= 64 tokens,
 * head dimension 32,
 *   S = (q * 32^-1/2) k^T  ->  qact_attn1  ->  + relative position bias (qact_table codes)  ->  qact2
 *     ->  - 100 where the shifted-window mask separates row and key  ->  log-int-softmax  ->  P V  ->  qact3.
 * Replaces WindowAttention.forward between its qact1 and qact3 (models/swin_quant.py:188-215) together with the
 * roll / window_partition / window_reverse around it (models/swin_quant.py:362-385): rows are loaded and stored
 * through `perm`, the composition of the cyclic shift and the window partition, so qkv [images * tokens, 3 * channels]
 * (per row [3][heads][32]) and out [images * tokens, channels] stay in token order.  All quantizers symmetric and on
 * power-of-two grids (the minmax observer); the fp32 scaling of q is reproduced exactly (csrc/p2v_swin.cu). */
typedef struct p2v_window_attention {
  const int32_t* perm;     /* [windows * n] token (inside its image) of window w's i-th row                            */
  const uint8_t* region;   /* [windows * n] region id of the shift mask (swin_quant.py:317-340); NULL without a shift   */
  const float* bias;       /* [heads][n (key)][n (row)] dequantized qact_table entry of the pair's relative position    */
  const float* exp_lut;    /* [lut_n] integer exp of d = rowmax - x (layers.py:334-358); d >= lut_n reads the last one  */
  const float* r3;         /* [lut_n][2] fl32((1 -+ 2^-20) / (3 exp_lut[d])): brackets of the fast log2 code            */
  const double* exp_lut64; /* [lut_n] the same table as doubles (the exact row sum adds them without a conversion)      */
  int32_t lut_n;
  int32_t n, heads, windows, tokens, channels;   /* tokens = windows * n per image, channels = heads * 32              */
  int32_t qshift;          /* fl32(code * qscale) * 2^qshift is an integer below 2^31 for every int8 code              */
  float qscale;            /* head_dim^-1/2 as fp32 (swin_quant.py:83,190)                                             */
  double acc_scale;        /* s_qkv^2 * 2^-qshift                                                                      */
  float qk_scale;          /* qscale * s_qkv^2 (exact in fp32): the fast score is (sum_c q_c k_c) * qk_scale           */
  float err_mul;           /* 2^-24 qscale * 128 s_qkv^2 / s_a1 (rounded up): times sum_c |q_c| it bounds the fast
                              score's distance from the exact one in units of the a1 grid                              */
  float a1_scale, a1_rscale;   /* qact_attn1 grid and its exact reciprocal                                             */
  float a2_rscale;         /* 1 / s of qact2                                                                           */
  int32_t mask_int;        /* 100 / s of qact2 (an integer)                                                            */
  float out_unit;          /* 2^-15 * s_qkv: the value of one unit of sum_j 2^(15 - k_j) v_j                           */
  float out_rscale;        /* 1 / s of qact3                                                                           */
  int32_t softmax_levels;  /* 2^bits = 16                                                                              */
  int8_t* dump_a1;         /* optional (all three or none) [images * windows][heads][n][n]: qact_attn1 codes           */
  int8_t* dump_a2;         /* optional, same shape: qact2 codes (before the mask)                                      */
  uint8_t* dump_softmax;   /* optional, same shape: log2 codes (softmax_levels = probability 0)                        */
} p2v_window_attention;
int p2v_window_attention_int(const int8_t* qkv, int8_t* out, int images, const p2v_window_attention* p, void* stream);
/* out[img][r][s * seg_bytes ...] = in[img][idx[r * segs + s]][0 .. seg_bytes): the x0 / x1 / x2 / x3 concat of
 * PatchMerging (models/swin_quant.py:449-456) as one byte gather.  seg_bytes: multiple of 16. */
int p2v_gather_row_segments(const int8_t* in, int8_t* out, const int32_t* idx, int images, int rows_in, int rows_out,
                            int segs, int seg_bytes, void* stream);
/* AdaptiveAvgPool1d(1) over the tokens of in [images, tokens, channels] (codes on the power-of-two grid in_scale)
 * followed by qact3 (models/swin_quant.py:810-813): out [images, channels]. */
int p2v_avgpool_requant(const int8_t* in, int8_t* out, int images, int tokens, int channels, float in_scale,
                        float out_scale, float out_zp, void* stream);

/* The whole quantized Swin forward as one call.  Replaces SwinTransformer.forward in quantized mode
 * (models/swin_quant.py:790-817) with everything below it (PatchEmbed, BasicLayer, SwinTransformerBlock :345-399,
 * WindowAttention :177-221, Mlp, PatchMerging :445-467).  Stateless: the descriptor and the HOST arrays it points at
 * (stages, blocks) are read during the call, every device buffer belongs to the caller, all launches go to `stream`
 * - so a caller may capture the call into a CUDA graph.  workspace: p2v_swin_workspace_bytes(desc, b) bytes,
 * 1024-byte aligned.  x fp32 [b, in_chans, img, img]; logits fp32 [b, classes]; logit_codes int8 [b, classes]. */
typedef struct p2v_swin_block_desc p2v_swin_block_desc;
typedef struct p2v_swin_stage_desc p2v_swin_stage_desc;
typedef struct p2v_swin_desc p2v_swin_desc;
int64_t p2v_swin_workspace_bytes(const p2v_swin_desc* desc, int b);
int p2v_swin_launches_per_forward(const p2v_swin_desc* desc);
int p2v_swin_forward(const p2v_swin_desc* desc, const float* x, float* logits, int8_t* logit_codes, int b,
                     void* workspace, void* stream);
/* The same from 8-bit pixels x [b, in_chans, img, img] (device) with the loader's normalisation constants (HOST arrays
 * [in_chans]): (pixel / 255 - mean[c]) / std[c] is evaluated on the device op for op (as p2v_vit_forward_u8). */
int p2v_swin_forward_u8(const p2v_swin_desc* desc, const uint8_t* x, const float* mean, const float* stdv, float* logits,
                        int8_t* logit_codes, int b, void* workspace, void* stream);

/* ---- the operators on their own, fp32 in / fp32 out (module-level use) -------------------------------- */
/* QIntLayerNorm.forward in mode 'int' (models/ptq/layers.py:255-289) on dequantized fp32 rows x [rows, d]:
 * x_q = RNE(x / in_scale[c]) * in_mask[c] with in_mask = RNE(in_scale / in_scale1), in_scale1 = min(in_scale);
 * integer row statistics; dyadic (M, N) affine; out = code * out_scale[c].  *overflow_flag (device int, pre-set
 * to 0) becomes 1 if some |x_q| >= 2^20, i.e. the input was not on the in_scale grid. */
int p2v_layernorm_int_f32(const float* x, float* out, int64_t rows, int d, const float* in_scale,
                          const float* in_mask, float in_scale1, const float* gamma, const float* beta,
                          const float* out_scale, int* overflow_flag, void* stream);
/* QIntSoftmax.forward with log_i_softmax (models/ptq/layers.py:323-376) over the last dimension of x [rows, n]:
 * I-BERT integer exp with the constants (x0_int, b_int, c_int) of layers.py:334-352 and n = exp_bits, exact row
 * sum, k = log_round(RNE(sum / exp)).  out (optional) = 2^-k, 0 where k >= levels; codes (optional) = min(k, levels). */
int p2v_softmax_log_int_f32(const float* x, float* out, uint8_t* codes, int64_t rows, int n, float scale,
                            float x0_int, float b_int, float c_int, int exp_bits, int levels, void* stream);
/* QAct applied to a sum of code tensors (the residual adds of models/vit_fquant.py:449,466):
 * out = clamp(RNE((a * a_scale[c] + b * b_scale[c]) / out_scale[c] + out_zp), -128, 127); b may be NULL (re-quantize a). */
int p2v_requant_eltwise(const int8_t* a, const int8_t* b, int8_t* out, int64_t rows, int d, const float* a_scale,
                        const float* b_scale, const float* out_scale, float out_zp, void* stream);
/* int4-packed weight codes (the storage form of 4-bit layers in serialised plans, BASELINE config 4) -> the int8
 * codes the tensor-core GEMM consumes (tcgen05 has no int4 kind): byte i holds code 2i in its low nibble and code
 * 2i + 1 in its high nibble, two's complement in [-8, 7].  out has 2 * nbytes entries. */
int p2v_unpack_int4(const uint8_t* packed, int8_t* out, int64_t nbytes, void* stream);
/* One pass of an exact radix select over fp32 data (the order statistics behind torch.quantile / np.percentile in
 * models/ptq/observer/percentile.py:27-38): hist[2048] (device, uint64, ACCUMULATED) counts bits
 * [shift, shift + 11) of the order-preserving key of every element whose key matches prefix under prefix_mask.
 * key(v) = bits(v) ^ (v < 0 ? 0xffffffff : 0x80000000).  Histograms from several ranks add up, which makes the
 * select exact across a data-parallel calibration batch. */
int p2v_select_histogram(const float* x, int64_t total, uint32_t prefix, uint32_t prefix_mask, int shift,
                         unsigned long long* hist, void* stream);

/* ---- calibration statistics ------------------------------------------------------- */
/* Running range of an activation: x is fp32 [rows, channels]; per_channel = 0 reduces the whole tensor to
 * out_min[0] / out_max[0], per_channel = 1 gives one pair per (contiguous, innermost) channel.  The outputs are
 * ACCUMULATED: pre-fill with +inf / -inf.  Replaces the max/min of MinmaxObserver.update and PtfObserver.update
 * (models/ptq/observer/minmax.py:16-39, ptf.py:14-31). */
int p2v_observe_minmax(const float* x, int64_t rows, int channels, int per_channel, float* out_min, float* out_max,
                       void* stream);
/* Squared error of fake-quantizing x with each of n_scales (<= 8, HOST array) candidate scales, zero point 0:
 * out[k] (per tensor) or out[k * channels + c] (per channel) += sum (x - clamp(RNE(x / s_k), qmin, qmax) * s_k)^2,
 * fp64.  Replaces the candidate scoring of the activation PoT search (minmax.py:180-242) and of the PTF factor
 * search (ptf.py:110-131). */
int p2v_observe_scale_sse(const float* x, int64_t rows, int channels, int per_channel, const float* scales_host,
                          int n_scales, float qmin, float qmax, double* out, void* stream);

/* ---- whole-model engine ------------------------------------------------------------------------ */
typedef struct p2v_linear_desc {
  const int8_t* w;          /* [n, k] codes */
  int32_t n, k;
  p2v_epilogue epi;         /* residual / aux / out pointers are filled by the engine */
} p2v_linear_desc;

typedef struct p2v_block_desc {
  p2v_layernorm norm1, norm2;
  p2v_linear_desc qkv, proj, fc1, fc2;
  p2v_attention attn;
} p2v_block_desc;

typedef struct p2v_vit_desc {
  int32_t img_size, patch_size, in_chans, embed_dim, depth, num_heads, hidden_dim, num_classes;
  float input_scale, input_zp;             /* qact_input */
  p2v_linear_desc patch_embed;             /* conv as GEMM, epilogue = patch_embed.qact */
  float pe_scale, pe_zp, embed_scale, embed_zp;
  const float* cls_value;                  /* [d] */
  const float* pos_value;                  /* [(np+1) * d] */
  const float* embed_out_scale;            /* [d] qact1 PTF scale */
  const p2v_block_desc* blocks;            /* HOST array [depth] */
  p2v_layernorm norm;                      /* final norm (CLS rows) -> qact2 */
  p2v_linear_desc head;                    /* epilogue = act_out, OUT_F32 */
} p2v_vit_desc;

struct p2v_swin_block_desc {
  p2v_layernorm norm1, norm2;               /* norm2: LN2 + qact3 + / channel scale + mlp.qact0 (pre_clamp) */
  p2v_linear_desc qkv, proj, fc1, fc2;      /* proj / fc2 run with the residual epilogue */
  p2v_window_attention attn;
};
struct p2v_swin_stage_desc {
  int32_t height, width, dim, depth;        /* token grid, channels and blocks of the stage */
  const p2v_swin_block_desc* blocks;        /* HOST array [depth] */
  int32_t has_merge;                        /* PatchMerging behind the blocks */
  const int32_t* merge_idx;                 /* device [height / 2 * width / 2 * 4]: x0 / x1 / x2 / x3 source tokens */
  p2v_layernorm merge_norm;                 /* over 4 * dim channels */
  p2v_linear_desc reduction;                /* 4 * dim -> 2 * dim, epilogue = downsample.qact2 */
};
struct p2v_swin_desc {
  int32_t img_size, patch_size, in_chans, embed_dim, num_stages, num_classes;
  float input_scale;                        /* qact_input (symmetric) */
  p2v_linear_desc patch_embed;              /* 4 x 4 conv as GEMM, epilogue = patch_embed.qact_before_norm */
  p2v_layernorm pe_norm;                    /* patch_embed.norm + patch_embed.qact */
  const p2v_swin_stage_desc* stages;        /* HOST array [num_stages] */
  p2v_layernorm norm;                       /* final norm + qact2 */
  float pool_in_scale, pool_out_scale;      /* qact2 / qact3 around the token average */
  p2v_linear_desc head;                     /* epilogue = act_out, fp32 logits */
};

typedef struct p2v_vit p2v_vit;

/* Replaces VisionTransformer.forward in quantized mode (models/vit_fquant.py:700-799).
 * The descriptor (and the host block array) is copied; device buffers it points to stay owned by the
 * caller and must outlive the handle. */
int p2v_vit_create(const p2v_vit_desc* desc, int device, p2v_vit** out);
void p2v_vit_destroy(p2v_vit* h);
/* Bytes of device workspace the forward needs for a batch of b images. */
int64_t p2v_vit_workspace_bytes(const p2v_vit* h, int b);
/* x: fp32 [b, c, h, w] device; logits: fp32 [b, classes] device; workspace: device scratch.
 * dump (optional, device): receives every intermediate code tensor at the offsets reported by
 * p2v_vit_dump_layout.  The launch sequence for (b, workspace, logits, x) is captured into a CUDA graph
 * on first use when use_graph != 0. */
int p2v_vit_forward(p2v_vit* h, const float* x, float* logits, int8_t* logit_codes, int b, void* workspace,
                    void* dump, int use_graph, void* stream);
/* p2v_vit_forward on 8-bit pixels x [b, c, h, w] (device) with the loader's normalisation constants (HOST arrays
 * [in_chans]): a quarter of the input traffic of the fp32 entry, identical codes (see p2v_quant_patchify_u8). */
int p2v_vit_forward_u8(p2v_vit* h, const uint8_t* x, const float* mean, const float* stdv, float* logits,
                       int8_t* logit_codes, int b, void* workspace, int use_graph, void* stream);
/* End-to-end call with HOST buffers (pinned or pageable): H2D copy, forward, D2H copy, stream sync. */
int p2v_vit_forward_host(p2v_vit* h, const float* x_host, float* logits_host, int b, void* workspace,
                         void* x_dev, void* logits_dev, void* stream);
/* Number of kernels one forward launches (for bench accounting). */
int p2v_vit_launches_per_forward(const p2v_vit* h);
/* Dump layout: entry i -> name (static string), byte offset, byte size, element size. Returns count. */
int p2v_vit_dump_layout(const p2v_vit* h, int b, int i, const char** name, int64_t* offset, int64_t* bytes,
                        int32_t* elem_size);
int64_t p2v_vit_dump_bytes(const p2v_vit* h, int b);

#ifdef __cplusplus
}
#endif
#endif /* P2V_H_ */
