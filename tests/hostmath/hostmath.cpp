// hostmath.cpp - TEST INFRASTRUCTURE.  Compiles diff_vit_b200/csrc/p2v_math.cuh (the exact scalar
// arithmetic the sm_100a kernels inline) for the host, so tests can check the integer plan and the
// kernels' fp32 op order against the oracle on a CPU-only box.  Build: g++ -O2 -ffp-contract=off.
// Never part of the product library.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../diff_vit_b200/csrc/p2v_math.cuh"

using namespace p2v;

extern "C" {

// out[m,n] codes from int32 accumulators; flags as in p2v.h
void hm_epilogue(const int32_t* acc, int m, int n, const float* acc_scale, const float* bias, const float* out_scale,
                 const float* out_rscale, float out_zp, uint32_t flags, const int8_t* residual, const float* res_scale,
                 const float* out2_scale, int8_t* out, int8_t* aux, float* out_f32) {
  for (int r = 0; r < m; ++r)
    for (int c = 0; c < n; ++c) {
      EpiChannel ch = {acc_scale[c], bias[c], out_scale[c], out_rscale[c], res_scale ? res_scale[c] : 0.f,
                       out2_scale ? out2_scale[c] : 1.f};
      int a = acc[(int64_t)r * n + c], q;
      switch (flags & 5u) {
        case 0: q = epilogue_code<0>(a, ch, out_zp); break;
        case 1: q = epilogue_code<1>(a, ch, out_zp); break;
        case 4: q = epilogue_code<4>(a, ch, out_zp); break;
        default: q = epilogue_code<5>(a, ch, out_zp); break;
      }
      if (out_f32) out_f32[(int64_t)r * n + c] = fmul(fsub((float)q, out_zp), ch.out_scale);
      if (flags & EPI_RESIDUAL) {
        if (aux) aux[(int64_t)r * n + c] = (int8_t)q;
        q = residual_code(q, residual[(int64_t)r * n + c], ch);
      }
      out[(int64_t)r * n + c] = (int8_t)q;
    }
}

void hm_layernorm(const int8_t* in, int64_t in_row_stride, int8_t* out, int32_t* ln_codes, int rows, int d,
                  const float* in_mask, const float* gamma, const float* beta, const float* ln_out_scale,
                  const float* ln_out_rscale, const float* post_mul, const float* post_div1, float post_div2,
                  float post_zp, float in_scale1, int pot) {
  std::vector<float> xq(d);
  for (int r = 0; r < rows; ++r) {
    const int8_t* src = in + r * in_row_stride;
    long long sum = 0, sumsq = 0;
    for (int c = 0; c < d; ++c) {
      int v = (int)src[c] * (int)in_mask[c];
      xq[c] = (float)v;
      sum += v;
      sumsq += (long long)v * v;
    }
    LnRow st = ln_row_stats(sum, sumsq, d, in_scale1);
    for (int c = 0; c < d; ++c) {
      float code = pot ? ln_code<true>(xq[c], st, gamma[c], beta[c], ln_out_scale[c], ln_out_rscale[c])
                       : ln_code<false>(xq[c], st, gamma[c], beta[c], ln_out_scale[c], 0.f);
      if (ln_codes) ln_codes[(int64_t)r * d + c] = (int)code;
      float v = pot ? rne(fadd(fmul(code, post_mul[c]), post_zp))
                    : rne(fadd(fdiv(fdiv(fmul(code, ln_out_scale[c]), post_div1[c]), post_div2), post_zp));
      out[(int64_t)r * d + c] = (int8_t)clamp_i(v, -128, 127);
    }
  }
}

// qkv [b, n, 3, heads, 64] -> out [b, n, heads*64]; scores/softmax dumps [b, heads, n, n]
void hm_attention(const int8_t* qkv, int8_t* out, int b, int n, int heads, float score_mul, float score_zp,
                  double out_mul, float out_zp, int levels, const float* lut, int8_t* scores, uint8_t* softmax,
                  float in_zp) {
  const int z = (int)in_zp;
  const int hd = 64;
  const int64_t rs = 3 * heads * hd;
  std::vector<int> sc(n);
  std::vector<int> kk(n);
  for (int img = 0; img < b; ++img)
    for (int h = 0; h < heads; ++h)
      for (int i = 0; i < n; ++i) {
        const int8_t* q = qkv + ((int64_t)img * n + i) * rs + h * hd;
        int mx = -128;
        for (int j = 0; j < n; ++j) {
          const int8_t* k = qkv + ((int64_t)img * n + j) * rs + (heads + h) * hd;
          int acc = 0;
          for (int d = 0; d < hd; ++d) acc += ((int)q[d] - z) * ((int)k[d] - z);
          sc[j] = clamp_i(rne(fadd(fmul((float)acc, score_mul), score_zp)), -128, 127);
          mx = sc[j] > mx ? sc[j] : mx;
        }
        unsigned long long sum = 0;
        for (int j = 0; j < n; ++j) sum += (unsigned long long)lut[mx - sc[j]];
        const float fsum = (float)sum;
        for (int j = 0; j < n; ++j) {
          kk[j] = softmax_log_code(fsum, lut[mx - sc[j]], levels);
          if (scores) {
            const int64_t o = (((int64_t)img * heads + h) * n + i) * n + j;
            scores[o] = (int8_t)sc[j];
            softmax[o] = (uint8_t)kk[j];
          }
        }
        for (int d = 0; d < hd; ++d) {
          long long acc = 0;
          for (int j = 0; j < n; ++j) {
            if (kk[j] >= levels) continue;
            const int8_t* v = qkv + ((int64_t)img * n + j) * rs + (2 * heads + h) * hd;
            acc += (long long)((int)v[d] - z) * (1LL << (15 - kk[j]));
          }
          double v = rint((double)acc * out_mul) + (double)out_zp;
          v = v < -128.0 ? -128.0 : (v > 127.0 ? 127.0 : v);
          out[((int64_t)img * n + i) * heads * hd + h * hd + d] = (int8_t)v;
        }
      }
}

void hm_quant(const float* x, int8_t* out, int64_t total, float scale, float zp) {
  for (int64_t i = 0; i < total; ++i) out[i] = (int8_t)quant_div(x[i], scale, zp, -128, 127);
}

void hm_embed(const int8_t* pe, int8_t* out, int b, int np, int d, float pe_scale, float pe_zp, float embed_scale,
              float embed_zp, const float* cls_value, const float* pos_value, const float* out_scale) {
  for (int img = 0; img < b; ++img)
    for (int t = 0; t <= np; ++t)
      for (int c = 0; c < d; ++c) {
        float xe;
        if (t == 0) {
          xe = cls_value[c];
        } else {
          float pv = fmul(fsub((float)pe[((int64_t)img * np + t - 1) * d + c], pe_zp), pe_scale);
          int qe = quant_div(pv, embed_scale, embed_zp, -128, 127);
          xe = fmul(fsub((float)qe, embed_zp), embed_scale);
        }
        float xv = fadd(xe, pos_value[(int64_t)t * d + c]);
        out[((int64_t)img * (np + 1) + t) * d + c] = (int8_t)quant_div(xv, out_scale[c], 0.f, -128, 127);
      }
}

}  // extern "C"
