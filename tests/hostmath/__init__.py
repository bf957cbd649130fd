"""TEST INFRASTRUCTURE: runs a VitPlan on the CPU through the kernels' own scalar arithmetic
(diff_vit_b200/csrc/p2v_math.cuh compiled for the host by g++), integer GEMMs done in numpy int64.
Used by the CPU test-suite to pin the plan builder and the fp32 op order against the oracle; the
product never imports it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, 'libp2v_hostmath.so')
_SRC = os.path.join(_HERE, 'hostmath.cpp')
_HDR = os.path.join(_HERE, '..', '..', 'diff_vit_b200', 'csrc', 'p2v_math.cuh')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if (not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(_SRC), os.path.getmtime(_HDR))):
            subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC', '-x', 'c++',
                                   _SRC, '-o', _LIB])
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, 'detach') else t, dtype=np.float32)


def gemm_epilogue(a, lp, residual=None, want_f32=False):
    """a int8 [m,k]; lp LinearPlan -> (codes, aux, f32)"""
    w = lp.codes().numpy().astype(np.int64)
    acc = (a.astype(np.int64) @ w.T).astype(np.int32)
    m, n = acc.shape
    out = np.empty((m, n), np.int8)
    aux = np.empty((m, n), np.int8) if residual is not None else None
    f32 = np.empty((m, n), np.float32) if want_f32 else None
    flags = lp.flags | (2 if residual is not None else 0)
    keep = [_f(lp.acc_scale), _f(lp.bias), _f(lp.out_scale), _f(lp.out_rscale),
            _f(lp.res_scale) if residual is not None else None, _f(lp.out2_scale) if residual is not None else None]
    res = np.ascontiguousarray(residual) if residual is not None else None
    lib().hm_epilogue(_p(np.ascontiguousarray(acc)), m, n, _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]),
                      C.c_float(lp.out_zp), C.c_uint32(flags), _p(res), _p(keep[4]), _p(keep[5]), _p(out), _p(aux), _p(f32))
    return out, aux, f32


def layernorm(x, row_stride, rows, d, p):
    if getattr(p, 'pre_clamp', 0):
        # an int8 QAct on the LayerNorm's own grid in front of the re-gridding (Swin): LN to its clamped codes, then
        # the reference's two divisions (/ channel scale, / consumer scale)
        import copy
        q = copy.copy(p)
        q.pre_clamp, q.post_mul, q.post_div1, q.post_div2 = 0, np.ones(d, np.float32), _f(p.ln_out_scale), 1.0
        clamped, codes = layernorm(x, row_stride, rows, d, q)
        v = (clamped.astype(np.float32) * _f(p.ln_out_scale)[None, :]).astype(np.float32)
        v = ((v / _f(p.post_div1)[None, :]).astype(np.float32) / np.float32(p.post_div2)).astype(np.float32)
        return np.clip(np.rint(v), -128, 127).astype(np.int8), codes
    out = np.empty((rows, d), np.int8)
    codes = np.empty((rows, d), np.int32)
    keep = [_f(p.in_mask), _f(p.gamma), _f(p.beta), _f(p.ln_out_scale), _f(p.ln_out_rscale), _f(p.post_mul), _f(p.post_div1)]
    lib().hm_layernorm(_p(x), C.c_int64(row_stride), _p(out), _p(codes), rows, d, *[_p(k) for k in keep],
                       C.c_float(p.post_div2), C.c_float(p.post_zp), C.c_float(p.in_scale1), int(p.pot))
    return out, codes


def attention(qkv, b, n, heads, p):
    out = np.empty((b * n, heads * 64), np.int8)
    sc = np.empty((b, heads, n, n), np.int8)
    sm = np.empty((b, heads, n, n), np.uint8)
    lut = _f(p.exp_lut)
    lib().hm_attention(_p(qkv), _p(out), b, n, heads, C.c_float(p.score_mul), C.c_float(p.score_zp),
                       C.c_double(p.out_mul), C.c_float(p.out_zp), p.levels, _p(lut), _p(sc), _p(sm),
                       C.c_float(getattr(p, 'in_zp', 0.0)))
    return out, sc, sm


def run_plan(plan, x):
    """x: float32 numpy [b, c, h, w].  Returns (logits fp32, {golden-style key: codes})."""
    a = plan.arch
    b, d, np_, heads = x.shape[0], a['embed_dim'], a['num_patches'], a['num_heads']
    n = np_ + 1
    P = a['patch_size']
    g = a['img_size'] // P
    codes = {}
    xin = np.empty(x.shape, np.int8)
    xc = np.ascontiguousarray(x, dtype=np.float32)
    lib().hm_quant(_p(xc), _p(xin), C.c_int64(xc.size), C.c_float(plan.input_scale), C.c_float(plan.input_zp))
    codes['act/qact_input'] = xin
    patches = xin.reshape(b, a['in_chans'], g, P, g, P).transpose(0, 2, 4, 1, 3, 5).reshape(b * np_, -1)
    pe, _, _ = gemm_epilogue(np.ascontiguousarray(patches), plan.patch_embed)
    codes['act/patch_embed.qact'] = pe
    xs = np.empty((b * n, d), np.int8)
    keep = [_f(plan.cls_value), _f(plan.pos_value), _f(plan.embed_out_scale)]
    lib().hm_embed(_p(pe), _p(xs), b, np_, d, C.c_float(plan.pe_scale), C.c_float(plan.pe_zp),
                   C.c_float(plan.embed_scale), C.c_float(plan.embed_zp), _p(keep[0]), _p(keep[1]), _p(keep[2]))
    codes['act/qact1'] = xs
    for i, blk in enumerate(plan.blocks):
        pre = 'blocks.%d' % i
        a0, ln = layernorm(xs, d, b * n, d, blk.norm1)
        codes['ln/' + pre + '.norm1'], codes['act/' + pre + '.attn.qact0'] = ln, a0
        qkv, _, _ = gemm_epilogue(a0, blk.qkv)
        codes['act/' + pre + '.attn.qact1'] = qkv
        o, sc, sm = attention(qkv, b, n, heads, blk.attn)
        codes['act/' + pre + '.attn.qact_attn1'], codes['softmax/' + pre + '.attn.log_int_softmax'] = sc, sm
        codes['act/' + pre + '.attn.qact2'] = o
        x1, aux, _ = gemm_epilogue(o, blk.proj, residual=xs)
        codes['act/' + pre + '.attn.qact3'], codes['act/' + pre + '.qact2'] = aux, x1
        m0, ln = layernorm(x1, d, b * n, d, blk.norm2)
        codes['ln/' + pre + '.norm2'], codes['act/' + pre + '.mlp.qact0'] = ln, m0
        hid, _, _ = gemm_epilogue(m0, blk.fc1)
        codes['act/' + pre + '.mlp.qact1'] = hid
        xs, aux, _ = gemm_epilogue(hid, blk.fc2, residual=x1)
        codes['act/' + pre + '.mlp.qact2'], codes['act/' + pre + '.qact4'] = aux, xs
    cls, ln = layernorm(xs, n * d, b, d, plan.norm)
    codes['ln/norm'], codes['act/qact2'] = ln, cls
    lc, _, logits = gemm_epilogue(cls, plan.head, want_f32=True)
    codes['act/act_out'] = lc
    return logits, codes


# ---- Swin (diff_vit_b200.swin_engine plans) ---------------------------------------------------------------------------
def window_attention(qkv, images, p):
    """The integer formulation of csrc/p2v_swin.cu window_attention_kernel in numpy (fp32 / int64 / fp64 steps as in
    the kernel).  qkv int8 [images * tokens, 3 * channels] in token order -> (out int8 [images * tokens, channels],
    qact_attn1 codes, qact2 codes, log2 codes; the three in window order [images * windows, heads, n, n])."""
    f32 = np.float32
    n, heads, nw, L, Cc = p.n, p.heads, p.windows, p.tokens, p.channels
    perm = p.perm.numpy().astype(np.int64)
    x = qkv.reshape(images, L, 3, heads, 32)[:, perm].reshape(images, nw, n, 3, heads, 32)
    q, k, v = (x[:, :, :, i].transpose(0, 1, 3, 2, 4).astype(np.int64) for i in range(3))      # [img, w, head, n, 32]
    m = ((q.astype(f32) * f32(p.qscale)).astype(f32).astype(np.float64) * 2.0 ** p.qshift).astype(np.int64)
    acc = np.einsum('bwhic,bwhjc->bwhij', m, k)
    s = (acc.astype(np.float64) * p.acc_scale).astype(f32)
    c1 = np.clip(np.rint(s * f32(p.a1_rscale)), -128, 127).astype(f32)
    bias = p.bias.numpy().transpose(0, 2, 1)[None, None]                                           # [1, 1, head, row, key]
    t = (c1 * f32(p.a1_scale)).astype(f32) + bias.astype(f32)
    a2 = np.clip(np.rint(t.astype(f32) * f32(p.a2_rscale)), -128, 127).astype(np.int64)
    xm = a2.copy()
    if p.region is not None:
        rid = p.region.numpy().reshape(nw, n).astype(np.int64)
        masked = rid[:, :, None] != rid[:, None, :]                                                # [w, row, key]
        xm = xm - masked[None, :, None] * p.mask_int
    d = np.minimum(xm.max(-1, keepdims=True) - xm, p.exp_lut.numel() - 1)
    e = p.exp_lut.numpy().astype(f32)[d]
    fsum = e.astype(np.float64).sum(-1, keepdims=True).astype(f32)
    r = np.rint(fsum / e).astype(f32)
    mant, ex = np.frexp(r)                    # r = mant * 2^ex, mant in [0.5, 1)
    big = ex.astype(np.int64) - 1
    kk = np.maximum(big + (mant >= 0.75), 0)
    kk = np.minimum(kk, p.levels)
    prob = np.where(kk >= p.levels, 0, np.int64(0x8000) >> np.minimum(kk, 15))
    o = np.einsum('bwhij,bwhjc->bwhic', prob, v)
    val = (o.astype(f32) * f32(p.out_unit)).astype(f32)
    code = np.clip(np.rint(val * f32(p.out_rscale)), -128, 127).astype(np.int8)                    # [img, w, head, n, 32]
    out = np.empty((images, L, heads, 32), np.int8)
    out[:, perm] = code.transpose(0, 1, 3, 2, 4).reshape(images, nw * n, heads, 32)
    shape = (images * nw, heads, n, n)
    return out.reshape(images * L, Cc), c1.astype(np.int8).reshape(shape), a2.astype(np.int8).reshape(shape), \
        kk.astype(np.uint8).reshape(shape)


def run_swin_plan(plan, x):
    """x: float32 numpy [b, c, h, w].  Returns (logits fp32, {golden-style key: codes}) - the launch sequence of
    diff_vit_b200.swin_engine.SwinIntegerEngine._run on the host."""
    a = plan.arch
    b, P, cin = x.shape[0], a['patch_size'], a['in_chans']
    g = a['img_size'] // P
    L, Cd = g * g, a['embed_dim']
    codes = {}
    xin = np.empty(x.shape, np.int8)
    xc = np.ascontiguousarray(x, dtype=np.float32)
    lib().hm_quant(_p(xc), _p(xin), C.c_int64(xc.size), C.c_float(plan.input_scale), C.c_float(plan.input_zp))
    codes['act/qact_input'] = xin
    patches = np.ascontiguousarray(xin.reshape(b, cin, g, P, g, P).transpose(0, 2, 4, 1, 3, 5).reshape(b * L, -1))
    pe, _, _ = gemm_epilogue(patches, plan.patch_embed)
    codes['act/patch_embed.qact_before_norm'] = pe
    xs, ln = layernorm(pe, Cd, b * L, Cd, plan.pe_norm)
    codes['ln/patch_embed.norm'], codes['act/patch_embed.qact'] = ln, xs
    for si, st in enumerate(plan.stages):
        H, W = st.res
        L, Cd = H * W, st.dim
        rows = b * L
        for bi, blk in enumerate(st.blocks):
            pre = 'layers.%d.blocks.%d' % (si, bi)
            ap = blk.attn
            perm = ap.perm.numpy().astype(np.int64)
            win = lambda t, c: t.reshape(b, L, c)[:, perm].reshape(b * ap.windows, ap.n, c)
            y, ln = layernorm(xs, Cd, rows, Cd, blk.norm1)
            codes['ln/' + pre + '.norm1'], codes['act/' + pre + '.qact1'] = ln, y
            qkv, _, _ = gemm_epilogue(y, blk.qkv)
            codes['act/' + pre + '.attn.qact1'] = win(qkv, 3 * Cd)
            att, c1, c2, sm = window_attention(qkv, b, ap)
            codes['act/' + pre + '.attn.qact_attn1'], codes['act/' + pre + '.attn.qact2'] = c1, c2
            codes['act/' + pre + '.attn.qact_table'] = ap.table_codes.numpy()
            codes['softmax/' + pre + '.attn.log_int_softmax'] = sm
            codes['act/' + pre + '.attn.qact3'] = win(att, Cd)
            x1, aux, _ = gemm_epilogue(att, blk.proj, residual=xs)
            codes['act/' + pre + '.attn.qact4'], codes['act/' + pre + '.qact2'] = win(aux, Cd), x1
            m0, ln = layernorm(x1, Cd, rows, Cd, blk.norm2)
            codes['ln/' + pre + '.norm2'], codes['act/' + pre + '.qact3'] = ln, np.clip(ln, -128, 127)
            codes['act/' + pre + '.mlp.qact0'] = m0
            hid, _, _ = gemm_epilogue(m0, blk.fc1)
            codes['act/' + pre + '.mlp.qact1'] = hid
            xs, aux, _ = gemm_epilogue(hid, blk.fc2, residual=x1)
            codes['act/' + pre + '.mlp.qact2'], codes['act/' + pre + '.qact4'] = aux, xs
        if st.merge is not None:
            pre = 'layers.%d.downsample' % si
            idx = st.merge.idx.numpy().astype(np.int64)
            cat = np.ascontiguousarray(xs.reshape(b, L, Cd)[:, idx].reshape(b * (L // 4), 4 * Cd))
            y, ln = layernorm(cat, 4 * Cd, b * (L // 4), 4 * Cd, st.merge.norm)
            codes['ln/' + pre + '.norm'], codes['act/' + pre + '.qact1'] = ln, y
            xs, _, _ = gemm_epilogue(y, st.merge.reduction)
            codes['act/' + pre + '.qact2'] = xs
            L, Cd = L // 4, 2 * Cd
    y, ln = layernorm(xs, Cd, b * L, Cd, plan.norm)
    codes['ln/norm'], codes['act/qact2'] = ln, y
    s = y.reshape(b, L, Cd).astype(np.int64).sum(1)
    mean = ((s.astype(np.float32) * np.float32(plan.pool_in_scale)).astype(np.float32) / np.float32(L)).astype(np.float32)
    pooled = np.clip(np.rint((mean / np.float32(plan.pool_out_scale)).astype(np.float32)), -128, 127).astype(np.int8)
    codes['act/qact3'] = pooled
    lc, _, logits = gemm_epilogue(pooled, plan.head, want_f32=True)
    codes['act/act_out'] = lc
    return logits, codes
