"""Per-kernel parity on the GPU, every call through the C ABI (ctypes, raw device pointers).
Checkers: exact int64 matmul in numpy and the host build of the kernels' scalar arithmetic
(tests/hostmath), which the CPU suite pins to the reference's codes."""
import ctypes as C

import numpy as np
import pytest
import torch

import hostmath

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cabi():
    from diff_vit_b200 import _cabi
    _cabi.check(_cabi.lib().p2v_check_device(0))
    return _cabi


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rand_i8(rng, *shape, lo=-128, hi=127):
    return rng.integers(lo, hi + 1, size=shape, dtype=np.int64).astype(np.int8)


GEMM_SHAPES = [(128, 128, 128), (256, 384, 128), (300, 192, 192), (50, 1000, 384), (1970, 1536, 384),
               (1970, 384, 1536), (6, 16, 128), (129, 136, 80), (20000, 384, 384), (50432, 1152, 384),
               (5000, 4000, 256), (3000, 576, 192)]


@pytest.fixture(params=[1, 2], ids=['streaming', 'weight_stationary'])
def gemm_mode(request, cabi):
    """Both tensor-core kernels: operand-streaming (any k) and weight-stationary (k <= 384)."""
    cabi.check(cabi.lib().p2v_gemm_set_mode(request.param))
    yield request.param
    cabi.check(cabi.lib().p2v_gemm_set_mode(0))


@pytest.mark.parametrize('m,n,k', GEMM_SHAPES)
def test_gemm_tensor_core_accumulators_exact(cabi, gemm_mode, m, n, k):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a, w = _rand_i8(rng, m, k), _rand_i8(rng, n, k)
    ad, wd = torch.from_numpy(a).cuda(), torch.from_numpy(w).cuda()
    acc = torch.full((m, n), -7, dtype=torch.int32, device='cuda')
    cabi.check(cabi.lib().p2v_gemm_i8_acc(ad.data_ptr(), k, wd.data_ptr(), acc.data_ptr(), m, n, k, _stream()))
    torch.cuda.synchronize()
    ref = a.astype(np.int64) @ w.astype(np.int64).T
    np.testing.assert_array_equal(acc.cpu().numpy().astype(np.int64), ref)


def _epilogue_case(rng, n, pot, gelu, residual):
    from diff_vit_b200.plan import LinearPlan
    acc_scale = (2.0 ** rng.integers(-14, -9, size=n)).astype(np.float32)
    bias = (rng.standard_normal(n) * 0.05).astype(np.float32)
    if pot:
        out_scale = np.full(n, 2.0 ** -5, np.float32)
    else:
        out_scale = (rng.uniform(0.5, 1.0, size=n) * 2.0 ** -5 * 2.0 ** rng.integers(0, 3, size=n)).astype(np.float32)
    lp = LinearPlan(w=None, acc_scale=torch.from_numpy(acc_scale), bias=torch.from_numpy(bias),
                    out_scale=torch.from_numpy(out_scale), out_rscale=torch.from_numpy(1.0 / out_scale), out_zp=0.0,
                    flags=(1 if gelu else 0) | (4 if pot else 0))
    if residual:
        lp.res_scale = torch.from_numpy((rng.uniform(0.5, 1.0, size=n) * 2.0 ** -6).astype(np.float32))
        lp.out2_scale = torch.from_numpy((rng.uniform(0.5, 1.0, size=n) * 2.0 ** -5).astype(np.float32))
    return lp


@pytest.mark.parametrize('impl', ['p2v_gemm_i8', 'p2v_gemm_i8_simt'])
@pytest.mark.parametrize('pot,gelu,residual', [(True, False, False), (True, True, False), (False, False, True),
                                               (False, False, False), (True, False, True)])
@pytest.mark.parametrize('m,n,k', [(394, 384, 384), (197, 1000, 192), (300, 768, 768), (260, 3072, 768), (300, 768, 3072),
                                   (333, 192, 192), (130, 208, 96)])   # plain launches whose last tile / slab is ragged
@pytest.mark.parametrize('dump', [True, False], ids=['dump', 'plain'])
def test_gemm_epilogues_match_host_arithmetic(cabi, gemm_mode, impl, pot, gelu, residual, m, n, k, dump):
    if impl == 'p2v_gemm_i8_simt' and (gemm_mode == 2 or k > 384):
        pytest.skip('the CUDA-core cross-check has a single kernel; large shapes are covered by the tensor-core kernels')
    if gemm_mode == 2 and k > 384:
        pytest.skip('weight-stationary kernel needs k <= 384')
    if not dump and impl == 'p2v_gemm_i8_simt':
        pytest.skip('the launch without dump targets selects the plain tensor-core kernels; nothing changes for the cross-check')
    rng = np.random.default_rng(1 + m + n + k + 2 * pot + 4 * gelu + 8 * residual)
    a, w = _rand_i8(rng, m, k), _rand_i8(rng, n, k, lo=-100, hi=100)
    lp = _epilogue_case(rng, n, pot, gelu, residual)
    lp.w = torch.from_numpy(w)
    res = _rand_i8(rng, m, n) if residual else None
    want, want_aux, want_f32 = hostmath.gemm_epilogue(a, lp, residual=res, want_f32=True)

    dev = lambda t: None if t is None else (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).cuda().contiguous()
    keep = dict(a=dev(a), w=dev(w), acc=dev(lp.acc_scale), bias=dev(lp.bias), os=dev(lp.out_scale), ors=dev(lp.out_rscale),
                rs=dev(lp.res_scale), o2=dev(lp.out2_scale), res=dev(res))
    out = torch.zeros(m, n, dtype=torch.int8, device='cuda')
    aux = torch.zeros(m, n, dtype=torch.int8, device='cuda')
    f32 = torch.zeros(m, n, dtype=torch.float32, device='cuda')
    e = cabi.Epilogue()
    e.acc_scale, e.bias, e.out_scale, e.out_rscale = (keep[x].data_ptr() for x in ('acc', 'bias', 'os', 'ors'))
    # dump: also ask for the dequantized output and the branch codes (the general kernels); plain: what the fused
    # forward launches (no dump target; with n % 16 == 0 that selects the kernels without ragged-edge handling)
    e.flags = lp.flags | (cabi.EPI_OUT_F32 if dump else 0) | (cabi.EPI_RESIDUAL if residual else 0)
    if dump:
        e.out_f32 = f32.data_ptr()
    if residual:
        e.res_scale, e.out2_scale, e.residual = keep['rs'].data_ptr(), keep['o2'].data_ptr(), keep['res'].data_ptr()
        if dump:
            e.aux_codes = aux.data_ptr()
    fn = getattr(cabi.lib(), impl)
    cabi.check(fn(keep['a'].data_ptr(), k, keep['w'].data_ptr(), out.data_ptr(), n, m, n, k, C.byref(e), _stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
    if gelu:   # device erff vs host erff: <= 1 LSB on <= 0.1 % of elements
        assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3
    else:
        np.testing.assert_array_equal(got, want)
        if dump:
            np.testing.assert_array_equal(f32.cpu().numpy(), want_f32)
            if residual:
                np.testing.assert_array_equal(aux.cpu().numpy(), want_aux)


@pytest.mark.parametrize('pot', [True, False])
@pytest.mark.parametrize('m,n,k', [(6304, 384, 1536), (6304, 384, 384), (197 * 40, 384, 768), (300, 384, 384)])
def test_gemm_residual_small_m_automatic_kernel_choice(cabi, pot, m, n, k):
    """Automatic kernel choice (p2v_gemm_set_mode(0)) at the shard sizes of the multi-GPU runs (32 images: M = 6304,
    150 tiles of 128 x 128 on 148 SMs) and around them, plain launches of the residual epilogue: same codes as the
    host arithmetic, bit for bit.  (A 128 x 192-tile kernel that does these shapes in one round was measured at
    11.25 vs 11.14 us for fc2 and 7.5 vs 8.0 us for proj and not kept.)"""
    rng = np.random.default_rng(m + n + k + pot)
    a, w = _rand_i8(rng, m, k), _rand_i8(rng, n, k, lo=-100, hi=100)
    lp = _epilogue_case(rng, n, pot, False, True)
    lp.w = torch.from_numpy(w)
    res = _rand_i8(rng, m, n)
    want, _, _ = hostmath.gemm_epilogue(a, lp, residual=res, want_f32=True)
    dev = lambda t: (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).cuda().contiguous()
    keep = dict(a=dev(a), w=dev(w), acc=dev(lp.acc_scale), bias=dev(lp.bias), os=dev(lp.out_scale), ors=dev(lp.out_rscale),
                rs=dev(lp.res_scale), o2=dev(lp.out2_scale), res=dev(res))
    out = torch.zeros(m, n, dtype=torch.int8, device='cuda')
    e = cabi.Epilogue()
    e.acc_scale, e.bias, e.out_scale, e.out_rscale = (keep[x].data_ptr() for x in ('acc', 'bias', 'os', 'ors'))
    e.flags = lp.flags | cabi.EPI_RESIDUAL
    e.res_scale, e.out2_scale, e.residual = keep['rs'].data_ptr(), keep['o2'].data_ptr(), keep['res'].data_ptr()
    cabi.check(cabi.lib().p2v_gemm_set_mode(0))
    cabi.check(cabi.lib().p2v_gemm_i8(keep['a'].data_ptr(), k, keep['w'].data_ptr(), out.data_ptr(), n, m, n, k, C.byref(e), _stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def _ln_plan(rng, d, pot):
    from diff_vit_b200.plan import LayerNormPlan
    base = np.float32(0.0123)
    mask = (2.0 ** rng.integers(0, 4, size=d)).astype(np.float32)
    mask[rng.integers(0, d)] = 1.0
    gamma = rng.uniform(0.5, 1.5, size=d).astype(np.float32) * rng.choice([-1.0, 1.0], size=d, p=[0.1, 0.9]).astype(np.float32)
    beta = (rng.standard_normal(d) * 0.1).astype(np.float32)
    # the corners of get_MN (layers.py:234-238): A = 0, N clamped at 31 (tiny |A|) and at 0 with M clamped at 255
    # (huge |A|), plus an offset beyond the fast path's 2^21 bound
    gamma[2], gamma[3], gamma[5], gamma[6] = 0.0, 1e-9, -3e-10, 4e5
    beta[7] = 3e5
    cs = (2.0 ** rng.integers(0, 3, size=d)).astype(np.float32)
    cs2 = (2.0 ** rng.integers(0, 3, size=d)).astype(np.float32)
    s_out = np.float32(2.0 ** -5) if pot else np.float32(0.0291)
    ln_out = (s_out * cs).astype(np.float32)
    t = torch.from_numpy
    return LayerNormPlan(in_mask=t(mask), gamma=t(gamma), beta=t(beta), ln_out_scale=t(ln_out),
                         ln_out_rscale=t((1.0 / ln_out).astype(np.float32)), post_mul=t((ln_out / cs2 / s_out).astype(np.float32)),
                         post_div1=t(cs2), post_div2=float(s_out), post_zp=0.0, in_scale1=float(base), pot=int(pot))


@pytest.mark.parametrize('rows,d,stride_rows', [(197 * 3, 384, 1), (64, 192, 1), (33, 128, 1), (5, 768, 1), (4, 384, 197),
                                                (700, 768, 1), (300, 1024, 1), (90, 512, 1), (75, 640, 1), (3, 768, 197),
                                                # Swin: four / two rows per warp (d = 96 / 192) with ragged row counts, the
                                                # 4C LayerNorm of the last PatchMerging (d = 1536, general kernel)
                                                (1001, 96, 1), (3, 96, 1), (4099, 96, 1), (333, 192, 1), (1, 192, 1),
                                                (50, 1536, 1), (9, 2048, 1)])
@pytest.mark.parametrize('pot', [True, False])
@pytest.mark.parametrize('big_masks', [False, True, 'pre_clamp'], ids=['masks_le_8', 'masks_to_64', 'pre_clamp'])
def test_layernorm_int_matches_host_arithmetic(cabi, rows, d, stride_rows, pot, big_masks):
    rng = np.random.default_rng(rows + d + pot)
    p = _ln_plan(rng, d, pot)
    if big_masks == 'pre_clamp':   # an int8 QAct on the LayerNorm's own grid before the re-gridding (Swin's qact3)
        p.pre_clamp, big_masks = 1, False
        p.gamma = p.gamma * 3.0    # enough codes beyond +-127 for the clamp to matter
    if big_masks:   # beyond the PTF range {1, 2, 4, 8}: the kernel must fall back from fp32 to integer row statistics
        p.in_mask = p.in_mask * torch.from_numpy((2.0 ** rng.integers(0, 4, size=d)).astype(np.float32))
    x = _rand_i8(rng, rows * stride_rows, d)
    x[0] = 127
    x[0, ::2] = -128                                     # maximal variance row
    want, want_codes = hostmath.layernorm(x, stride_rows * d, rows, d, p)
    keep = {k: getattr(p, k).cuda() for k in ('in_mask', 'gamma', 'beta', 'ln_out_scale', 'ln_out_rscale', 'post_mul', 'post_div1')}
    c = cabi.LayerNorm()
    for k, v in keep.items():
        setattr(c, k, v.data_ptr())
    c.post_div2, c.post_zp, c.in_scale1, c.pot = p.post_div2, p.post_zp, p.in_scale1, p.pot
    c.pre_clamp = int(getattr(p, 'pre_clamp', 0))
    if c.pre_clamp:
        assert (np.abs(want_codes) > 127).mean() > 1e-3
    xd = torch.from_numpy(x).cuda()
    out = torch.zeros(rows, d, dtype=torch.int8, device='cuda')
    codes = torch.zeros(rows, d, dtype=torch.int32, device='cuda')
    cabi.check(cabi.lib().p2v_layernorm_int(xd.data_ptr(), stride_rows * d, out.data_ptr(), codes.data_ptr(), rows, d,
                                            C.byref(c), _stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(codes.cpu().numpy(), want_codes)
    np.testing.assert_array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize('b,n,heads', [(2, 197, 3), (3, 10, 2), (1, 224, 1), (2, 33, 6), (1, 1, 1)])
@pytest.mark.parametrize('spread', [1, 6])
@pytest.mark.parametrize('in_zp', [0, 7, -11])
@pytest.mark.parametrize('kernel', ['tcgen05', 'mma.sync'])
def test_attention_int_matches_host_arithmetic(cabi, b, n, heads, spread, in_zp, kernel):
    """Both attention kernels against the host arithmetic, bit for bit: the tcgen05 / TMEM / TMA kernel
    (p2v_attention_tc.cu: power-of-two grids, n <= 208) and the mma.sync kernel (p2v_attention.cu: everything else).
    in_zp != 0: asymmetric q/k/v codes (omse observer); the kernel completes the raw int8 products with row / key
    sums and must equal the direct sum over (q - z)(k - z) and p (v - z) of the host arithmetic, with non-zero
    score and output zero points as well."""
    if kernel == 'tcgen05' and (in_zp != 0 or n > 208):
        pytest.skip('outside the tcgen05 kernel: the library routes this call to the mma.sync kernel')
    from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut
    rng = np.random.default_rng(b * 100 + n + heads + spread)
    qkv = _rand_i8(rng, b * n, 3 * heads * 64, lo=max(-128, -20 * spread + in_zp), hi=min(127, 20 * spread + in_zp))
    s_att = torch.tensor([2.0 ** -4])
    p = AttentionPlan(score_mul=float(2.0 ** -7), score_zp=3.0 if in_zp else 0.0, out_mul=2.0 ** -15 * 2.0 ** -1,
                      out_zp=-5.0 if in_zp else 0.0, levels=16, exp_lut=softmax_exp_lut(s_att), in_zp=float(in_zp))
    want, want_sc, want_sm = hostmath.attention(qkv, b, n, heads, p)
    lut = p.exp_lut.cuda()
    c = cabi.Attention()
    c.score_mul, c.score_zp, c.out_mul, c.out_zp, c.softmax_levels = p.score_mul, p.score_zp, p.out_mul, p.out_zp, 16
    c.in_zp = p.in_zp
    c.exp_lut = lut.data_ptr()
    c.lut_sig_bits = p.lut_sig_bits
    assert 0 < p.lut_sig_bits <= 21
    c.force_legacy = int(kernel == 'mma.sync')
    sc = torch.zeros(b, heads, n, n, dtype=torch.int8, device='cuda')
    sm = torch.zeros(b, heads, n, n, dtype=torch.uint8, device='cuda')
    c.dump_scores, c.dump_softmax = sc.data_ptr(), sm.data_ptr()
    qd = torch.from_numpy(qkv).cuda()
    out = torch.zeros(b * n, heads * 64, dtype=torch.int8, device='cuda')
    cabi.check(cabi.lib().p2v_attention_int(qd.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), _stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    np.testing.assert_array_equal(sm.cpu().numpy(), want_sm)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    if spread == 6 and n >= 33:
        assert want_sm.max() == 16 and len(np.unique(want_sm)) >= 8    # a wide code range, including 'zero'
    # no dump pointers: same result
    c.dump_scores = c.dump_softmax = None
    out2 = torch.zeros_like(out)
    cabi.check(cabi.lib().p2v_attention_int(qd.data_ptr(), out2.data_ptr(), b, n, heads, C.byref(c), _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def _run_attention(cabi, qkv, b, n, heads, p, legacy=False, dump=True):
    lut = p.exp_lut.cuda()
    c = cabi.Attention()
    c.score_mul, c.score_zp, c.out_mul, c.out_zp, c.softmax_levels = p.score_mul, p.score_zp, p.out_mul, p.out_zp, 16
    c.in_zp, c.exp_lut, c.lut_sig_bits, c.force_legacy = p.in_zp, lut.data_ptr(), p.lut_sig_bits, int(legacy)
    sc = torch.zeros(b, heads, n, n, dtype=torch.int8, device='cuda')
    sm = torch.zeros(b, heads, n, n, dtype=torch.uint8, device='cuda')
    if dump:
        c.dump_scores, c.dump_softmax = sc.data_ptr(), sm.data_ptr()
    qd = torch.from_numpy(qkv).cuda()
    out = torch.zeros(b * n, heads * 64, dtype=torch.int8, device='cuda')
    cabi.check(cabi.lib().p2v_attention_int(qd.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), _stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy(), sc.cpu().numpy(), sm.cpu().numpy()


@pytest.mark.parametrize('n', [8, 16, 17, 128, 129, 160, 161, 192, 193, 197, 208])
def test_attention_tcgen05_token_counts_and_zero_points(cabi, n):
    """The tcgen05 attention kernel at every boundary of its tiling: one / two row tiles, ragged last key chunk of
    1..32 columns (8-, 16- and 32-column TMEM loads), second-tile segments of 1, 32, 33, 64, 65 and 80 rows; non-zero
    score and output zero points; score scales that clamp a large share of the scores at both ends."""
    from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut
    rng = np.random.default_rng(1000 + n)
    b, heads = 3, 2
    for score_mul, szp, ozp, spread in ((2.0 ** -7, 0.0, 0.0, 2), (2.0 ** -5, 5.0, -3.0, 6), (2.0 ** -9, -100.0, 9.0, 6)):
        qkv = _rand_i8(rng, b * n, 3 * heads * 64, lo=-20 * spread, hi=20 * spread)
        p = AttentionPlan(score_mul=float(score_mul), score_zp=szp, out_mul=2.0 ** -15 * 2.0 ** -2, out_zp=ozp,
                          levels=16, exp_lut=softmax_exp_lut(torch.tensor([2.0 ** -4])), in_zp=0.0)
        want, want_sc, want_sm = hostmath.attention(qkv, b, n, heads, p)
        got, sc, sm = _run_attention(cabi, qkv, b, n, heads, p)
        np.testing.assert_array_equal(sc, want_sc)
        np.testing.assert_array_equal(sm, want_sm)
        np.testing.assert_array_equal(got, want)
        got2, _, _ = _run_attention(cabi, qkv, b, n, heads, p, dump=False)
        np.testing.assert_array_equal(got2, want)
        old, _, _ = _run_attention(cabi, qkv, b, n, heads, p, legacy=True)
        np.testing.assert_array_equal(old, want)


def test_attention_tcgen05_many_items_per_cta(cabi):
    """More (image, head) items than SMs, so that every persistent CTA walks five items through its two-stage operand
    ring, both TMEM tile pipelines and all four rotations of the second row tile over the lane quarters; peaked rows
    (one dominant key: the irregular first steps of the code function) and flat rows in the same batch."""
    from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut
    rng = np.random.default_rng(77)
    b, n, heads = 110, 197, 6                      # 660 items on 148 CTAs
    qkv = _rand_i8(rng, b * n, 3 * heads * 64, lo=-40, hi=40)
    q3 = qkv.reshape(b, n, 3, heads, 64)
    q3[::3, :, 0] //= 8                            # flat rows: small queries
    q3[1::3, 5, 1] = np.clip(q3[1::3, 5, 1].astype(np.int32) * 3, -128, 127).astype(np.int8)   # a dominant key
    p = AttentionPlan(score_mul=float(2.0 ** -8), score_zp=0.0, out_mul=2.0 ** -15 * 2.0 ** -1, out_zp=0.0,
                      levels=16, exp_lut=softmax_exp_lut(torch.tensor([2.0 ** -3])), in_zp=0.0)
    want, want_sc, want_sm = hostmath.attention(qkv, b, n, heads, p)
    got, sc, sm = _run_attention(cabi, qkv, b, n, heads, p)
    np.testing.assert_array_equal(sc, want_sc)
    np.testing.assert_array_equal(sm, want_sm)
    np.testing.assert_array_equal(got, want)
    assert want_sm.min() == 0 and want_sm.max() == 16
    got2, _, _ = _run_attention(cabi, qkv, b, n, heads, p, dump=False)
    np.testing.assert_array_equal(got2, want)


def test_quant_patchify_and_embed(cabi):
    rng = np.random.default_rng(5)
    b, c, hw, p, d = 3, 3, 48, 16, 128
    x = (rng.standard_normal((b, c, hw, hw)) * 1.5).astype(np.float32)
    x[0, 0, 0, :8] = [0.046875, -0.046875, 0.078125, 4.5, -4.5, 0.015625, -0.015625, 1e-9]   # ties and clamps at 2^-5
    xd = torch.from_numpy(x).cuda()
    g = hw // p
    out = torch.zeros(b * g * g, c * p * p, dtype=torch.int8, device='cuda')
    cabi.check(cabi.lib().p2v_quant_patchify(xd.data_ptr(), out.data_ptr(), b, c, hw, hw, p, 2.0 ** -5, 0.0, _stream()))
    torch.cuda.synchronize()
    q = np.clip(np.rint(x / np.float32(2.0 ** -5)), -128, 127).astype(np.int8)
    want = q.reshape(b, c, g, p, g, p).transpose(0, 2, 4, 1, 3, 5).reshape(b * g * g, -1)
    np.testing.assert_array_equal(out.cpu().numpy(), want)

    npatch = g * g
    pe = _rand_i8(rng, b * npatch, d)
    cls = (rng.standard_normal(d) * 0.02).astype(np.float32)
    pos = (np.rint(rng.standard_normal((npatch + 1, d)) * 20) * 2.0 ** -10).astype(np.float32)
    osc = (rng.uniform(0.5, 1, size=d) * 2.0 ** -6).astype(np.float32)
    want = np.empty((b * (npatch + 1), d), np.int8)
    hostmath.lib().hm_embed(pe.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), b, npatch, d,
                            C.c_float(2.0 ** -6), C.c_float(0), C.c_float(2.0 ** -5), C.c_float(0),
                            cls.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), osc.ctypes.data_as(C.c_void_p))
    ped, clsd, posd, oscd = (torch.from_numpy(t).cuda() for t in (pe, cls, pos, osc))
    got = torch.zeros(b * (npatch + 1), d, dtype=torch.int8, device='cuda')
    cabi.check(cabi.lib().p2v_embed_assemble(ped.data_ptr(), got.data_ptr(), b, npatch, d, 2.0 ** -6, 0.0, 2.0 ** -5, 0.0,
                                             clsd.data_ptr(), posd.data_ptr(), oscd.data_ptr(), _stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_error_reporting(cabi):
    rc = cabi.lib().p2v_gemm_i8(None, 0, None, None, 0, 0, 0, 0, None, None)
    assert rc == -1 and b'null' in cabi.lib().p2v_last_error()
    with pytest.raises(cabi.P2VError):
        cabi.check(cabi.lib().p2v_quant_patchify(1, 1, 1, 3, 30, 30, 16, 1.0, 0.0, None))


def test_fast_gelu_reproduces_the_erf_codes_for_every_input(cabi):
    """The fc1 epilogue evaluates erf-GELU with one polynomial + ex2 and a guard band; whenever the guard accepts,
    the int8 code must equal RNE(gelu_erf(y) / s_out).  Swept over ALL 2^32 fp32 inputs for every power-of-two
    output grid from 2^0 to 2^-12."""
    for e in range(0, 13):
        counts = torch.zeros(3, dtype=torch.int64, device='cuda')
        cabi.check(cabi.lib().p2v_test_gelu_fast(float(2.0 ** e), counts.data_ptr(), _stream()))
        bad, rejected, total = (int(v) for v in counts.cpu())
        assert bad == 0, 'grid 2^-%d: %d accepted inputs round to a different code' % (e, bad)
        assert total > 2_000_000_000 and rejected <= 2e-3 * total, (e, rejected, total)
