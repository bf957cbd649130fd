"""The reference's operators called one at a time on the GPU (module-level API, fp32 in / fp32 out) and the
per-module model path that forward hooks rely on, checked against the CPU oracle and the fused engine."""
import copy

import numpy as np
import pytest
import torch

from oracle import fakequant_forward as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cabi():
    from diff_vit_b200 import _cabi
    _cabi.lib()
    return _cabi


@pytest.fixture(scope='module')
def micro_cuda(micro_model):
    """The calibrated (on CPU) micro model moved to the GPU: calibration state follows .cuda()."""
    engine, micro_model._engine = micro_model._engine, None   # the bound engine holds library handles
    try:
        model = copy.deepcopy(micro_model)
    finally:
        micro_model._engine = engine
    return model.cuda()


class _Q:
    """Minimal quantizer stand-in: QIntLayerNorm.forward only reads .scale."""
    def __init__(self, scale):
        self.scale = scale


@pytest.mark.parametrize('pot', [True, False])
@pytest.mark.parametrize('rows,d', [(197, 192), (64, 384), (33, 768), (5, 100)])
def test_qint_layernorm_module_matches_oracle(cabi, rows, d, pot):
    import diff_vit_b200 as dv
    g = torch.Generator().manual_seed(rows * d + pot)
    base = 2.0 ** -4 if pot else 0.0713
    in_scale = base * 2.0 ** torch.randint(0, 4, (d,), generator=g).float()      # PTF: base * 2^m per channel
    in_scale[0] = base
    out_scale = torch.tensor(2.0 ** -3 if pot else 0.1371)
    cs = 2.0 ** torch.randint(-2, 3, (d,), generator=g).float() if pot else torch.rand(d, generator=g) + 0.5
    codes = torch.randint(-128, 128, (1, rows, d), generator=g).float()
    x = codes * in_scale
    ln = dv.QIntLayerNorm(d, eps=1e-6)
    with torch.no_grad():
        ln.weight.copy_(torch.rand(d, generator=g) + 0.5)
        ln.bias.copy_(torch.randn(d, generator=g) * 0.1)
        ln.weight[1] = -0.7
    ln.mode = 'int'
    want_codes, want = orc.int_layernorm(x, in_scale, out_scale * cs, ln.weight.detach(), ln.bias.detach())
    got = ln.cuda()(x.cuda(), _Q(in_scale.cuda()), _Q(out_scale.cuda()), cs.cuda()).cpu()
    diff = (got - want).abs() / (out_scale * cs)
    # the oracle's log2/sum run in fp32 on the CPU; anything beyond one code step would be a bug
    assert diff.max() <= 1.0001
    assert (diff > 1e-3).float().mean() <= 1e-3
    # half-width input scale (in_scale_expand of the reference, layers.py:257-259)
    if d % 2 == 0:
        half = in_scale[:d // 2]
        want2 = orc.int_layernorm(codes * half.repeat(2), half.repeat(2), out_scale * cs,
                                  ln.weight.detach().cpu(), ln.bias.detach().cpu())[1]
        got2 = ln(codes.cuda() * half.repeat(2).cuda(), _Q(half.cuda()), _Q(out_scale.cuda()), cs.cuda(), 2).cpu()
        assert ((got2 - want2).abs() / (out_scale * cs)).max() <= 1.0001


def test_qint_layernorm_module_rejects_off_grid_input(cabi):
    import diff_vit_b200 as dv
    ln = dv.QIntLayerNorm(64).cuda()
    ln.mode = 'int'
    x = torch.full((1, 4, 64), 1e9, device='cuda')
    with pytest.raises(ValueError):
        ln(x, _Q(torch.full((64,), 1e-3, device='cuda')), _Q(torch.tensor(0.1, device='cuda')))


@pytest.mark.parametrize('scale', [2.0 ** -3, 2.0 ** -2, 2.0 ** -5, 0.0917])
@pytest.mark.parametrize('shape', [(2, 3, 197, 197), (1, 2, 10, 10), (1, 1, 64, 300)])
def test_qint_softmax_module_matches_oracle(cabi, shape, scale):
    import diff_vit_b200 as dv
    g = torch.Generator().manual_seed(int(scale * 1e4) + shape[-1])
    codes = torch.randint(-128, 128, shape, generator=g).float()
    codes[..., 0, :] = 5.0                       # a flat row
    codes[..., 1, :] = -128.0
    codes[..., 1, 3] = 127.0                     # a peaked row
    x = codes * scale
    sm = dv.QIntSoftmax(log_i_softmax=True, quant=True, bit_type=dv.ptq.bit_type.BIT_TYPE_DICT['uint4'])
    want_codes, want = orc.log_int_softmax(x, torch.tensor(scale), 4)
    got = sm(x.cuda(), torch.tensor(scale, device='cuda')).cpu()
    # exact integer row sums vs the oracle's fp32 summation: a code may move by one at a rounding tie
    k_got = -torch.log2(got.clamp_min(2.0 ** -16))
    assert torch.equal(k_got, k_got.round())
    assert (k_got - want_codes).abs().max() <= 1 and (k_got != want_codes).float().mean() <= 1e-3
    assert torch.equal(got[..., 0, :], want[..., 0, :]) and torch.equal(got[..., 1, :], want[..., 1, :])
    # not a softmax over log codes: without log_i_softmax the module is the float softmax, as in the reference
    assert torch.allclose(dv.QIntSoftmax()(x.cuda(), None).cpu(), x.softmax(-1), atol=1e-6)


def test_requant_eltwise_matches_formula(cabi):
    g = torch.Generator().manual_seed(11)
    rows, d = 777, 384
    a = torch.randint(-128, 128, (rows, d), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (rows, d), generator=g, dtype=torch.int8)
    sa = 0.05 * 2.0 ** torch.randint(0, 4, (d,), generator=g).float()
    sb = 0.03 * 2.0 ** torch.randint(0, 4, (d,), generator=g).float()
    so = 0.11 * 2.0 ** torch.randint(0, 4, (d,), generator=g).float()
    for zp, use_b in ((0.0, True), (3.0, True), (-2.0, False)):
        val = a.float() * sa + (b.float() * sb if use_b else 0)
        want = (val / so + zp).round().clamp(-128, 127).to(torch.int8)
        out = torch.empty((rows, d), dtype=torch.int8, device='cuda')
        ad, bd, sad, sbd, sod = a.cuda(), b.cuda(), sa.cuda(), sb.cuda(), so.cuda()
        cabi.check(cabi.lib().p2v_requant_eltwise(ad.data_ptr(), bd.data_ptr() if use_b else None, out.data_ptr(), rows, d,
                                                  sad.data_ptr(), sbd.data_ptr() if use_b else None, sod.data_ptr(), zp,
                                                  cabi.current_stream()))
        assert torch.equal(out.cpu(), want)
    assert cabi.lib().p2v_requant_eltwise(ad.data_ptr(), bd.data_ptr(), out.data_ptr(), rows, d, sad.data_ptr(), None,
                                          sod.data_ptr(), 0.0, cabi.current_stream()) != 0


def test_percentile_select_kernel_exact(cabi):
    from diff_vit_b200.ptq.observer import build_observer, gpu_stats
    from diff_vit_b200.ptq.bit_type import BIT_TYPE_DICT
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1_000_003, generator=g) * 4
    x[:4] = torch.tensor([0.0, -0.0, 1e-30, -1e-30])
    ref = torch.sort(x).values
    ranks = [0, 1, 500000, 1000002, 999990]
    assert torch.equal(gpu_stats.order_statistics(x.cuda(), ranks).cpu(), ref[ranks])
    # the observer on a CUDA activation: same value as torch.quantile on the same tensor (<= 16M elements) ...
    obs = build_observer('percentile', 'activation', BIT_TYPE_DICT['int8'], 'layer_wise')
    act = x[:1_000_000].reshape(10, 100, 1000)
    obs.update(act.cuda())
    assert float(obs.max_val) == float(torch.quantile(act.cuda().reshape(-1), 0.99999))
    assert float(obs.min_val) == float(torch.quantile(act.cuda().reshape(-1), 1 - 0.99999))
    # ... and as np.percentile beyond, where torch.quantile refuses and the reference goes through the host
    big = torch.randn(16_500_000, generator=g)
    hi, lo = gpu_stats.quantile_pair(big.cuda(), 0.99999, big.numel())
    assert float(hi) == float(np.float32(np.percentile(big.numpy(), 99.999)))
    assert float(lo) == float(np.float32(np.percentile(big.numpy(), (1 - 0.99999) * 100)))


def test_qlinear_qconv_modules_fake_quantize_weights(cabi, micro_cuda):
    import diff_vit_b200 as dv
    blk = micro_cuda.blocks[0]
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 10, 128, generator=g).cuda()
    for bits in (8, 4):
        lin = blk.attn.proj
        y = lin(x, [], bits)
        q = lin.quantizer
        s, z = q.dic_scale['int%d' % bits].reshape(-1, 1), q.dic_zero_point['int%d' % bits].reshape(-1, 1)
        lo, hi = (-128, 127) if bits == 8 else (-8, 7)
        wq = ((lin.weight / s + z).round().clamp(lo, hi) - z) * s
        assert torch.equal(y, torch.nn.functional.linear(x, wq, lin.bias))
    with pytest.raises(KeyError):
        blk.attn.proj(x, [], 3)
    conv = micro_cuda.patch_embed.proj
    img = torch.randn(2, 3, 48, 48, generator=g).cuda()
    y = conv(img, 8)
    q = conv.quantizer
    s, z = q.dic_scale['int8'].reshape(-1, 1, 1, 1), q.dic_zero_point['int8'].reshape(-1, 1, 1, 1)
    wq = ((conv.weight / s + z).round().clamp(-128, 127) - z) * s
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = torch.nn.functional.conv2d(img, wq, conv.bias, conv.stride)
    assert torch.allclose(y, want, atol=1e-5, rtol=1e-5)
    # CPU tensors are refused: there is no CPU implementation of the quantized operators
    with pytest.raises(RuntimeError):
        copy.deepcopy(blk.attn.proj).cpu()(x.cpu(), [], 8)


@pytest.mark.parametrize('tag', ['w8', 'mixed'])
def test_per_module_path_with_hooks_matches_engine(micro_cuda, micro_golden, tag):
    """Forward hooks (cka_utility.get_activations / modeldiff_p2.add_hooks style) switch the model to the
    per-module kernels; what the hooks see are the dequantized tensors of the fused engine's codes."""
    import diff_vit_b200 as dv
    z = micro_golden
    bc = {'w8': [8] * 10, 'mixed': [int(v) for v in z['mixed/bit_config']]}[tag]
    x = torch.from_numpy(z['x_eval']).cuda()
    model = micro_cuda
    fused, flops, _ = model(x, bc, False)
    _, dump = model.integer_engine().forward_dump(x, bc)

    seen = {}
    handles = []
    names = ['blocks.0.attn.qact1', 'blocks.0.qact2', 'blocks.1.mlp.qact1', 'blocks.1.attn.qact_attn1', 'qact2']
    mods = dict(model.named_modules())
    for n in names:
        handles.append(mods[n].register_forward_hook(lambda m, i, o, n=n: seen.__setitem__(n, o.detach())))
    handles.append(mods['blocks.1.attn.log_int_softmax'].register_forward_hook(
        lambda m, i, o: seen.__setitem__('softmax', o.detach())))
    try:
        assert model._hooked()
        out, flops2, gd = model(x, bc, False)
    finally:
        for h in handles:
            h.remove()
    assert not model._hooked()
    assert flops2 == flops and gd == []
    step = float(model.act_out.quantizer.scale)
    assert (out - fused).abs().max() <= step
    assert (out != fused).float().mean() <= 0.02
    total = bad = 0
    for n in names:
        q = mods[n].quantizer
        codes = (seen[n] / q.scale.reshape(-1).to(seen[n].device) + q.zero_point.reshape(-1).to(seen[n].device)).round()
        got = codes.cpu().numpy().astype(np.int64)
        want = dump['act/' + n].astype(np.int64).reshape(got.shape)
        d = np.abs(got - want)
        assert d.max() <= 1, n
        total += d.size
        bad += int((d != 0).sum())
    assert bad <= 2e-3 * total
    k = -torch.log2(seen['softmax'].clamp_min(2.0 ** -20)).round().clamp(max=16).cpu().numpy().astype(np.int64)
    want = dump['softmax/blocks.1.attn.log_int_softmax'].astype(np.int64).reshape(k.shape)
    assert (k != want).mean() <= 2e-3
    # the analysis scripts read these attributes after a hooked forward (cka_utility.py:44-47)
    assert model.blocks[0].attn.qkv_output is not None and model.blocks[1].mlp.fc1_output is not None
    # explicit opt-in, and a bit_config with an fp32 layer (-1), run per module as well
    model.per_module = True
    try:
        out2, _, _ = model(x, bc, False)
    finally:
        model.per_module = False
    assert torch.equal(out2, out)
    # (on a copy: as in the reference, a -1 entry switches that block's LayerNorm to float mode for good)
    engine, model._engine = model._engine, None
    try:
        scratch = copy.deepcopy(model)
    finally:
        model._engine = engine
    bc_fp = list(bc)
    bc_fp[3] = -1
    out3, _, _ = scratch(x, bc_fp, False)
    assert out3.shape == out.shape and torch.isfinite(out3).all()
    assert scratch.blocks[0].norm2.mode == 'ln' and model.blocks[0].norm2.mode == 'int'
