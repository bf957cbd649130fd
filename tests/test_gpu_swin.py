"""BASELINE config 5 on the GPU.  (1) The integer engine (diff_vit_b200/swin_engine.py: int8 codes end to end,
tcgen05 GEMMs, integer LayerNorm, the window attention kernel) against the CPU oracle layer by layer and against the
reference's golden codes.  (2) The per-module path (every Q-module running its own sm_100a operator, fp32 library GEMMs
for the products as the reference itself does - what hooks / per_module select) against the same."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import swin_fakequant_forward as sorc
from test_swin_golden import build_swin_micro

pytestmark = pytest.mark.gpu


def capture_codes(model, x):
    """Quantized forward with hooks that turn every QAct / integer LayerNorm / log-int-softmax output back into its
    integer codes (the recipe of tests/golden/make_golden_swin.py)."""
    import diff_vit_b200 as dv
    store, hooks = {}, []

    def act_hook(name):
        def fn(mod, inp, outp):
            q = mod.quantizer
            s = q.scale.reshape(q.get_reshape_range(outp))
            zp = q.zero_point.reshape(q.get_reshape_range(outp))
            store['act/' + name] = (outp / s + zp).round().to(torch.int32).cpu()
        return fn

    def ln_hook(name):
        def fn(mod, inp, outp):
            if mod.mode == 'int':
                store['ln/' + name] = (outp / inp[2].scale.reshape(1, 1, -1)).round().to(torch.int32).cpu()
        return fn

    def sm_hook(name):
        def fn(mod, inp, outp):
            k = torch.where(outp > 0, -torch.log2(outp.clamp_min(1e-30)), torch.full_like(outp, 16.0))
            store['softmax/' + name] = k.round().to(torch.int32).cpu()
        return fn

    for name, m in model.named_modules():
        if isinstance(m, dv.QAct):
            hooks.append(m.register_forward_hook(act_hook(name)))
        elif isinstance(m, dv.QIntLayerNorm):
            hooks.append(m.register_forward_hook(ln_hook(name)))
        elif isinstance(m, dv.QIntSoftmax):
            hooks.append(m.register_forward_hook(sm_hook(name)))
    with torch.no_grad():
        logits = model(x)
    for h in hooks:
        h.remove()
    return logits, store


def _compare(got, ref, gelu_layers_only=True):
    total = bad = 0
    for k, r in ref.items():
        if k not in got:
            continue
        r = r.numpy().astype(np.int64) if torch.is_tensor(r) else np.asarray(r).astype(np.int64)
        d = np.abs(got[k].numpy().astype(np.int64).reshape(r.shape) - r)
        total += d.size
        bad += int((d != 0).sum())
        if k.endswith('.mlp.qact1') or not gelu_layers_only:
            assert d.max() <= 1 and (d != 0).mean() <= 1e-3, '%s: max %d, %.2e differ' % (k, d.max(), (d != 0).mean())
        else:
            assert d.max() == 0, '%s is bit-defined on identical inputs: max %d, %.2e differ' % (k, d.max(), (d != 0).mean())
    return total, bad


def test_swin_micro_on_gpu_vs_reference_golden_and_oracle():
    import diff_vit_b200 as dv
    from diff_vit_b200.swin_quant import extract_swin_state
    z = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
    model = build_swin_micro(z).cuda()
    dv.calibrate_model(model, [torch.from_numpy(z['x_calib']).cuda()])        # observers on the GPU kernels
    state = extract_swin_state(model)
    x = torch.from_numpy(z['x_eval'])
    logits, codes = capture_codes(model, x.cuda())
    # layer by layer on identical inputs: the oracle teacher-forced with the GPU's codes
    want, ref = sorc.forward(state, x, capture=True, override={k: v.numpy() for k, v in codes.items()})
    total, bad = _compare(codes, ref)
    assert total > 500000
    lsb = float(state['act']['act_out'][0])
    assert (logits.cpu() - want).abs().max().item() <= lsb
    # ... and against the reference's own golden codes of the first block and its logits (free-running)
    for k in ('act/qact_input', 'act/patch_embed.qact', 'ln/layers.0.blocks.0.norm1', 'act/layers.0.blocks.0.attn.qact1',
              'act/layers.0.blocks.0.attn.qact2', 'softmax/layers.0.blocks.0.attn.log_int_softmax',
              'act/layers.0.blocks.0.attn.qact3', 'act/layers.0.blocks.0.qact2'):
        g = z['w8/' + k].astype(np.int64)
        d = np.abs(codes[k].numpy().astype(np.int64).reshape(g.shape) - g)
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3, k
    assert np.abs(logits.cpu().numpy() - z['w8/logits']).max() <= 2 * lsb
    print('swin micro: %d codes compared on identical inputs, %d differ' % (total, bad))


def test_swin_tiny_batch128_config5_per_module_path_vs_oracle():
    """BASELINE config 5: swin_tiny, W8A8 PoT, shifted-window quantized attention, batch 128 synthetic images,
    random-init weights.  The full batch runs on the GPU; every quantizer's codes of 2 of the images are compared with
    the CPU oracle on identical inputs (teacher forcing), and the batch rows must not depend on the batch size."""
    import diff_vit_b200 as dv
    from diff_vit_b200.swin_quant import extract_swin_state
    torch.manual_seed(0)
    model = dv.swin_tiny_patch4_window7_224(cfg=dv.Config(True, True, 'minmax')).eval().cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    dv.calibrate_model(model, [torch.randn(8, 3, 224, 224, device='cuda', generator=g)])
    x = torch.randn(128, 3, 224, 224, device='cuda', generator=g)
    model.per_module = True          # this test is about the operators on their own; the engine has its own below
    with torch.no_grad():
        full = model(x)
    lsb = float(model.act_out.quantizer.scale)
    c = full / lsb
    assert torch.equal(c, c.round()) and c.abs().max() <= 128 and full.std() > 0
    sub = x[:2].contiguous()
    logits, codes = capture_codes(model, sub)
    assert torch.equal(logits, full[:2])
    state = extract_swin_state(model)
    want, ref = sorc.forward(state, sub.cpu(), capture=True, override={k: v.numpy() for k, v in codes.items()})
    total, bad = _compare(codes, ref)
    assert total > 3e7
    assert (logits.cpu() - want).abs().max().item() <= lsb
    print('swin_tiny config 5: %d codes compared on identical inputs, %d differ' % (total, bad))


# ---- the integer engine ---------------------------------------------------------------------------------------------
def _engine_vs_oracle(model, x_sub, bits=None):
    """Every quantizer's codes of the engine against the oracle on identical inputs (teacher forcing), exact
    accumulation: the q * 32^-1/2 operand of the window attention is not on an integer grid, so an fp32 BLAS product
    carries summation-order noise that an integer engine does not have (as with float scales, config 3)."""
    from diff_vit_b200.swin_quant import extract_swin_state
    eng = model.integer_engine()
    logits, codes = eng.forward_dump(x_sub, bits)
    state = extract_swin_state(model)
    want, ref = sorc.forward(state, x_sub.cpu(), bits, capture=True, accum='fp64',
                             override={k: v.numpy() for k, v in codes.items()})
    assert set(ref) <= set(codes), sorted(set(ref) - set(codes))
    total, bad = _compare(codes, ref)
    lsb = float(state['act']['act_out'][0])
    assert (logits.cpu() - want).abs().max().item() <= lsb
    return logits, codes, total, bad


def test_swin_micro_engine_vs_oracle_and_reference_golden():
    import diff_vit_b200 as dv
    z = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
    model = build_swin_micro(z).cuda()
    dv.calibrate_model(model, [torch.from_numpy(z['x_calib']).cuda()])
    x = torch.from_numpy(z['x_eval']).cuda()
    logits, codes, total, bad = _engine_vs_oracle(model, x)
    assert total > 500000
    with torch.no_grad():
        eager = model(x)                       # un-hooked model call: the engine (first call eager, then the graph)
        replay = model(x)
    assert torch.equal(eager, logits) and torch.equal(replay, logits)
    assert model._engine_off is None and model.integer_engine().launches > 30
    # the reference's own run (fp32 BLAS accumulation, free-running): first block and logits
    for k in ('act/qact_input', 'act/patch_embed.qact', 'ln/layers.0.blocks.0.norm1', 'act/layers.0.blocks.0.attn.qact1',
              'act/layers.0.blocks.0.attn.qact_attn1', 'act/layers.0.blocks.0.attn.qact2',
              'softmax/layers.0.blocks.0.attn.log_int_softmax', 'act/layers.0.blocks.0.attn.qact3',
              'act/layers.0.blocks.0.qact2'):
        g = z['w8/' + k].astype(np.int64)
        d = np.abs(codes[k].numpy().astype(np.int64).reshape(g.shape) - g)
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3, k
    lsb = float(model.act_out.quantizer.scale)
    assert np.abs(logits.cpu().numpy() - z['w8/logits']).max() <= 2 * lsb
    # 4-bit weights through the same engine
    n = model.num_linear_layers()
    _, _, total4, bad4 = _engine_vs_oracle(model, x, [4] * n)
    # a hook sends the model back to the per-module path, whose logits agree within the GELU / BLAS allowance
    model.per_module = True
    with torch.no_grad():
        pm = model(x)
    assert (pm - logits).abs().max().item() <= 2 * lsb
    print('swin micro engine: %d codes compared on identical inputs, %d differ (W8), %d of %d (W4)' % (total, bad, bad4, total4))


def test_swin_tiny_engine_batch128_config5_vs_oracle():
    """BASELINE config 5 on the integer engine: swin_tiny, W8A8 PoT, batch 128 synthetic images, random-init weights.
    The full batch runs as one CUDA-graph replay; every quantizer's codes of 2 of the images are compared with the CPU
    oracle on identical inputs, and the rows of the batch must not depend on the batch size."""
    import diff_vit_b200 as dv
    torch.manual_seed(0)
    model = dv.swin_tiny_patch4_window7_224(cfg=dv.Config(True, True, 'minmax')).eval().cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    dv.calibrate_model(model, [torch.randn(8, 3, 224, 224, device='cuda', generator=g)])
    x = torch.randn(128, 3, 224, 224, device='cuda', generator=g)
    with torch.no_grad():
        first = model(x)
        full = model(x)                        # graph replay
    assert model._engine_off is None, model._engine_off
    assert torch.equal(first, full)
    lsb = float(model.act_out.quantizer.scale)
    c = full / lsb
    assert torch.equal(c, c.round()) and c.abs().max() <= 128 and full.std() > 0
    sub = x[:2].contiguous()
    logits, codes, total, bad = _engine_vs_oracle(model, sub)
    assert torch.equal(logits, full[:2])
    assert total > 3e7
    # the same model with 4-bit weights everywhere, and with the reference's kind of mixed assignment
    n = model.num_linear_layers()
    _, _, total4, bad4 = _engine_vs_oracle(model, sub, [4] * n)
    mixed = [8 if i % 3 == 0 else 4 for i in range(n)]
    _, _, totalm, badm = _engine_vs_oracle(model, sub, mixed)
    print('swin_tiny config 5 on the integer engine: %d codes compared on identical inputs, %d differ; W4: %d of %d; '
          'mixed 4 / 8: %d of %d' % (total, bad, bad4, total4, badm, totalm))


def test_swin_base_engine_vs_oracle():
    """swin_base (C = 128, heads 4 / 8 / 16 / 32, 24 blocks, the 2048-wide LayerNorm of the last PatchMerging) through
    the same engine: 2 images, every quantizer's codes against the oracle on identical inputs."""
    import diff_vit_b200 as dv
    torch.manual_seed(0)
    model = dv.swin_base_patch4_window7_224(cfg=dv.Config(True, True, 'minmax')).eval().cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    dv.calibrate_model(model, [torch.randn(4, 3, 224, 224, device='cuda', generator=g)])
    x = torch.randn(2, 3, 224, 224, device='cuda', generator=g)
    with torch.no_grad():
        full = model(x)
    assert model._engine_off is None, model._engine_off
    logits, codes, total, bad = _engine_vs_oracle(model, x)
    assert torch.equal(logits, full)
    assert total > 5e7
    print('swin_base on the integer engine: %d codes compared on identical inputs, %d differ' % (total, bad))


# ---- the window attention kernel on its own -----------------------------------------------------------------------------
@pytest.fixture(scope='module')
def cabi():
    from diff_vit_b200 import _cabi
    _cabi.lib()
    return _cabi


def _wa_plan(rng, heads, res, ws, shift, s1, sa, s2, s3, st):
    from diff_vit_b200.swin_engine import _SwinBuilder
    from diff_vit_b200.swin_quant import relative_position_index
    C_ = heads * 32
    b = _SwinBuilder.__new__(_SwinBuilder)
    t = lambda v: torch.tensor([v], dtype=torch.float32)
    z = torch.zeros(1)
    b.arch = {'softmax_bits': 4}
    b.P = {'a.qkv.weight': torch.zeros(3 * C_, C_),
           'a.relative_position_bias_table': torch.from_numpy(rng.standard_normal(((2 * ws - 1) ** 2, heads)).astype(np.float32)) * 4 * st,
           'a.relative_position_index': relative_position_index((ws, ws))}
    b.s = {'act': {'a.qact1': (t(s1), z, -128, 127), 'a.qact_attn1': (t(sa), z, -128, 127), 'a.qact2': (t(s2), z, -128, 127),
                   'a.qact3': (t(s3), z, -128, 127), 'a.qact_table': (t(st), z, -128, 127)}}
    return b.window_attention('a', heads, res, ws, shift)


@pytest.mark.parametrize('heads,res,ws,shift,images', [(3, (14, 14), 7, 0, 3), (3, (14, 14), 7, 3, 3), (1, (7, 7), 7, 0, 5),
                                                       (2, (8, 8), 4, 2, 2), (4, (16, 16), 8, 4, 1), (6, (28, 28), 7, 3, 2)])
@pytest.mark.parametrize('spread,scales', [(12, (2.0 ** -3, 2.0 ** -2, 2.0 ** -2, 2.0 ** -4)),      # flat rows
                                            (127, (2.0 ** -3, 2.0 ** 0, 2.0 ** -1, 2.0 ** -3)),      # peaked rows, saturating scores
                                            (40, (2.0 ** -2, 2.0 ** -4, 2.0 ** -5, 2.0 ** -5)),      # fine score grid: long exp table
                                            (60, (2.0 ** -4, 2.0 ** -3, 2.0 ** -3, 2.0 ** -6))])
def test_window_attention_kernel_matches_its_integer_formulation(cabi, heads, res, ws, shift, images, spread, scales):
    """`p2v_window_attention_int` (dp4a fast paths with tie guards, five items per CTA, ragged last CTA) against the
    exact integer formulation in numpy (tests/hostmath.window_attention, itself pinned to the reference golden by
    tests/test_swin_plan.py): scores after both re-quantisations, log2 codes and the output, bit for bit - flat and
    peaked rows, saturating scores, shifted windows, window sizes 4 / 7 / 8."""
    import ctypes as C
    import hostmath
    from diff_vit_b200.swin_engine import _Bound
    rng = np.random.default_rng(heads * 1000 + res[0] * 10 + shift + spread)
    s1, sa, s2, s3 = scales
    p = _wa_plan(rng, heads, res, ws, shift, s1, sa, s2, s3, st=2.0 ** -3)
    L, C_ = res[0] * res[1], heads * 32
    qkv = rng.integers(-spread, spread + 1, size=(images * L, 3 * C_)).astype(np.int8)
    qkv[0, :C_] = -128                                       # a row of the most negative code (|q| bound, saturation)
    want, w1, w2, wsm = hostmath.window_attention(qkv, images, p)
    holder = _Bound.__new__(_Bound)
    holder.device, holder.keep = torch.device('cuda'), []
    d = holder.attn(p)
    nw = images * p.windows
    d1 = torch.zeros(nw, heads, p.n, p.n, dtype=torch.int8, device='cuda')
    d2 = torch.zeros_like(d1)
    d3 = torch.zeros(nw, heads, p.n, p.n, dtype=torch.uint8, device='cuda')
    xq = torch.from_numpy(qkv).cuda()
    out = torch.zeros(images * L, C_, dtype=torch.int8, device='cuda')
    out2 = torch.zeros_like(out)
    st = cabi.current_stream()
    cabi.check(cabi.lib().p2v_window_attention_int(xq.data_ptr(), out2.data_ptr(), images, C.byref(d), st))   # plain kernel
    d.dump_a1, d.dump_a2, d.dump_softmax = d1.data_ptr(), d2.data_ptr(), d3.data_ptr()
    cabi.check(cabi.lib().p2v_window_attention_int(xq.data_ptr(), out.data_ptr(), images, C.byref(d), st))    # dump kernel
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d1.cpu().numpy(), w1)
    np.testing.assert_array_equal(d2.cpu().numpy(), w2)
    np.testing.assert_array_equal(d3.cpu().numpy(), wsm)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    np.testing.assert_array_equal(out2.cpu().numpy(), want)
    assert (wsm < 16).mean() > 0.02 and len(np.unique(want)) > 8


def test_swin_engine_scope_float_scale_grids_keep_the_per_module_path():
    """The engine covers symmetric power-of-two activation grids (the minmax observer of config 5).  A model
    calibrated with a float-scale observer is refused once (NotImplementedError inside, recorded on the model) and its
    forward keeps running on the per-module operators - same logits as with `per_module` forced."""
    import diff_vit_b200 as dv
    z = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
    a = [int(v) for v in z['arch']]
    model = dv.SwinTransformer(img_size=a[0], patch_size=a[1], num_classes=a[2], embed_dim=a[3], window_size=a[4],
                               depths=tuple(a[5:7]), num_heads=tuple(a[7:9]), norm_layer=dv.QIntLayerNorm,
                               input_quant=True, cfg=dv.Config(True, True, 'percentile')).eval()
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}, strict=True)
    model = model.cuda()
    dv.calibrate_model(model, [torch.from_numpy(z['x_calib']).cuda()])
    x = torch.from_numpy(z['x_eval']).cuda()
    with torch.no_grad():
        got = model(x)
    assert model._engine_off, 'a percentile-calibrated model must be outside the integer engine'
    model.per_module = True
    with torch.no_grad():
        want = model(x)
    assert torch.equal(got, want) and got.std() > 0


def test_swin_engine_from_a_serialised_plan(tmp_path):
    """A plan written by save_swin_plan runs on a fresh engine without the float model: same logits as the model's
    own engine, for the 8-bit and for a mixed 4 / 8-bit assignment (int4-packed layers in the file)."""
    import diff_vit_b200 as dv
    from diff_vit_b200.swin_engine import SwinIntegerEngine, load_swin_plan, save_swin_plan
    z = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
    model = build_swin_micro(z).cuda()
    dv.calibrate_model(model, [torch.from_numpy(z['x_calib']).cuda()])
    x = torch.from_numpy(z['x_eval']).cuda()
    n = model.num_linear_layers()
    eng = model.integer_engine()
    plans = []
    for i, bits in enumerate(([8] * n, [4 if j % 2 else 8 for j in range(n)])):
        _, bound = eng.bound(bits)
        path = str(tmp_path / ('plan%d.npz' % i))
        save_swin_plan(bound.plan, path)
        plans.append(load_swin_plan(path))
    fresh = SwinIntegerEngine(device='cuda', plans=plans)
    for bits in ([8] * n, [4 if j % 2 else 8 for j in range(n)]):
        assert torch.equal(fresh.forward(x, bits), eng.forward(x, bits))
    with pytest.raises(KeyError):
        fresh.forward(x, [4] * n)              # no such plan in the files and no calibrated state to build it from


def test_swin_uint8_pixel_entry_equals_fp32_entry_on_normalised_images():
    """p2v_swin_forward_u8 / SwinIntegerEngine.forward_u8: 8-bit pixels plus the loader's mean / std give the logits of
    the fp32 entry on torchvision's ToTensor + Normalize of the same pixels, bit for bit (all 256 pixel values occur in
    every channel), eager and through the replayed graph."""
    import diff_vit_b200 as dv
    z = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
    model = build_swin_micro(z).cuda()
    dv.calibrate_model(model, [torch.from_numpy(z['x_calib']).cuda()])
    eng = model.integer_engine()
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (5, 3, 56, 56), dtype=torch.uint8, generator=g)
    img[1].reshape(3, -1)[:, :256].copy_(torch.arange(256, dtype=torch.uint8).expand(3, 256))
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = img.float().div(255)
    x = (x - torch.tensor(mean).reshape(1, 3, 1, 1)) / torch.tensor(std).reshape(1, 3, 1, 1)     # ToTensor + Normalize
    want = eng.forward(x.cuda())
    got = eng.forward_u8(img.cuda(), mean, std)
    again = eng.forward_u8(img.cuda(), mean, std)
    assert torch.equal(got, want) and torch.equal(again, want) and want.std() > 0
