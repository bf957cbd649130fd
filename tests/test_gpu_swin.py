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
    print('swin_tiny config 5 on the integer engine: %d codes compared on identical inputs, %d differ' % (total, bad))


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
