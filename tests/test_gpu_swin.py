"""BASELINE config 5 on the GPU: the quantized Swin forward through the package's Q-modules, each running its own
sm_100a operator (fake-quant, integer LayerNorm, log-int-softmax kernels; fp32 library GEMMs for the products, as the
reference itself does), against the reference's golden codes and against the CPU oracle."""
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


def test_swin_tiny_batch128_config5_vs_oracle():
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
