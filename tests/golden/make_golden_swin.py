#!/usr/bin/env python
"""Generate tests/golden/swin_micro.npz by running the reference's Swin graph (/root/reference, models/swin_quant.py)
on CPU in the build container.

The reference's Swin file is stale and raises TypeError as shipped: it calls
`self.patch_embed(x)` (swin_quant.py:793) and `self.mlp(x)` (swin_quant.py:392-396) although
`PatchEmbed.forward(x, FLOPs, bit_config)` / `Mlp.forward(x, FLOPs, global_distance, ffn_bit_config, ...)`
(layers_quant.py:202,416) now require more arguments.  On top of the two shims of make_golden.py (matplotlib stub,
Tensor.cuda = identity) this script therefore wraps exactly those two forwards so that the missing arguments default
to `FLOPs=[]`, `global_distance=[]` and 8-bit weights (`bit_config=8`, `ffn_bit_config=(8, 8)`: what every other
`QLinear(x)` call of the file resolves to after calibration).  A third defect shows once those run: `PatchMerging.reduction` is
built with `bias=False` (swin_quant.py:430) but the weight observer's scale search indexes the bias
(`self.others[0][j]`, observer/minmax.py:126) and dies on None.  The script gives every bias-less QLinear an
all-zero bias before calibration (x W^T + 0 is x W^T exactly; the extra `*.bias` keys are dropped from the stored
state dict).  A fourth: `PatchMerging.forward` calls
`self.norm(x, last_quantizer, self.qact1.quantizer, 4)` (swin_quant.py:463), written for upstream FQ-ViT's
`forward(x, in_quantizer, out_quantizer, in_scale_expand)`; the fork inserted `out_quantizer_scale` before
`in_scale_expand` (ptq/layers.py:240-245), so the 4 lands in the wrong parameter and the integer LayerNorm fails on
a shape mismatch.  The script routes a plain-int fourth argument to `in_scale_expand`, where the call meant it.
Call plumbing only - no arithmetic is touched: every tensor in the fixture is produced by the reference's own
modules.

Usage:  python tests/golden/make_golden_swin.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import torch, ref_config  # noqa: E402  (imports the reference with the two base shims)
from models import layers_quant as ref_lq  # noqa: E402
from models import swin_quant as ref_swin  # noqa: E402
from models.ptq.layers import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear  # noqa: E402

_pe_forward = ref_lq.PatchEmbed.forward
_mlp_forward = ref_lq.Mlp.forward


def _pe_shim(self, x, FLOPs=None, bit_config=8):
    return _pe_forward(self, x, [] if FLOPs is None else FLOPs, bit_config)


def _mlp_shim(self, x, FLOPs=None, global_distance=None, ffn_bit_config=(8, 8), *args, **kwargs):
    return _mlp_forward(self, x, [] if FLOPs is None else FLOPs, [] if global_distance is None else global_distance,
                        ffn_bit_config, *args, **kwargs)


ref_lq.PatchEmbed.forward = _pe_shim
ref_lq.Mlp.forward = _mlp_shim

_ln_forward = QIntLayerNorm.forward


def _ln_shim(self, x, in_quantizer=None, out_quantizer=None, out_quantizer_scale=None, in_scale_expand=1):
    if isinstance(out_quantizer_scale, int):       # shim 4 (see the header)
        out_quantizer_scale, in_scale_expand = None, out_quantizer_scale
    return _ln_forward(self, x, in_quantizer, out_quantizer, out_quantizer_scale, in_scale_expand)


QIntLayerNorm.forward = _ln_shim

ARCH = dict(img_size=56, patch_size=4, in_chans=3, num_classes=10, embed_dim=32, depths=(2, 2), num_heads=(1, 2),
            window_size=7)


def perturb(model, seed):
    """Non-degenerate parameters: non-zero biases, non-unit LayerNorm affine, a bias table that matters, wider qkv."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if 'relative_position_bias_table' in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif name.endswith('.bias') and 'norm' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif 'norm' in name and name.endswith('.weight'):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif 'norm' in name and name.endswith('.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith('attn.qkv.weight'):
                p.mul_(6.0)


def main():
    torch.manual_seed(0)
    cfg = ref_config.Config(True, True, 'minmax')
    model = ref_swin.SwinTransformer(norm_layer=QIntLayerNorm, input_quant=True, cfg=cfg, **ARCH).eval()
    perturb(model, 1)
    out = {'sd/' + k: v.detach().numpy() for k, v in model.state_dict().items()}
    for m in model.modules():          # shim 3 (see the header): a zero bias where the reference built none
        if isinstance(m, QLinear) and m.bias is None:
            m.bias = torch.nn.Parameter(torch.zeros(m.out_features))
    g = torch.Generator().manual_seed(2)
    x_calib = torch.randn(6, 3, 56, 56, generator=g)
    x_eval = torch.randn(4, 3, 56, 56, generator=g)
    out['x_calib'], out['x_eval'] = x_calib.numpy(), x_eval.numpy()
    # calibration exactly as the reference's flow drives it (test_quant.py:222-249)
    model.model_open_calibrate()
    with torch.no_grad():
        model.model_open_last_calibrate()
        model(x_calib)
    model.model_close_calibrate()
    model.model_quant()
    for name, m in model.named_modules():
        if isinstance(m, QAct) and m.quantizer.scale is not None:
            out['scale/' + name] = m.quantizer.scale.detach().numpy().astype(np.float32)
            out['zp/' + name] = m.quantizer.zero_point.detach().numpy().astype(np.int64)
        if isinstance(m, (QLinear, QConv2d)):
            for bit, s in m.quantizer.dic_scale.items():
                out['wscale/%s/%s' % (name, bit)] = s.detach().numpy().astype(np.float32)
        if isinstance(m, ref_lq.Mlp) and m.channel_scale is not None:
            out['cs/' + name] = m.best_scale[0].detach().numpy().astype(np.float32)
    store = {}

    def act_hook(name):
        def fn(mod, inp, outp):
            q = mod.quantizer
            s = q.scale.reshape(q.get_reshape_range(outp))
            z = q.zero_point.reshape(q.get_reshape_range(outp))
            store['act/' + name] = (outp / s + z).round().to(torch.int16).numpy()
        return fn

    def ln_hook(name):
        def fn(mod, inp, outp):
            if mod.mode == 'int':
                store['ln/' + name] = (outp / inp[2].scale.reshape(1, 1, -1)).round().to(torch.int32).numpy()
        return fn

    def sm_hook(name):
        def fn(mod, inp, outp):
            k = torch.where(outp > 0, -torch.log2(outp.clamp_min(1e-30)), torch.full_like(outp, 16.0))
            store['softmax/' + name] = k.round().to(torch.uint8).numpy()
        return fn

    for name, m in model.named_modules():
        if isinstance(m, QAct):
            m.register_forward_hook(act_hook(name))
        elif isinstance(m, QIntLayerNorm):
            m.register_forward_hook(ln_hook(name))
        elif isinstance(m, QIntSoftmax):
            m.register_forward_hook(sm_hook(name))
    with torch.no_grad():
        logits = model(x_eval)
    out['w8/logits'] = logits.numpy().astype(np.float32)
    for k, v in store.items():
        out['w8/' + k] = v
    out['arch'] = np.array([ARCH['img_size'], ARCH['patch_size'], ARCH['num_classes'], ARCH['embed_dim'],
                            ARCH['window_size']] + list(ARCH['depths']) + list(ARCH['num_heads']), dtype=np.int64)
    path = os.path.join(HERE, 'swin_micro.npz')
    np.savez_compressed(path, **out)
    sm = [v for k, v in store.items() if k.startswith('softmax/')]
    print('wrote %s: %d arrays, %.1f KB; logits %s; softmax codes %d..%d; layers with codes: %d' % (
        path, len(out), os.path.getsize(path) / 1e3, logits.shape, min(int(v.min()) for v in sm),
        max(int(v.max()) for v in sm), len(store)))


if __name__ == '__main__':
    main()
