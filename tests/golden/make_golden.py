#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, LeSN-Lab/diff-ViT) on CPU in the build container.

The reference cannot travel to the GPU box, so its outputs are committed as .npz
fixtures together with this script.  Two import shims are needed and change call
plumbing only, never arithmetic:
  1. matplotlib stubs (models/plot_distrib.py:1-4 imports it unconditionally);
  2. torch.Tensor.cuda = identity (the reference hard-codes .cuda(),
     models/ptq/quantizer/uniform.py:85,126, models/ptq/observer/minmax.py:67-71).

Usage:  python tests/golden/make_golden.py [micro] [deit_tiny]
"""
import hashlib
import os
import sys
import types
from functools import partial

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec', 'matplotlib.collections'):
        sys.modules[n] = types.ModuleType(n)
    sys.modules['matplotlib.collections'].PolyCollection = object
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    import config as ref_config
    import models as ref_models
    return torch, ref_config, ref_models


torch, ref_config, ref_models = import_reference()
from models.ptq.layers import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear  # noqa: E402
from models.vit_fquant import Attention, VisionTransformer  # noqa: E402
from models.layers_quant import Mlp  # noqa: E402


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def perturb(model, seed):
    """Make the degenerate random-init paths non-trivial: non-zero biases,
    non-unit LayerNorm affine, wider qkv so that softmax codes span the whole range."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('.bias') and ('norm' not in name):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif 'norm' in name and name.endswith('.weight'):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif 'norm' in name and name.endswith('.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith('attn.qkv.weight'):
                p.mul_(4.0)


def calibrate(model, x):
    model.model_open_calibrate()
    with torch.no_grad():
        model.model_open_last_calibrate()
        _, flops, gd = model(x, plot=False)
    model.model_close_calibrate()
    model.model_quant()
    return flops, gd


def collect_scales(model, out):
    for name, m in model.named_modules():
        if isinstance(m, QAct) and m.quantizer.scale is not None:
            out['scale/' + name] = m.quantizer.scale.detach().numpy().astype(np.float32)
            out['zp/' + name] = m.quantizer.zero_point.detach().numpy().astype(np.int64)
        if isinstance(m, (QLinear, QConv2d)):
            for bit, s in m.quantizer.dic_scale.items():
                out['wscale/%s/%s' % (name, bit)] = s.detach().numpy().astype(np.float32)
        if isinstance(m, (Attention, Mlp)) and m.channel_scale is not None:
            out['cs/' + name] = m.best_scale[0].detach().numpy().astype(np.float32)
            for i in range(2):
                assert torch.equal(m.best_scale[i], m.best_scale[0])
                out['best_act_scale/%s/%d' % (name, i)] = m.best_act_scale[i].detach().numpy().astype(np.float32)


def run_eval(model, x, bit_config, prefix, out, keep=None, full_names=None):
    """Run the quantized forward with hooks; store integer codes of every QAct / int-LN / softmax.

    keep: number of leading images whose codes are stored in full (None = all).
    Every tensor additionally gets an int64 checksum pair over the whole batch."""
    store = {}

    def act_hook(name):
        def fn(mod, inp, outp):
            q = mod.quantizer
            s = q.scale.reshape(q.get_reshape_range(outp))
            z = q.zero_point.reshape(q.get_reshape_range(outp))
            store['act/' + name] = (outp / s + z).round().to(torch.int16)
        return fn

    def ln_hook(name):
        def fn(mod, inp, outp):
            if mod.mode != 'int':
                return
            out_q, out_cs = inp[2], (inp[3] if len(inp) > 3 else None)
            s = out_q.scale if out_cs is None else out_q.scale * out_cs
            store['ln/' + name] = (outp / s.reshape(1, 1, -1)).round().to(torch.int32)
        return fn

    def sm_hook(name):
        def fn(mod, inp, outp):
            k = torch.where(outp > 0, -torch.log2(outp.clamp_min(1e-30)), torch.full_like(outp, 16.0))
            store['softmax/' + name] = k.round().to(torch.uint8)
        return fn

    hs = []
    for name, m in model.named_modules():
        if isinstance(m, QAct):
            hs.append(m.register_forward_hook(act_hook(name)))
        elif isinstance(m, QIntLayerNorm):
            hs.append(m.register_forward_hook(ln_hook(name)))
        elif isinstance(m, QIntSoftmax):
            hs.append(m.register_forward_hook(sm_hook(name)))
    with torch.no_grad():
        logits, flops, gd = model(x, list(bit_config), False)
    for h in hs:
        h.remove()
    for name, m in model.named_modules():
        if isinstance(m, Attention):
            store['float/%s.qkv_output' % name] = m.qkv_output
        if isinstance(m, Mlp):
            store['float/%s.fc1_output' % name] = m.fc1_output
    out[prefix + '/logits'] = logits.numpy().astype(np.float32)
    out[prefix + '/flops'] = np.asarray(flops, dtype=np.int64)
    for k, v in store.items():
        if k.startswith('float/'):
            if keep is None:
                out['%s/%s' % (prefix, k)] = v.numpy().astype(np.float32)
            continue
        a = v.numpy()
        flat = a.astype(np.int64).reshape(-1)
        w = (np.arange(flat.size, dtype=np.int64) % 65521) + 1
        out['%s/sum/%s' % (prefix, k)] = np.asarray([flat.sum(), (flat * w).sum()], dtype=np.int64)
        if keep is None or (full_names is not None and any(k.endswith(n) for n in full_names)):
            a = a if keep is None else a[:keep]
            if a.min() >= -128 and a.max() <= 127:
                a = a.astype(np.int8)
            out['%s/%s' % (prefix, k)] = a


def make_micro():
    """Tiny 2-block / 2-head ViT (hd=64 so that the attention scale is the exact power of two 2^-3,
    like every DeiT/ViT config of the reference) with perturbed parameters.  Everything is stored."""
    for variant, seed in (('micro_minmax', 0),):
        torch.manual_seed(seed)
        cfg = ref_config.Config(True, True, 'minmax')
        model = VisionTransformer(img_size=48, patch_size=16, embed_dim=128, depth=2, num_heads=2,
                                  mlp_ratio=4, qkv_bias=True, norm_layer=partial(QIntLayerNorm, eps=1e-6),
                                  input_quant=True, cfg=cfg, num_classes=16).eval()
        perturb(model, seed + 100)
        out = {}
        for k, v in model.state_dict().items():
            out['sd/' + k] = v.numpy().copy()
        out['sd_hash'] = np.asarray(sd_hash(model.state_dict()))
        g = torch.Generator().manual_seed(seed + 1)
        x_cal = torch.randn(8, 3, 48, 48, generator=g)
        x_eval = torch.randn(6, 3, 48, 48, generator=g) * 1.3
        out['x_calib'] = x_cal.numpy()
        out['x_eval'] = x_eval.numpy()
        flops, gd = calibrate(model, x_cal)
        out['calib/flops'] = np.asarray(flops, dtype=np.int64)
        out['calib/global_distance'] = np.asarray([[float(d) for d in row] for row in gd], dtype=np.float32)
        collect_scales(model, out)
        nl = 4 * 2 + 2
        run_eval(model, x_eval, [8] * nl, 'w8', out)
        run_eval(model, x_eval, [4] * nl, 'w4', out)
        mixed = [4] * nl
        for i in (0, 2, 5, 9):
            mixed[i] = 8
        out['mixed/bit_config'] = np.asarray(mixed)
        run_eval(model, x_eval, mixed, 'mixed', out)
        path = os.path.join(HERE, variant + '.npz')
        np.savez_compressed(path, **out)
        print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_deit_tiny():
    """Config C1 (BASELINE.json configs[0]): deit_tiny, minmax, W8A8 PoT, randn(32,3,224,224) as both the
    calibration and the eval batch.  Weights are the factory's random init under torch.manual_seed(0); only
    their hash is stored (the package's factories reproduce them bit-for-bit from the seed)."""
    torch.manual_seed(0)
    cfg = ref_config.Config(True, True, 'minmax')
    model = ref_models.deit_tiny_patch16_224(pretrained=False, cfg=cfg).eval()
    out = {'sd_hash': np.asarray(sd_hash(model.state_dict()))}
    x = torch.randn(32, 3, 224, 224)
    out['x_hash'] = np.asarray(hashlib.sha256(x.numpy().tobytes()).hexdigest())
    flops, gd = calibrate(model, x)
    out['calib/flops'] = np.asarray(flops, dtype=np.int64)
    out['calib/global_distance'] = np.asarray([[float(d) for d in row] for row in gd], dtype=np.float32)
    collect_scales(model, out)
    full = ['blocks.0.attn.qact0', 'blocks.0.attn.qact1', 'blocks.0.attn.qact_attn1', 'blocks.0.attn.log_int_softmax',
            'blocks.0.attn.qact2', 'blocks.0.attn.qact3', 'blocks.0.qact2', 'blocks.0.mlp.qact0', 'blocks.0.mlp.qact1',
            'blocks.0.mlp.qact2', 'blocks.0.qact4', 'blocks.11.qact4', 'blocks.11.mlp.qact1', 'act/qact1',
            'patch_embed.qact', 'act/qact2', 'act_out']
    run_eval(model, x, [8] * 50, 'w8', out, keep=2, full_names=full)
    run_eval(model, x, [4] * 50, 'w4', out, keep=2, full_names=['act_out', 'blocks.11.qact4'])
    path = os.path.join(HERE, 'deit_tiny_c1.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    what = sys.argv[1:] or ['micro', 'deit_tiny']
    if 'micro' in what:
        make_micro()
    if 'deit_tiny' in what:
        make_deit_tiny()
