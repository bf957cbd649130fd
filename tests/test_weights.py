"""Checkpoint entry points (diff_vit_b200/weights.py): the Flax .npz layout round-trips, `pretrained=` takes local
files and never downloads, and `build_model` answers to the timm names the reference calls it with
(reference: models/utils.py:11-197, models/vit_fquant.py:820-932, utils/build_model.py:64-93)."""
import os
from functools import partial

import numpy as np
import pytest
import torch

import diff_vit_b200 as dv
from diff_vit_b200 import weights
from diff_vit_b200.vit_fquant import VisionTransformer

CFG = dv.Config(True, True, 'minmax')


def _micro(seed, **kw):
    torch.manual_seed(seed)
    m = VisionTransformer(img_size=32, patch_size=8, embed_dim=32, depth=2, num_heads=2, num_classes=10, cfg=CFG, norm_layer=partial(dv.QIntLayerNorm, eps=1e-6), **kw)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn_like(p) * 0.1)
    return m.eval()


def _float_state(m):
    return {k: v for k, v in m.state_dict().items() if v.dtype.is_floating_point and v.numel() > 1 and
            any(k.endswith(s) for s in ('weight', 'bias', 'cls_token', 'pos_embed'))}


def test_npz_round_trip_restores_every_weight(tmp_path):
    src, dst = _micro(0), _micro(1)
    path = weights.export_weights_to_npz(src, str(tmp_path / 'micro.npz'))
    z = np.load(path)
    # the layout of Google's Flax checkpoints: HWIO stem, [in, heads, head_dim] attention kernels
    assert z['embedding/kernel'].shape == (8, 8, 3, 32)
    assert z['Transformer/encoderblock_1/MultiHeadDotProductAttention_1/key/kernel'].shape == (32, 2, 16)
    assert z['Transformer/encoderblock_0/MultiHeadDotProductAttention_1/out/kernel'].shape == (2, 16, 32)
    assert z['Transformer/encoderblock_0/MlpBlock_3/Dense_0/kernel'].shape == (32, 128)
    dv.load_weights_from_npz(dst, path)
    a, b = _float_state(src), _float_state(dst)
    assert a.keys() == b.keys() and len(a) > 20
    for k in a:
        assert torch.equal(a[k], b[k]), k
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        assert torch.equal(src(x)[0], dst(x)[0])      # forward returns (logits, flops, distances) in float mode


def test_npz_prefix_and_position_embedding_resize(tmp_path):
    src = _micro(2)
    path = weights.export_weights_to_npz(src, str(tmp_path / 'opt.npz'), prefix='opt/target/')
    dst = _micro(3)
    weights.load_weights_from_npz(dst, path)             # prefix detected from the keys
    assert torch.equal(dst.blocks[1].mlp.fc2.weight, src.blocks[1].mlp.fc2.weight)
    big = VisionTransformer(img_size=64, patch_size=8, embed_dim=32, depth=2, num_heads=2, num_classes=10,
                            cfg=CFG, norm_layer=partial(dv.QIntLayerNorm, eps=1e-6)).eval()
    weights.load_weights_from_npz(big, path)             # 4 x 4 grid -> 8 x 8, bicubic, class token kept
    assert big.pos_embed.shape == (1, 65, 32)
    assert torch.equal(big.pos_embed[:, 0], src.pos_embed[:, 0])
    assert torch.isfinite(big.pos_embed).all()


def test_pretrained_takes_local_files_and_never_downloads(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.hub, 'get_dir', lambda: str(tmp_path))
    with pytest.raises(RuntimeError, match='never downloads'):
        dv.deit_tiny_patch16_224(pretrained=True, cfg=CFG)
    with pytest.raises(RuntimeError, match='does not exist'):
        dv.deit_tiny_patch16_224(pretrained=str(tmp_path / 'nope.pth'), cfg=CFG)
    # the reference's DeiT checkpoints are {'model': state_dict}; put one where torch.hub would have cached it
    torch.manual_seed(4)
    src = dv.deit_tiny_patch16_224(cfg=CFG)
    os.makedirs(tmp_path / 'checkpoints')
    torch.save({'model': {k: v for k, v in src.state_dict().items() if 'quantizer' not in k and 'observer' not in k}},
               tmp_path / 'checkpoints' / weights.CHECKPOINTS['deit_tiny'])
    dst = dv.deit_tiny_patch16_224(pretrained=True, cfg=CFG)
    assert torch.equal(dst.blocks[11].attn.qkv.weight, src.blocks[11].attn.qkv.weight)
    assert torch.equal(dst.head.bias, src.head.bias)


def test_build_model_alias(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.hub, 'get_dir', lambda: str(tmp_path))
    net = dv.build_model('deit_tiny_patch16_224', Pretrained=False)
    assert isinstance(net, VisionTransformer) and not net.training
    assert net.blocks[0].attn.num_heads == 3 and net.embed_dim == 192
    assert type(dv.build_model('swin_tiny_patch4_window7_224', Pretrained=False)).__name__ == 'SwinTransformer'
    with pytest.raises(ValueError, match='no model named'):
        dv.build_model('resnet50', Pretrained=False)
    with pytest.raises(RuntimeError, match='never downloads'):
        dv.build_model('deit_small_patch16_224')         # Pretrained=True is the reference's default


def test_model_dequant_clears_what_the_reference_clears(micro_golden):
    """model_dequant() (reference: models/vit_fquant.py:680-683) clears `quant` on QConv2d / QLinear / QAct /
    QIntSoftmax and nothing else: QIntLayerNorm stays in 'int' mode, as in the reference.  Here it must also switch the
    fused engine off, and model_quant() must bring everything back."""
    from conftest import build_micro
    model = build_micro(micro_golden)
    dv.calibrate_model(model, [torch.from_numpy(micro_golden['x_calib'])])
    flagged = [m for m in model.modules() if isinstance(m, (dv.QConv2d, dv.QLinear, dv.QAct, dv.QIntSoftmax))]
    norms = [m for m in model.modules() if isinstance(m, dv.QIntLayerNorm)]
    assert model._int_active and all(m.quant for m in flagged) and all(m.mode == 'int' for m in norms)
    model.model_dequant()
    assert not model._int_active and model._engine is None
    assert not any(m.quant for m in flagged)
    assert all(m.mode == 'int' for m in norms)
    model.model_quant()
    assert model._int_active and all(m.quant for m in flagged)
