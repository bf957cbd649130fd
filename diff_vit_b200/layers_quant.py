"""Shared blocks of the quantized ViT graph: SmoothQuant calibration, Mlp, PatchEmbed
(reference: models/layers_quant.py:141-492).
"""
import collections.abc
from itertools import repeat

import torch
import torch.nn.functional as F
from torch import nn

from .ptq import QAct, QConv2d, QLinear
from .ptq.observer.utils import ln2_round

# SmoothQuant migration strength candidates for fc1 and the weight bit widths the calibrated
# state is indexed by (reference: models/layers_quant.py:13-15).
alpha_pool = [0.5]
bit_pool = [4, 8]


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable):
        return x
    return tuple(repeat(x, 2))


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Truncated normal init; the reference carries a private copy of the algorithm that became
    torch.nn.init.trunc_normal_ (models/layers_quant.py:51-104), which consumes the RNG identically."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def drop_path(x, drop_prob: float = 0.0, training: bool = False):
    if drop_prob == 0.0 or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    mask = (keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
    return x.div(keep_prob) * mask


class DropPath(nn.Module):
    """Stochastic depth; identity at inference."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


class SmoothQuantState:
    """Calibrated SmoothQuant state of one (QAct -> QLinear) pair, shared by Attention and Mlp.

    The reference keeps five parallel ``best_*`` lists on the owner module, indexed by
    ``bit_pool.index(bit)`` (models/vit_fquant.py:229-280, models/layers_quant.py:250-302).  With a
    single alpha in the pool the "search" is degenerate and both entries alias the same objects;
    the candidate losses it computes therefore never influence the result and are not evaluated here.
    """

    FIELDS = ('channel_scale', 'best_scale', 'best_act_scale', 'best_act_zp', 'best_weight_scale', 'best_weight_zp')

    @staticmethod
    def move(owner, fn):
        """Carry the owner's SmoothQuant state through a Module._apply (model.cuda(), model.to(...))."""
        from .ptq.quantizer.core import map_state
        for name in SmoothQuantState.FIELDS:
            if getattr(owner, name, None) is not None:
                setattr(owner, name, map_state(fn, getattr(owner, name)))

    @staticmethod
    def channel_scale(x, weight, alpha):
        """PoT-rounded max|x|_c^alpha / max|W|_c^(1-alpha) (per input channel)."""
        global_max_x = torch.abs(x).max(axis=1).values.max(axis=0).values
        from . import dist as _dist
        _dist.reduce_max_(global_max_x)     # activations are sharded over ranks while calibrating
        max_weight = torch.abs(weight).max(axis=0).values
        cs = global_max_x ** alpha / (max_weight ** (1 - alpha))
        return 2 ** ln2_round(cs)

    @staticmethod
    def calibrate(owner, x, qact0, linear, pool, bit_config, global_distance, **linear_kw):
        """Calibration-time forward of the smoothed pair; returns (float output, smoothed input)."""
        if owner.channel_scale is None:
            owner.best_scale, owner.best_act_scale, owner.best_act_zp = [], [], []
            owner.best_weight_scale, owner.best_weight_zp = [], []
        found = None
        for alpha in pool:
            cs = SmoothQuantState.channel_scale(x, linear.weight, alpha)
            x_smoothed = x / cs.reshape((1, 1, -1))
            weight_smoothed = linear.weight * cs.reshape((1, -1))
            gt = F.linear(x_smoothed, weight_smoothed, linear.bias)
            qact0(x_smoothed)
            if qact0.last_calibrate and bit_config != -1:
                act = (qact0.quantizer.scale, qact0.quantizer.zero_point)
                linear(x_smoothed, global_distance, bit_config, weight_smoothed, **linear_kw)
                found = (cs, act, linear.quantizer.dic_scale, linear.quantizer.dic_zero_point)
                # the reference leaves the pair on the last pool entry after scoring the candidates
                linear.quantizer.bit_type = linear.observer.bit_type = \
                    linear._bit_type_of(bit_pool[-1])
        if found is not None:
            cs, act, w_scale, w_zp = found
            for _ in bit_pool:
                owner.channel_scale = cs
                owner.best_scale.append(cs)
                owner.best_act_scale.append(act[0])
                owner.best_act_zp.append(act[1])
                owner.best_weight_scale.append(w_scale)
                owner.best_weight_zp.append(w_zp)
        return gt, x_smoothed

    @staticmethod
    def select(owner, qact0, linear, bit_config):
        """Quantized-mode lookup; ``bit_config=None`` raises ValueError exactly like the reference's
        ``bit_pool.index(None)``."""
        indx = bit_pool.index(bit_config)
        owner.channel_scale = owner.best_scale[indx]
        qact0.quantizer.scale = owner.best_act_scale[indx]
        qact0.quantizer.zero_point = owner.best_act_zp[indx]
        linear.quantizer.dic_scale = owner.best_weight_scale[indx]
        linear.quantizer.dic_zero_point = owner.best_weight_zp[indx]
        return owner.channel_scale


class Mlp(nn.Module):

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0,
                 quant=False, calibrate=False, cfg=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        act = dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_A,
                   calibration_mode=cfg.CALIBRATION_MODE_A, observer_str=cfg.OBSERVER_A,
                   quantizer_str=cfg.QUANTIZER_A)
        act_ln = dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_A,
                      calibration_mode=cfg.CALIBRATION_MODE_A_LN, observer_str=cfg.OBSERVER_A_LN,
                      quantizer_str=cfg.QUANTIZER_A_LN)
        wgt = dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_W,
                   calibration_mode=cfg.CALIBRATION_MODE_W, observer_str=cfg.OBSERVER_W,
                   quantizer_str=cfg.QUANTIZER_W)
        self.qact0 = QAct(**act)
        self.fc1 = QLinear(in_features, hidden_features, **wgt)
        self.act = act_layer()
        self.qact1 = QAct(**act)
        self.fc2 = QLinear(hidden_features, out_features, **wgt)
        self.qact2 = QAct(**act_ln)
        self.drop = nn.Dropout(drop)
        self.channel_scale = None
        self.fc1_output = None

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        SmoothQuantState.move(self, fn)
        return self

    def forward(self, x, FLOPs, global_distance, ffn_bit_config, plot=False, quant=True, smoothquant=True,
                activation=[], hessian_statistic=False):
        B, N, C = x.shape
        bit_config = ffn_bit_config[0] if ffn_bit_config else None
        if smoothquant and not hessian_statistic:
            if self.channel_scale is None or bit_config == -1:
                x, _ = SmoothQuantState.calibrate(self, x, self.qact0, self.fc1, alpha_pool, bit_config,
                                                  global_distance)
            else:
                cs = SmoothQuantState.select(self, self.qact0, self.fc1, bit_config)
                weight_smoothed = self.fc1.weight * cs.reshape((1, -1))
                x = self.qact0(x / cs.reshape((1, 1, -1)))
                x = self.fc1(x, global_distance, bit_config, weight_smoothed)
        else:
            x = self.fc1(self.qact0(x), global_distance, bit_config, None)
        self.fc1_output = x.detach().clone()
        FLOPs.append(N * C * x.shape[2])
        x = self.act(x)
        x = self.qact1(x, asymmetric=False)
        x = self.drop(x)
        B, N, C = x.shape
        bit_config = ffn_bit_config[1] if ffn_bit_config else None
        x = self.fc2(x, global_distance, bit_config)
        FLOPs.append(N * C * x.shape[2])
        x = self.qact2(x)
        return self.drop(x)


class PatchEmbed(nn.Module):
    """Image to patch embedding: stride-P QConv2d, flatten to tokens, QAct."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        act = dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_A,
                   calibration_mode=cfg.CALIBRATION_MODE_A, observer_str=cfg.OBSERVER_A,
                   quantizer_str=cfg.QUANTIZER_A)
        self.proj = QConv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, quant=quant,
                            calibrate=calibrate, bit_type=cfg.BIT_TYPE_W,
                            calibration_mode=cfg.CALIBRATION_MODE_W, observer_str=cfg.OBSERVER_W,
                            quantizer_str=cfg.QUANTIZER_W)
        if norm_layer:
            self.qact_before_norm = QAct(**act)
            self.norm = norm_layer(embed_dim)
        else:
            self.qact_before_norm = nn.Identity()
            self.norm = nn.Identity()
        self.qact = QAct(**act)

    def forward(self, x, FLOPs, bit_config):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = self.proj(x, bit_config)
        B, M, H, W = x.shape
        FLOPs.append(C * self.patch_size[0] * self.patch_size[0] * M * H * W)
        x = x.flatten(2).transpose(1, 2)
        x = self.qact_before_norm(x)
        if isinstance(self.norm, nn.Identity):
            x = self.norm(x)
        else:
            x = self.norm(x, self.qact_before_norm.quantizer, self.qact.quantizer)
        return self.qact(x)
