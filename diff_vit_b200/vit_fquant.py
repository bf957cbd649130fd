"""Quantized ViT / DeiT graph (reference: models/vit_fquant.py:57-933).

Same modules, attribute names, state-dict keys, calibration switches and
``forward(x, bit_config, plot, hessian_statistic) -> (logits, FLOPs, global_distance)`` contract as
the reference.  Float / calibration passes run the graph below in PyTorch.  After
``model_quant()`` the forward is handed to the integer engine: one pre-extracted plan of int8/int4
weight codes, power-of-two exponents and per-channel PTF scales executed by the sm_100a kernels
behind the C-ABI (``diff_vit_b200.engine``).
"""
from collections import OrderedDict
from functools import partial

import torch
from torch import nn
from torch.nn.modules import module as _module_globals

from .layers_quant import DropPath, Mlp, PatchEmbed, SmoothQuantState, trunc_normal_
from .ptq import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear

__all__ = [
    'deit_tiny_patch16_224', 'deit_small_patch16_224', 'deit_base_patch16_224',
    'vit_base_patch16_224', 'vit_large_patch16_224'
]

# SmoothQuant migration strength for qkv (reference: models/vit_fquant.py:32-33)
alpha_pool = [0.35]
bit_pool = [4, 8]


def _act_kw(cfg, quant, calibrate, ln=False):
    return dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_A,
                calibration_mode=cfg.CALIBRATION_MODE_A_LN if ln else cfg.CALIBRATION_MODE_A,
                observer_str=cfg.OBSERVER_A_LN if ln else cfg.OBSERVER_A,
                quantizer_str=cfg.QUANTIZER_A_LN if ln else cfg.QUANTIZER_A)


def _weight_kw(cfg, quant, calibrate):
    return dict(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_W,
                calibration_mode=cfg.CALIBRATION_MODE_W, observer_str=cfg.OBSERVER_W,
                quantizer_str=cfg.QUANTIZER_W)


class Attention(nn.Module):

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.calibrate = calibrate
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = QLinear(dim, dim * 3, bias=qkv_bias, **_weight_kw(cfg, quant, calibrate))
        self.qact0 = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate))
        self.proj = QLinear(dim, dim, **_weight_kw(cfg, quant, calibrate))
        self.qact3 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))
        self.qact_attn1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.log_int_softmax = QIntSoftmax(log_i_softmax=cfg.INT_SOFTMAX, quant=quant, calibrate=calibrate,
                                           bit_type=cfg.BIT_TYPE_S, calibration_mode=cfg.CALIBRATION_MODE_S,
                                           observer_str=cfg.OBSERVER_S, quantizer_str=cfg.QUANTIZER_S)
        self.channel_scale = None
        self.qkv_output = None

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        SmoothQuantState.move(self, fn)
        return self

    def forward(self, x, FLOPs, global_distance, atten_bit_config, plot=False, quant=False, smoothquant=True,
                hessian_statistic=False):
        self.atten_bit_config = atten_bit_config
        B, N, C = x.shape
        bit_config = atten_bit_config[0] if atten_bit_config else None
        attn_para = [self.num_heads, C, self.scale]
        if smoothquant and not hessian_statistic:
            if self.channel_scale is None or bit_config == -1:
                x, _ = SmoothQuantState.calibrate(self, x, self.qact0, self.qkv, alpha_pool, bit_config,
                                                  global_distance, attn=False, attn_para=attn_para)
            else:
                cs = SmoothQuantState.select(self, self.qact0, self.qkv, bit_config)
                weight_smoothed = self.qkv.weight * cs.reshape((1, -1))
                x = self.qact0(x / cs.reshape((1, 1, -1)))
                x = self.qkv(x, global_distance, bit_config, weight_smoothed, attn=False, attn_para=attn_para)
        else:
            x = self.qkv(self.qact0(x), global_distance, bit_config, None, attn=False, attn_para=attn_para)
        self.qkv_output = x.detach().clone()
        FLOPs.append(N * C * x.shape[2])
        x = self.qact1(x, attn=False, attn_para=attn_para)
        qkv = x.reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = self.qact_attn1(attn)
        attn = self.log_int_softmax(attn, self.qact_attn1.quantizer.scale)
        attn = self.attn_drop(attn)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        x = self.qact2(x)
        bit_config = atten_bit_config[1] if atten_bit_config else None
        x = self.proj(x, global_distance, bit_config)
        FLOPs.append(N * C * x.shape[2])
        x = self.qact3(x)
        return self.proj_drop(x)

    def get_requant_scale(self):
        bits = 'int' + str(self.atten_bit_config[1])
        return (self.qact2.quantizer.scale * self.proj.quantizer.dic_scale[bits]) / self.qact3.quantizer.scale


class Block(nn.Module):

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, quant=False, calibrate=False,
                 cfg=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        # The reference builds Attention without forwarding quant/calibrate (vit_fquant.py:379-385).
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                              attn_drop=attn_drop, proj_drop=drop, cfg=cfg)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop,
                       quant=quant, calibrate=calibrate, cfg=cfg)
        self.qact4 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))

    def forward(self, x, last_quantizer=None, FLOPs=[], global_distance=[], local_bit_config=None, plot=False,
                quant=False, hessian_statistic=False):
        atten_bit_config = local_bit_config[0:2] if local_bit_config else None
        if atten_bit_config is not None and -1 in atten_bit_config:
            self.norm1.mode = 'ln'
        y = self.norm1(x, last_quantizer, self.attn.qact0.quantizer, self.attn.channel_scale)
        y = self.attn(y, FLOPs, global_distance, atten_bit_config, plot=False, quant=quant,
                      hessian_statistic=hessian_statistic)
        x = self.qact2(x + self.drop_path(y))
        ffn_bit_config = local_bit_config[2:4] if local_bit_config else None
        if ffn_bit_config is not None and -1 in ffn_bit_config:
            self.norm2.mode = 'ln'
        # norm2 is handed the ATTENTION block's SmoothQuant scale, as in the reference (vit_fquant.py:464);
        # Mlp then re-grids the result by cs_attn / cs_mlp.
        y = self.norm2(x, self.qact2.quantizer, self.mlp.qact0.quantizer, self.attn.channel_scale)
        y = self.mlp(y, FLOPs, global_distance, ffn_bit_config, plot, quant, activation=[],
                     hessian_statistic=hessian_statistic)
        return self.qact4(x + self.drop_path(y))


class VisionTransformer(nn.Module):

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, hybrid_backbone=None, norm_layer=None,
                 quant=False, calibrate=False, input_quant=False, cfg=None):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError('HybridEmbed is unquantized and unused by the reference factories')
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.mlp_ratio = mlp_ratio
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.cfg = cfg
        self.quant = False
        self.input_quant = input_quant
        if input_quant:
            self.qact_input = QAct(**_act_kw(cfg, quant, calibrate))
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                      embed_dim=embed_dim, quant=quant, calibrate=calibrate, cfg=cfg)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.qact_embed = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact_pos = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact1 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, quant=quant,
                  calibrate=calibrate, cfg=cfg) for i in range(depth)
        ])
        self.depth = depth
        self.norm = norm_layer(embed_dim)
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate))
        if representation_size:
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([('fc', nn.Linear(embed_dim, representation_size)),
                                                         ('act', nn.Tanh())]))
        else:
            self.pre_logits = nn.Identity()
        self.head = (QLinear(self.num_features, num_classes, **_weight_kw(cfg, quant, calibrate))
                     if num_classes > 0 else nn.Identity())
        self.act_out = QAct(**_act_kw(cfg, quant, calibrate))
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)
        self._engine = None
        self._submodules = None
        self._int_active = False  # model_quant() arms the fused integer engine, model_dequant() disarms it
        self.per_module = False   # True: run every Q-module's own forward even when the fused engine could
        # a loaded checkpoint invalidates the plan the engine extracted from the old parameters
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_engine())

    def _drop_engine(self):
        """The engine's plan is a snapshot of the calibrated state (weights, scales, device): anything that changes
        one of them drops it; the next quantized forward rebuilds it."""
        self._engine = None
        self._submodules = None

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda() / .float(): parameters move or change
        super()._apply(fn, *args, **kwargs)
        self._drop_engine()
        return self

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self._drop_engine()

    # -- mode switches (reference: vit_fquant.py:667-698) -------------------------------------
    _Q_MODULES = (QConv2d, QLinear, QAct, QIntSoftmax)

    def _set_flag(self, name, value):
        for m in self.modules():
            if type(m) in self._Q_MODULES:
                setattr(m, name, value)

    def model_quant(self, flag='on'):
        if flag == 'on':
            self.quant = True
        self._set_flag('quant', True)
        if self.cfg.INT_NORM and flag != 'off':
            for m in self.modules():
                if type(m) is QIntLayerNorm:
                    m.mode = 'int'
        self._int_active = True
        self._drop_engine()

    def model_dequant(self):
        """The reference clears the modules' flags and leaves the model-level `quant` set (vit_fquant.py:680-683): the
        next forward runs the float modules.  So does this one: the fused engine stays off until model_quant()."""
        self._set_flag('quant', False)
        self._int_active = False
        self._drop_engine()

    def model_open_calibrate(self):
        self._set_flag('calibrate', True)
        self._drop_engine()       # scales are about to change

    def model_open_last_calibrate(self):
        self._set_flag('last_calibrate', True)

    def model_close_calibrate(self):
        self._set_flag('calibrate', False)

    # -- graph ---------------------------------------------------------------------------------
    def forward_features(self, x, FLOPs, global_distance, bit_config, global_plot, hessian_statistic=False):
        B = x.shape[0]
        if self.input_quant:
            x = self.qact_input(x)
        patch_bit = bit_config[0] if bit_config else None
        x = self.patch_embed(x, FLOPs, patch_bit)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        x = self.qact_embed(x)
        x = x + self.qact_pos(self.pos_embed)
        x = self.qact1(x)
        x = self.pos_drop(x)
        for i, blk in enumerate(self.blocks):
            local_bit_config = bit_config[i * 4 + 1:i * 4 + 5] if bit_config else None
            last_quantizer = self.qact1.quantizer if i == 0 else self.blocks[i - 1].qact4.quantizer
            x = blk(x, last_quantizer, FLOPs, global_distance, local_bit_config, False, self.quant,
                    hessian_statistic)
        x = self.norm(x, self.blocks[-1].qact4.quantizer, self.qact2.quantizer)[:, 0]
        x = self.qact2(x)
        return self.pre_logits(x)

    def flops(self):
        """The per-layer MAC list the reference accumulates during a forward (len 4*depth+2)."""
        pe = self.patch_embed
        gh, gw = pe.grid_size
        n = pe.num_patches + 1
        d, hid = self.embed_dim, self.blocks[0].mlp.fc1.out_features
        out = [pe.proj.in_channels * pe.patch_size[0] * pe.patch_size[0] * d * gh * gw]
        for _ in range(self.depth):
            out += [n * d * 3 * d, n * d * d, n * d * hid, n * hid * d]
        out.append(self.num_features * self.num_classes)
        return out

    def integer_engine(self):
        """The sm_100a execution engine bound to this model's calibrated state (built lazily)."""
        if self._engine is None:
            from .engine import IntegerEngine
            self._engine = IntegerEngine(self)
        return self._engine

    def _hooked(self):
        """True when somebody listens on a submodule (forward hooks of cka_utility.py:39-66 /
        modeldiff_p2.py:50-82) or asked for per-module execution: those callers need every Q-module's own
        forward to run, which the fused engine skips."""
        if self.per_module:
            return True
        if self._submodules is None:
            self._submodules = [m for m in self.modules() if m is not self]
        if _module_globals._global_forward_hooks or _module_globals._global_forward_pre_hooks:
            return True
        return any(m._forward_hooks or m._forward_pre_hooks for m in self._submodules)

    def forward(self, x, bit_config=None, plot=False, hessian_statistic=False):
        if (self.quant and self._int_active and not hessian_statistic and self._integer_path(bit_config)
                and not self._hooked()):
            logits = self.integer_engine().forward(x, bit_config)
            return logits, self.flops(), []
        FLOPs, global_distance = [], []
        x = self.forward_features(x, FLOPs, global_distance, bit_config, plot, hessian_statistic)
        B, C = x.shape
        head_bit = bit_config[-1] if bit_config else None
        x = self.head(x, global_distance, head_bit)
        FLOPs.append(C * x.shape[1])
        x = self.act_out(x)
        return x, FLOPs, global_distance

    def _integer_path(self, bit_config):
        """Whole-graph integer execution needs every layer quantized: a -1 entry (that layer in fp32,
        layers.py:144) keeps the per-module path."""
        if bit_config is None:
            bit_pool.index(None)  # ValueError, as the reference's Attention.forward does
        return all(b != -1 for b in bit_config)


def _factory(short_name, embed_dim, depth, num_heads, input_quant=True):
    def build(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
        """pretrained: False, True (the reference's checkpoint, taken from the torch hub cache: nothing is downloaded)
        or the path of a .pth / .npz checkpoint (weights.load_pretrained)."""
        model = VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                                  mlp_ratio=4, qkv_bias=True, norm_layer=partial(QIntLayerNorm, eps=1e-6),
                                  quant=quant, calibrate=calibrate, input_quant=input_quant, cfg=cfg, **kwargs)
        if pretrained:
            from .weights import load_pretrained
            load_pretrained(model, short_name, pretrained)
        return model
    return build


# reference: models/vit_fquant.py:802-933
deit_tiny_patch16_224 = _factory('deit_tiny', 192, 12, 3)
deit_small_patch16_224 = _factory('deit_small', 384, 12, 6)
deit_base_patch16_224 = _factory('deit_base', 768, 12, 12)
vit_base_patch16_224 = _factory('vit_base', 768, 12, 12)
vit_large_patch16_224 = _factory('vit_large', 1024, 24, 16, input_quant=False)
for _n in __all__:
    globals()[_n].__name__ = _n
