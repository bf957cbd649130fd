"""Quantized Swin Transformer graph (reference: models/swin_quant.py:19-901, BASELINE config 5).

Same modules, attribute names and state-dict keys as the reference.  The reference file is stale: it calls
``self.patch_embed(x)`` (swin_quant.py:793) and ``self.mlp(x)`` (swin_quant.py:392-396) with signatures that
``PatchEmbed.forward(x, FLOPs, bit_config)`` / ``Mlp.forward(x, FLOPs, global_distance, ffn_bit_config, ...)``
(models/layers_quant.py:202,416) no longer have, and raises TypeError as shipped.  This graph is the same arithmetic
with the current signatures: every weight quantizer at 8 bits unless a ``bit_config`` says otherwise (what the
reference's ``QLinear(x)`` calls resolve to after calibration: the last calibrated bit type, int8), Mlp with its
SmoothQuant pair, FLOPs / weight distances collected like the ViT graph does.

``forward(x)`` returns the logits (the reference signature); ``forward(x, bit_config, plot)`` returns
``(logits, FLOPs, global_distance)`` like the ViT models, so the calibration / evaluation drivers work on both
families.  bit_config layout: [patch embed] + per stage (per block [qkv, proj, fc1, fc2] ... + [reduction] where the
stage downsamples) + [head].

After ``model_quant()`` the model-level forward of a CUDA tensor runs on the integer engine
(diff_vit_b200/swin_engine.py): int8 codes from the patch embedding to the head, tcgen05 GEMMs with fused
re-quantisation / GELU / residual epilogues, the integer LayerNorm and the window attention kernel (49 tokens, head
dimension 32, bias table and shift mask between the two re-quantisations; csrc/p2v_swin.cu).  With a forward hook on a
submodule, ``per_module = True``, a ``-1`` in ``bit_config`` or a configuration the engine does not cover (float
scales, zero points) every Q-module runs its own sm_100a operator instead (diff_vit_b200/standalone.py: fake-quant,
integer LayerNorm, log-int-softmax kernels; fp32 library GEMMs for the products, as the reference itself computes
them) - the analysis path, an order of magnitude slower.
"""
import torch
from torch import nn

from torch.nn.modules import module as _module_globals

from .layers_quant import DropPath, Mlp, PatchEmbed, to_2tuple, trunc_normal_
from .ptq import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear
from .vit_fquant import _act_kw, _weight_kw

__all__ = ['swin_tiny_patch4_window7_224', 'swin_small_patch4_window7_224', 'swin_base_patch4_window7_224']


def window_partition(x, window_size):
    """(B, H, W, C) -> (num_windows * B, window_size, window_size, C)   (swin_quant.py:19-33)"""
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    """(num_windows * B, window_size, window_size, C) -> (B, H, W, C)   (swin_quant.py:36-51)"""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(window_size):
    """Pair-wise relative position index inside a window (swin_quant.py:90-105)."""
    coords = torch.stack(torch.meshgrid([torch.arange(window_size[0]), torch.arange(window_size[1])], indexing='ij'))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window_size[0] - 1
    rel[:, :, 1] += window_size[1] - 1
    rel[:, :, 0] *= 2 * window_size[1] - 1
    return rel.sum(-1)


def shift_attention_mask(input_resolution, window_size, shift_size):
    """0 / -100 mask of the shifted-window attention (swin_quant.py:317-340); None without a shift."""
    if shift_size <= 0:
        return None
    H, W = input_resolution
    img_mask = torch.zeros((1, H, W, 1))
    slices = (slice(0, -window_size), slice(-window_size, -shift_size), slice(-shift_size, None))
    cnt = 0
    for h in slices:
        for w in slices:
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mask_windows = window_partition(img_mask, window_size).view(-1, window_size * window_size)
    attn_mask = mask_windows.unsqueeze(1) - mask_windows.unsqueeze(2)
    return attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))


class WindowAttention(nn.Module):
    """Window multi-head self attention with a quantized relative position bias (swin_quant.py:54-221)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        self.register_buffer('relative_position_index', relative_position_index(window_size))
        self.qkv = QLinear(dim, dim * 3, bias=qkv_bias, **_weight_kw(cfg, quant, calibrate))
        self.qact1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact_attn1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact_table = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate))
        self.attn_drop = nn.Dropout(attn_drop)
        self.log_int_softmax = QIntSoftmax(log_i_softmax=cfg.INT_SOFTMAX, quant=quant, calibrate=calibrate,
                                           bit_type=cfg.BIT_TYPE_S, calibration_mode=cfg.CALIBRATION_MODE_S,
                                           observer_str=cfg.OBSERVER_S, quantizer_str=cfg.QUANTIZER_S)
        self.qact3 = QAct(**_act_kw(cfg, quant, calibrate))
        self.qact4 = QAct(**_act_kw(cfg, quant, calibrate))
        self.proj = QLinear(dim, dim, **_weight_kw(cfg, quant, calibrate))
        self.proj_drop = nn.Dropout(proj_drop)
        trunc_normal_(self.relative_position_bias_table, std=0.02)

    def forward(self, x, mask=None, FLOPs=None, global_distance=None, bits=(None, None)):
        FLOPs = [] if FLOPs is None else FLOPs
        global_distance = [] if global_distance is None else global_distance
        B_, N, C = x.shape
        x = self.qkv(x, global_distance, bits[0])
        FLOPs.append(N * C * x.shape[2])
        x = self.qact1(x)
        qkv = x.reshape(B_, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        q = q * self.scale
        attn = q @ k.transpose(-2, -1)
        attn = self.qact_attn1(attn)
        table_q = self.qact_table(self.relative_position_bias_table)
        n = self.window_size[0] * self.window_size[1]
        bias = table_q[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
        attn = attn + bias.unsqueeze(0)
        attn = self.qact2(attn)
        if mask is not None:
            nW = mask.shape[0]
            attn = attn.view(B_ // nW, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn = self.log_int_softmax(attn, self.qact2.quantizer.scale)
        attn = self.attn_drop(attn)
        x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
        x = self.qact3(x)
        x = self.proj(x, global_distance, bits[1])
        FLOPs.append(N * C * x.shape[2])
        x = self.qact4(x)
        return self.proj_drop(x)


class SwinTransformerBlock(nn.Module):
    """(swin_quant.py:224-399)"""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:      # a single window: no partition, no shift
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, 'shift_size must in 0-window_size'
        self.norm1 = norm_layer(dim)
        self.qact1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop, quant=quant,
                                    calibrate=calibrate, cfg=cfg)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))
        self.norm2 = norm_layer(dim)
        self.qact3 = QAct(**_act_kw(cfg, quant, calibrate))
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop,
                       quant=quant, calibrate=calibrate, cfg=cfg)
        self.qact4 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))
        self.register_buffer('attn_mask', shift_attention_mask(self.input_resolution, self.window_size, self.shift_size))

    def forward(self, x, last_quantizer=None, FLOPs=None, global_distance=None, bits=(None, None, None, None)):
        FLOPs = [] if FLOPs is None else FLOPs
        global_distance = [] if global_distance is None else global_distance
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, 'input feature has wrong size'
        shortcut = x
        x = self.norm1(x, last_quantizer, self.qact1.quantizer)
        x = self.qact1(x)
        x = x.view(B, H, W, C)
        shifted = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2)) if self.shift_size > 0 else x
        windows = window_partition(shifted, self.window_size).view(-1, self.window_size * self.window_size, C)
        attn_windows = self.attn(windows, self.attn_mask, FLOPs, global_distance, bits[0:2])
        attn_windows = attn_windows.view(-1, self.window_size, self.window_size, C)
        shifted = window_reverse(attn_windows, self.window_size, H, W)
        x = torch.roll(shifted, shifts=(self.shift_size, self.shift_size), dims=(1, 2)) if self.shift_size > 0 else shifted
        x = x.view(B, H * W, C)
        x = shortcut + self.drop_path(x)
        x = self.qact2(x)
        y = self.qact3(self.norm2(x, self.qact2.quantizer, self.qact3.quantizer))
        x = x + self.drop_path(self.mlp(y, FLOPs, global_distance, tuple(bits[2:4])))
        return self.qact4(x)


class PatchMerging(nn.Module):
    """2x2 patch merging: concat, LayerNorm over 4C (input scales tiled 4x), QLinear 4C -> 2C (swin_quant.py:402-476)."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm, quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.norm = norm_layer(4 * dim)
        self.qact1 = QAct(**_act_kw(cfg, quant, calibrate))
        self.reduction = QLinear(4 * dim, 2 * dim, bias=False, **_weight_kw(cfg, quant, calibrate))
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate, ln=True))

    def forward(self, x, last_quantizer=None, FLOPs=None, global_distance=None, bit=None):
        FLOPs = [] if FLOPs is None else FLOPs
        global_distance = [] if global_distance is None else global_distance
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, 'input feature has wrong size'
        assert H % 2 == 0 and W % 2 == 0, f'x size ({H}*{W}) are not even.'
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
        x = x.view(B, -1, 4 * C)
        x = self.norm(x, last_quantizer, self.qact1.quantizer, None, 4)
        x = self.qact1(x)
        n_tok, c_in = x.shape[1], x.shape[2]
        x = self.reduction(x, global_distance, bit)
        FLOPs.append(n_tok * c_in * x.shape[2])
        return self.qact2(x)


class BasicLayer(nn.Module):
    """One stage (swin_quant.py:479-568)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads,
                                 window_size=window_size, shift_size=0 if (i % 2 == 0) else window_size // 2,
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, quant=quant, calibrate=calibrate, cfg=cfg)
            for i in range(depth)])
        self.downsample = (downsample(input_resolution, dim=dim, norm_layer=norm_layer, quant=quant,
                                      calibrate=calibrate, cfg=cfg) if downsample is not None else None)

    def num_linear_layers(self):
        return 4 * self.depth + (1 if self.downsample is not None else 0)

    def forward(self, x, last_quantizer=None, FLOPs=None, global_distance=None, bits=None):
        bits = bits if bits is not None else [None] * self.num_linear_layers()
        for i, blk in enumerate(self.blocks):
            lq = last_quantizer if i == 0 else self.blocks[i - 1].qact4.quantizer
            x = blk(x, lq, FLOPs, global_distance, tuple(bits[4 * i:4 * i + 4]))
        if self.downsample is not None:
            x = self.downsample(x, self.blocks[-1].qact4.quantizer, FLOPs, global_distance, bits[4 * self.depth])
        return x


class SwinTransformer(nn.Module):
    """(swin_quant.py:570-817)"""

    _Q_MODULES = (QConv2d, QLinear, QAct, QIntSoftmax)

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 quant=False, calibrate=False, input_quant=False, cfg=None, **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.input_quant = input_quant
        self.cfg = cfg
        self.quant = quant
        self.per_module = False      # True: every Q-module runs its own operator even without hooks
        self._engine = None
        self._engine_off = None      # why the integer engine refused this model (NotImplementedError text)
        self._int_active = bool(quant)
        self._submodules = None
        if input_quant:
            self.qact_input = QAct(**_act_kw(cfg, quant, calibrate))
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                      embed_dim=embed_dim, norm_layer=norm_layer if self.patch_norm else None,
                                      quant=quant, calibrate=calibrate, cfg=cfg)
        num_patches = self.patch_embed.num_patches
        self.patch_grid = self.patch_embed.grid_size
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
            trunc_normal_(self.absolute_pos_embed, std=0.02)
            self.qact1 = QAct(**_act_kw(cfg, quant, calibrate))
        else:
            self.absolute_pos_embed = None
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        layers = []
        for i in range(self.num_layers):
            layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i),
                input_resolution=(self.patch_grid[0] // (2 ** i), self.patch_grid[1] // (2 ** i)),
                depth=depths[i], num_heads=num_heads[i], window_size=window_size, mlp_ratio=self.mlp_ratio,
                qkv_bias=qkv_bias, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if (i < self.num_layers - 1) else None, quant=quant, calibrate=calibrate,
                cfg=cfg))
        self.layers = nn.Sequential(*layers)
        self.norm = norm_layer(self.num_features)
        self.qact2 = QAct(**_act_kw(cfg, quant, calibrate))
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.qact3 = QAct(**_act_kw(cfg, quant, calibrate))
        self.head = (QLinear(self.num_features, num_classes, **_weight_kw(cfg, quant, calibrate))
                     if num_classes > 0 else nn.Identity())
        self.act_out = QAct(**_act_kw(cfg, quant, calibrate))
        self.apply(self._init_weights)

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
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def get_classifier(self):
        return self.head

    # -- mode switches (swin_quant.py:759-788) -------------------------------------------------------------------
    def _set_flag(self, name, value):
        for m in self.modules():
            if type(m) in self._Q_MODULES:
                setattr(m, name, value)

    def model_quant(self):
        self.quant = True
        self._set_flag('quant', True)
        if self.cfg.INT_NORM:
            for m in self.modules():
                if type(m) is QIntLayerNorm:
                    m.mode = 'int'
        self._int_active = True
        self._drop_engine()

    def model_dequant(self):
        self._set_flag('quant', False)
        self._int_active = False
        self._drop_engine()

    def model_open_calibrate(self):
        self._set_flag('calibrate', True)
        self._drop_engine()       # scales are about to change

    # -- the integer engine ---------------------------------------------------------------------------------------------
    def _drop_engine(self):
        self._engine = None
        self._engine_off = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._drop_engine()       # .to() / .cuda(): plans live on one device
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._drop_engine()
        return out

    def integer_engine(self):
        """The sm_100a execution engine bound to this model's calibrated state (built lazily)."""
        if self._engine is None:
            from .swin_engine import SwinIntegerEngine
            self._engine = SwinIntegerEngine(self)
        return self._engine

    def _hooked(self):
        if self.per_module:
            return True
        if self._submodules is None:
            self._submodules = [m for m in self.modules() if m is not self]
        if _module_globals._global_forward_hooks or _module_globals._global_forward_pre_hooks:
            return True
        return any(m._forward_hooks or m._forward_pre_hooks for m in self._submodules)

    def _integer_forward(self, x, bits):
        """Logits from the integer engine, or None when this call stays on the per-module path."""
        if not (self.quant and self._int_active and torch.is_tensor(x) and x.is_cuda and self.input_quant
                and self.absolute_pos_embed is None and self.cfg.INT_NORM and self.cfg.INT_SOFTMAX):
            return None
        if self._engine_off is not None or any(b == -1 for b in bits) or self._hooked():
            return None
        try:
            return self.integer_engine().forward(x, bits)
        except NotImplementedError as e:      # float scales, zero points, ...: documented scope of swin_engine.py
            self._engine_off = str(e)
            return None

    def model_open_last_calibrate(self):
        self._set_flag('last_calibrate', True)

    def model_close_calibrate(self):
        self._set_flag('calibrate', False)

    # -- graph ------------------------------------------------------------------------------------------------------
    def num_linear_layers(self):
        return 1 + sum(layer.num_linear_layers() for layer in self.layers) + 1

    def forward_features(self, x, FLOPs, global_distance, bit_config):
        if self.input_quant:
            x = self.qact_input(x)
        x = self.patch_embed(x, FLOPs, bit_config[0])
        if self.absolute_pos_embed is not None:
            x = x + self.absolute_pos_embed
            x = self.qact1(x)
        x = self.pos_drop(x)
        pos = 1
        for i, layer in enumerate(self.layers):
            lq = self.patch_embed.qact.quantizer if i == 0 else self.layers[i - 1].downsample.qact2.quantizer
            cnt = layer.num_linear_layers()
            x = layer(x, lq, FLOPs, global_distance, list(bit_config[pos:pos + cnt]))
            pos += cnt
        x = self.norm(x, self.layers[-1].blocks[-1].qact4.quantizer, self.qact2.quantizer)
        x = self.qact2(x)
        x = self.avgpool(x.transpose(1, 2))
        x = self.qact3(x)
        return torch.flatten(x, 1)

    def forward(self, x, bit_config=None, plot=None, hessian_statistic=False):
        """forward(x) -> logits (the reference signature); with a bit_config or an explicit `plot` argument (the ViT
        models' call, used by calibrate_model / dist.validate): (logits, FLOPs, global_distance)."""
        n = self.num_linear_layers()
        bits = [8] * n if bit_config is None else list(bit_config)
        if len(bits) < n:
            raise IndexError('bit_config has %d entries, the model has %d quantized layers' % (len(bits), n))
        FLOPs, global_distance = [], []
        logits = self._integer_forward(x, bits[:n])
        if logits is not None:
            return logits if bit_config is None and plot is None else (logits, self.flops(), [])
        x = self.forward_features(x, FLOPs, global_distance, bits)
        c = x.shape[1]
        x = self.head(x, global_distance, bits[n - 1])
        FLOPs.append(c * x.shape[1])
        x = self.act_out(x)
        return x if bit_config is None and plot is None else (x, FLOPs, global_distance)


def _swin_flops(self):
    """The per-layer MAC list a per-module forward accumulates (patch embed, per block qkv / proj / fc1 / fc2, the
    reductions, head)."""
    pe = self.patch_embed
    out = [pe.proj.in_channels * pe.patch_size[0] * pe.patch_size[0] * self.embed_dim * pe.grid_size[0] * pe.grid_size[1]]
    for layer in self.layers:
        for blk in layer.blocks:
            n, c = blk.window_size * blk.window_size, blk.dim
            out += [n * c * 3 * c, n * c * c]
            t = blk.input_resolution[0] * blk.input_resolution[1]
            hid = blk.mlp.fc1.out_features
            out += [t * c * hid, t * hid * c]
        if layer.downsample is not None:
            t = layer.input_resolution[0] * layer.input_resolution[1] // 4
            out.append(t * 4 * layer.dim * 2 * layer.dim)
    out.append(self.num_features * self.num_classes)
    return out


SwinTransformer.flops = _swin_flops


def _factory(embed_dim, depths, num_heads):
    def build(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
        if pretrained:
            raise RuntimeError('pretrained checkpoints need network access (torch.hub); load a state_dict with the '
                               'reference key names instead')
        kw = dict(patch_size=4, window_size=7, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
                  norm_layer=QIntLayerNorm, quant=quant, calibrate=calibrate, input_quant=True, cfg=cfg)
        kw.update(kwargs)
        return SwinTransformer(**kw)
    return build


# reference: models/swin_quant.py:820-901
swin_tiny_patch4_window7_224 = _factory(96, (2, 2, 6, 2), (3, 6, 12, 24))
swin_small_patch4_window7_224 = _factory(96, (2, 2, 18, 2), (3, 6, 12, 24))
swin_base_patch4_window7_224 = _factory(128, (2, 2, 18, 2), (4, 8, 16, 32))
for _n in __all__:
    globals()[_n].__name__ = _n


def extract_swin_state(model):
    """Calibrated state of a ``SwinTransformer`` as CPU tensors (see oracle/swin_fakequant_forward.py): float
    parameters and buffers, activation scales / zero points, per-bit-width weight scales, the Mlp SmoothQuant scales."""
    cfg = model.cfg
    f32 = lambda t: t.detach().to('cpu', torch.float32).clone()
    pe = model.patch_embed
    arch = dict(img_size=pe.img_size[0], patch_size=pe.patch_size[0], in_chans=pe.proj.in_channels,
                num_classes=model.num_classes, embed_dim=model.embed_dim,
                depths=tuple(layer.depth for layer in model.layers),
                num_heads=tuple(layer.blocks[0].num_heads for layer in model.layers),
                window_size=max(blk.window_size for layer in model.layers for blk in layer.blocks),
                softmax_bits=cfg.BIT_TYPE_S.bits)
    state = dict(arch=arch, params={k: f32(v) for k, v in model.state_dict().items()}, act={}, weight={}, cs={})
    for name, m in model.named_modules():
        if isinstance(m, QAct):
            if m.quantizer.scale is None:
                raise RuntimeError('%s has no scale: calibrate the model before model_quant()' % name)
            state['act'][name] = (f32(m.quantizer.scale), f32(m.quantizer.zero_point),
                                  m.quantizer.bit_type.lower_bound, m.quantizer.bit_type.upper_bound)
        elif isinstance(m, (QLinear, QConv2d)):
            state['weight'][name] = {b: (f32(s), f32(m.quantizer.dic_zero_point[b]))
                                     for b, s in m.quantizer.dic_scale.items()}
        elif isinstance(m, Mlp):
            if m.channel_scale is None:
                raise RuntimeError('%s was never calibrated' % name)
            state['cs'][name] = f32(m.best_scale[-1])
    for name, m in model.named_modules():       # the SmoothQuant pair keeps its calibrated grids on the owner
        if isinstance(m, Mlp):
            bt = m.qact0.quantizer.bit_type
            state['act'][name + '.qact0'] = (f32(m.best_act_scale[-1]), f32(m.best_act_zp[-1]), bt.lower_bound,
                                             bt.upper_bound)
            ws, wz = m.best_weight_scale[-1], m.best_weight_zp[-1]
            state['weight'][name + '.fc1'] = {b: (f32(ws[b]), f32(wz[b])) for b in ws}
    return state
