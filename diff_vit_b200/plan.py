"""From a calibrated model to the integer plan the sm_100a engine executes.

``extract_state`` snapshots what calibration left on the modules (float parameters, activation
scales / zero points, per-bit-width weight scales, SmoothQuant channel scales) as plain CPU tensors;
the same dict drives the CPU oracle in tests.  ``build_plan`` turns it, for one ``bit_config``, into
integer weight codes plus the per-channel fp32 vectors of every fused epilogue
(SURVEY.md section 8a "per-block dataflow").
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .ptq import QAct, QConv2d, QIntSoftmax, QLinear

_W_RANGE = {4: (-8, 7), 8: (-128, 127)}


def extract_state(model):
    """Calibrated state of a ``VisionTransformer`` as CPU tensors (see oracle/fakequant_forward.py)."""
    cfg = model.cfg
    if not (cfg.INT_NORM and cfg.INT_SOFTMAX):
        raise NotImplementedError('the integer engine needs Config(ptf=True, lis=True): without them the '
                                  'reference runs float LayerNorm / softmax in quantized mode')
    pe = model.patch_embed
    arch = dict(img_size=pe.img_size[0], patch_size=pe.patch_size[0], in_chans=pe.proj.in_channels,
                embed_dim=model.embed_dim, depth=model.depth, num_heads=model.num_heads,
                hidden_dim=model.blocks[0].mlp.fc1.out_features, num_classes=model.num_classes,
                attn_scale=float(model.blocks[0].attn.scale), softmax_bits=cfg.BIT_TYPE_S.bits,
                num_patches=pe.num_patches)
    if not model.input_quant:
        raise NotImplementedError('input_quant=False (vit_large) multiplies unquantized pixels by integer '
                                  'weights, which is not an integer GEMM (SURVEY.md a15)')
    f32 = lambda t: t.detach().to('cpu', torch.float32).clone()
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
    for i, blk in enumerate(model.blocks):
        for sub, owner, qact0, lin in (('attn', blk.attn, blk.attn.qact0, 'attn.qkv'),
                                       ('mlp', blk.mlp, blk.mlp.qact0, 'mlp.fc1')):
            pre = 'blocks.%d.%s' % (i, sub)
            if owner.channel_scale is None:
                raise RuntimeError('%s was never calibrated' % pre)
            # both bit-pool slots alias the same calibrated objects (single-alpha pool)
            state['cs'][pre] = f32(owner.best_scale[-1])
            bt = qact0.quantizer.bit_type
            state['act'][pre + '.qact0'] = (f32(owner.best_act_scale[-1]), f32(owner.best_act_zp[-1]),
                                            bt.lower_bound, bt.upper_bound)
            ws, wz = owner.best_weight_scale[-1], owner.best_weight_zp[-1]
            state['weight']['blocks.%d.%s' % (i, lin)] = {b: (f32(ws[b]), f32(wz[b])) for b in ws}
    return state


def is_pot(t):
    """True when every element is an exact power of two."""
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    if not bool((t > 0).all()):
        return False
    mant, _ = torch.frexp(t)
    return bool((mant == 0.5).all())


@dataclass
class LinearPlan:
    w: torch.Tensor            # int8 [n, k]
    acc_scale: torch.Tensor    # fp32 [n]
    bias: torch.Tensor         # fp32 [n]
    out_scale: torch.Tensor    # fp32 [n]
    out_rscale: torch.Tensor   # fp32 [n]
    out_zp: float
    flags: int
    res_scale: Optional[torch.Tensor] = None
    out2_scale: Optional[torch.Tensor] = None
    bits: int = 8                             # weight width of this layer (4: codes in [-8, 7])
    w4: Optional[torch.Tensor] = None         # uint8 [n, k / 2]: the int4 codes two per byte (even k in the low nibble);
                                              # serialised plans carry 4-bit layers this way and leave `w` empty

    def codes(self):
        """int8 [n, k] weight codes, unpacking the nibbles of a packed 4-bit layer (host side)."""
        if self.w is not None:
            return self.w
        return unpack_int4(self.w4)

    def pack(self):
        """Move a 4-bit layer to its packed form (in place); 8-bit layers are left alone."""
        if self.bits == 4 and self.w is not None and self.w.shape[1] % 2 == 0:
            self.w4, self.w = pack_int4(self.w), None
        return self


def pack_int4(w):
    """int8 codes in [-8, 7], [n, k] with k even -> uint8 [n, k / 2], element 2i in the low nibble of byte i."""
    w = w.to(torch.int16)
    if int(w.min()) < -8 or int(w.max()) > 7:
        raise ValueError('pack_int4: codes outside [-8, 7]')
    return ((w[:, 0::2] & 0xF) | ((w[:, 1::2] & 0xF) << 4)).to(torch.uint8).contiguous()


def unpack_int4(packed):
    p = packed.to(torch.int16)
    lo, hi = p & 0xF, (p >> 4) & 0xF
    out = torch.stack((lo, hi), dim=-1).reshape(packed.shape[0], -1)
    return torch.where(out > 7, out - 16, out).to(torch.int8).contiguous()


@dataclass
class LayerNormPlan:
    in_mask: torch.Tensor
    gamma: torch.Tensor
    beta: torch.Tensor
    ln_out_scale: torch.Tensor
    ln_out_rscale: torch.Tensor
    post_mul: torch.Tensor
    post_div1: torch.Tensor
    post_div2: float
    post_zp: float
    in_scale1: float
    pot: int
    pre_clamp: int = 0         # the LN code passes an int8 QAct of its own grid before the re-gridding (Swin)


@dataclass
class AttentionPlan:
    score_mul: float
    score_zp: float
    out_mul: float
    out_zp: float
    levels: int
    exp_lut: torch.Tensor      # fp32 [256]
    in_zp: float = 0.0         # zero point of the q/k/v codes (asymmetric observers)

    @property
    def lut_sig_bits(self):
        """Widest entry of exp_lut in significant bits (p2v_attention.lut_sig_bits: the kernel choice of the library)."""
        widest = 0
        for e in self.exp_lut.double().tolist():
            v = int(e)
            if v > 0:
                widest = max(widest, v.bit_length() - ((v & -v).bit_length() - 1))
        return widest


@dataclass
class BlockPlan:
    norm1: LayerNormPlan
    norm2: LayerNormPlan
    qkv: LinearPlan
    proj: LinearPlan
    fc1: LinearPlan
    fc2: LinearPlan
    attn: AttentionPlan


@dataclass
class VitPlan:
    arch: dict
    bit_config: tuple
    input_scale: float
    input_zp: float
    patch_embed: LinearPlan
    pe_scale: float
    pe_zp: float
    embed_scale: float
    embed_zp: float
    cls_value: torch.Tensor
    pos_value: torch.Tensor
    embed_out_scale: torch.Tensor
    blocks: List[BlockPlan] = field(default_factory=list)
    norm: LayerNormPlan = None
    head: LinearPlan = None

    def tensors(self):
        """Every tensor of the plan (for upload)."""
        def walk(obj):
            if isinstance(obj, torch.Tensor):
                yield obj
            elif isinstance(obj, (list, tuple)):
                for o in obj:
                    yield from walk(o)
            elif hasattr(obj, '__dataclass_fields__'):
                for f in obj.__dataclass_fields__:
                    yield from walk(getattr(obj, f))
        yield from walk(self)


def _scalar(t, what):
    t = torch.as_tensor(t).reshape(-1)
    if t.numel() != 1:
        raise NotImplementedError('%s: expected a layer-wise (scalar) quantizer, got %d values' % (what, t.numel()))
    return t


def _expand(t, n):
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    return (t.expand(n) if t.numel() == 1 else t).contiguous().clone()


def softmax_exp_lut(score_scale):
    """Integer exp of the log-int-softmax for every possible distance d = rowmax - code in [0, 255],
    evaluated with the reference's own tensor expressions (models/ptq/layers.py:334-358)."""
    x0_int, b_int, c_int = QIntSoftmax.exp_constants(score_scale)
    n = QIntSoftmax.EXP_BITS
    x_int = -torch.arange(256, dtype=torch.float32)
    x_int = torch.max(x_int, n * x0_int)
    q = torch.floor(x_int / x0_int)
    r = x_int - x0_int * q
    z = r + b_int
    z = r * z
    z = z + c_int
    e = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
    if not bool((e > 0).all()) or float(e.max()) >= 2.0 ** 62 or float(z.max()) >= 2.0 ** 24:
        raise NotImplementedError('score scale %g puts the integer exp outside the exact range' % float(score_scale))
    return e.contiguous()


class _Builder:

    def __init__(self, state, bit_config):
        self.s = state
        self.P = state['params']
        self.arch = state['arch']
        self.bits = list(bit_config)
        n_layers = 4 * self.arch['depth'] + 2
        if len(self.bits) < n_layers:
            raise IndexError('bit_config has %d entries, the model has %d quantized layers' % (len(self.bits), n_layers))

    def act(self, name):
        return self.s['act'][name]

    def linear(self, name, weight, bits, in_act, out_act, gelu=False, residual=None):
        """residual = (res_act_name, out2_act_name) for the two residual-carrying layers."""
        if bits not in _W_RANGE:
            raise KeyError('int%s' % bits)  # the reference indexes BIT_TYPE_DICT['int' + str(bit)]
        w_scale, w_zp = self.s['weight'][name]['int%d' % bits]
        lo, hi = _W_RANGE[bits]
        n = weight.shape[0]
        w2 = weight.reshape(n, -1)
        codes = (w2 / w_scale.reshape(-1, 1) + w_zp.reshape(-1, 1)).round().clamp(lo, hi)
        if bool((w_zp != 0).any()):
            raise NotImplementedError('%s: asymmetric weight quantizers are not produced by the minmax observer' % name)
        in_scale, in_zp, _, _ = self.act(in_act)
        in_scale = _scalar(in_scale, in_act)
        in_zp = float(_scalar(in_zp, in_act))
        out_scale, out_zp, _, _ = self.act(out_act)
        out_vec = _expand(out_scale, n)
        pot = is_pot(out_vec) and float(torch.as_tensor(out_zp).reshape(-1)[0]) == 0.0
        flags = (1 if gelu else 0) | (4 if pot else 0)
        acc_scale = _expand(in_scale * w_scale.reshape(-1), n)
        bias = self.P[name + '.bias'].clone()
        if in_zp != 0.0:
            # asymmetric input (omse observer): sum_k (q_k - z) w_nk = acc_n - z * sum_k w_nk; the second term is a
            # per-channel constant and joins the bias, so the GEMM itself still multiplies raw int8 codes
            bias = (bias.double() - in_zp * acc_scale.double() * codes.double().sum(1)).float()
        plan = LinearPlan(w=codes.to(torch.int8).contiguous(), acc_scale=acc_scale, bits=int(bits),
                          bias=bias, out_scale=out_vec, out_rscale=1.0 / out_vec,
                          out_zp=float(torch.as_tensor(out_zp).reshape(-1)[0]), flags=flags)
        if residual is not None:
            res_act, out2_act = residual
            if float(self.act(res_act)[1].abs().max()) != 0.0 or float(self.act(out2_act)[1].abs().max()) != 0.0 \
                    or plan.out_zp != 0.0:
                raise NotImplementedError('residual quantizers are expected to be symmetric (ptf observer)')
            plan.res_scale = _expand(self.act(res_act)[0], n)
            plan.out2_scale = _expand(self.act(out2_act)[0], n)
        return plan

    def layernorm(self, in_act, gamma, beta, out_act, ln_cs=None, next_cs=None):
        """ln_cs: SmoothQuant scale handed to the LayerNorm (out grid = s_out * ln_cs);
        next_cs: the scale the consumer divides by before its QAct."""
        d = gamma.numel()
        in_scale = _expand(self.act(in_act)[0], d)
        in_scale1 = in_scale.min()
        out_s, out_zp, _, _ = self.act(out_act)
        out_s = _scalar(out_s, out_act)
        ln_out = _expand(out_s * ln_cs if ln_cs is not None else out_s, d)
        div1 = _expand(next_cs if next_cs is not None else torch.ones(1), d)
        pot = is_pot(ln_out) and is_pot(div1) and is_pot(out_s)
        return LayerNormPlan(in_mask=(in_scale / in_scale1).round().contiguous(), gamma=gamma.clone(), beta=beta.clone(),
                             ln_out_scale=ln_out, ln_out_rscale=1.0 / ln_out, post_mul=(ln_out / div1 / out_s).contiguous(),
                             post_div1=div1, post_div2=float(out_s), post_zp=float(torch.as_tensor(out_zp).reshape(-1)[0]),
                             in_scale1=float(in_scale1), pot=int(pot))

    def attention(self, pre):
        s1, z1, _, _ = self.act(pre + '.qact1')
        sa, za, _, _ = self.act(pre + '.qact_attn1')
        s2, z2, _, _ = self.act(pre + '.qact2')
        s1, sa, s2 = (float(_scalar(v, pre)) for v in (s1, sa, s2))
        score_mul = s1 * s1 * self.arch['attn_scale'] / sa
        return AttentionPlan(score_mul=float(torch.tensor(score_mul, dtype=torch.float32)),
                             score_zp=float(_scalar(za, pre)), out_mul=(2.0 ** -15) * s1 / s2,
                             out_zp=float(_scalar(z2, pre)), levels=2 ** self.arch['softmax_bits'],
                             exp_lut=softmax_exp_lut(_scalar(self.act(pre + '.qact_attn1')[0], pre)),
                             in_zp=float(_scalar(z1, pre)))

    def build(self):
        P, arch, bits = self.P, self.arch, self.bits
        in_s, in_z, _, _ = self.act('qact_input')
        pe_s, pe_z, _, _ = self.act('patch_embed.qact')
        em_s, em_z, lo, hi = self.act('qact_embed')
        po_s, po_z, plo, phi = self.act('qact_pos')
        em_s, em_z, po_s, po_z = (_scalar(v, 'embed') for v in (em_s, em_z, po_s, po_z))
        cls = P['cls_token'].reshape(-1)
        cls_value = ((cls / em_s + em_z).round().clamp(lo, hi) - em_z) * em_s
        pos = P['pos_embed'].reshape(-1, arch['embed_dim'])
        pos_value = ((pos / po_s + po_z).round().clamp(plo, phi) - po_z) * po_s
        plan = VitPlan(arch=arch, bit_config=tuple(bits), input_scale=float(_scalar(in_s, 'qact_input')),
                       input_zp=float(_scalar(in_z, 'qact_input')),
                       patch_embed=self.linear('patch_embed.proj', P['patch_embed.proj.weight'], bits[0], 'qact_input',
                                               'patch_embed.qact'),
                       pe_scale=float(_scalar(pe_s, 'pe')), pe_zp=float(_scalar(pe_z, 'pe')), embed_scale=float(em_s),
                       embed_zp=float(em_z), cls_value=cls_value.contiguous(), pos_value=pos_value.contiguous(),
                       embed_out_scale=_expand(self.act('qact1')[0], arch['embed_dim']))
        if float(self.act('qact1')[1].abs().max()) != 0.0:
            raise NotImplementedError('qact1 is expected to be symmetric (ptf observer)')
        stream = 'qact1'
        for i in range(arch['depth']):
            pre = 'blocks.%d' % i
            b = bits[4 * i + 1:4 * i + 5]
            cs_a, cs_m = self.s['cs'][pre + '.attn'], self.s['cs'][pre + '.mlp']
            blk = BlockPlan(
                norm1=self.layernorm(stream, P[pre + '.norm1.weight'], P[pre + '.norm1.bias'], pre + '.attn.qact0',
                                     ln_cs=cs_a, next_cs=cs_a),
                qkv=self.linear(pre + '.attn.qkv', P[pre + '.attn.qkv.weight'] * cs_a.reshape(1, -1), b[0],
                                pre + '.attn.qact0', pre + '.attn.qact1'),
                attn=self.attention(pre + '.attn'),
                proj=self.linear(pre + '.attn.proj', P[pre + '.attn.proj.weight'], b[1], pre + '.attn.qact2',
                                 pre + '.attn.qact3', residual=(stream, pre + '.qact2')),
                # norm2 receives the ATTENTION channel scale (reference quirk, vit_fquant.py:464); Mlp re-grids
                norm2=self.layernorm(pre + '.qact2', P[pre + '.norm2.weight'], P[pre + '.norm2.bias'],
                                     pre + '.mlp.qact0', ln_cs=cs_a, next_cs=cs_m),
                fc1=self.linear(pre + '.mlp.fc1', P[pre + '.mlp.fc1.weight'] * cs_m.reshape(1, -1), b[2],
                                pre + '.mlp.qact0', pre + '.mlp.qact1', gelu=True),
                fc2=self.linear(pre + '.mlp.fc2', P[pre + '.mlp.fc2.weight'], b[3], pre + '.mlp.qact1',
                                pre + '.mlp.qact2', residual=(pre + '.qact2', pre + '.qact4')))
            plan.blocks.append(blk)
            stream = pre + '.qact4'
        plan.norm = self.layernorm(stream, P['norm.weight'], P['norm.bias'], 'qact2')
        plan.head = self.linear('head', P['head.weight'], bits[-1], 'qact2', 'act_out')
        return plan


def build_plan(state, bit_config):
    """Integer plan for one bit_config (index map: 0 conv, 1+4i.. block i qkv/proj/fc1/fc2, -1 head)."""
    return _Builder(state, bit_config).build()


# ---- serialised plans -------------------------------------------------------------------------------------------
# The reference keeps calibrated scales only as Python attributes ("checkpoint / resume");
# an integer plan is self-contained (int8 codes + fp32 vectors + a few scalars), so it can be written once and
# executed later without the float model, the calibration data or PyTorch modules.
_PLAN_CLASSES = {c.__name__: c for c in (LinearPlan, LayerNormPlan, AttentionPlan, BlockPlan, VitPlan)}


def _is_record(obj):
    from types import SimpleNamespace
    return hasattr(obj, '__dataclass_fields__') or isinstance(obj, SimpleNamespace)


def _flatten(obj, prefix, out):
    from types import SimpleNamespace
    if isinstance(obj, torch.Tensor):
        out[prefix] = obj.detach().cpu().numpy()
    elif hasattr(obj, '__dataclass_fields__'):
        out[prefix + '/__class__'] = type(obj).__name__
        for f in obj.__dataclass_fields__:
            _flatten(getattr(obj, f), prefix + '/' + f, out)
    elif isinstance(obj, SimpleNamespace):           # the Swin plans (swin_engine.py) are namespaces of the same leaves
        out[prefix + '/__ns__'] = ','.join(vars(obj))
        for f, v in vars(obj).items():
            _flatten(v, prefix + '/' + f, out)
    elif isinstance(obj, (list, tuple)) and obj and _is_record(obj[0]):
        out[prefix + '/__len__'] = len(obj)
        for i, o in enumerate(obj):
            _flatten(o, '%s/%d' % (prefix, i), out)
    elif isinstance(obj, (list, tuple)):             # plain tuples (bit_config, a stage's resolution): keep the type
        out[prefix + '/__tuple__'] = repr(tuple(obj))
    elif isinstance(obj, dict):
        out[prefix + '/__dict__'] = repr(sorted(obj.items()))
    elif obj is None:
        out[prefix + '/__none__'] = 1
    else:
        out[prefix] = obj


def _linear_plans(plan):
    yield plan.patch_embed
    for blk in plan.blocks:
        yield from (blk.qkv, blk.proj, blk.fc1, blk.fc2)
    yield plan.head


def save_plan(plan, path, pack4=True):
    """Write a VitPlan to a compressed .npz file.  4-bit layers are stored int4-packed (two codes per byte);
    the engine unpacks them on the device when the plan is bound (`p2v_unpack_int4`)."""
    import copy
    import numpy as np
    if pack4:
        plan = copy.copy(plan)
        plan.patch_embed, plan.head = copy.copy(plan.patch_embed).pack(), copy.copy(plan.head).pack()
        plan.blocks = [copy.copy(b) for b in plan.blocks]
        for b in plan.blocks:
            b.qkv, b.proj, b.fc1, b.fc2 = (copy.copy(l).pack() for l in (b.qkv, b.proj, b.fc1, b.fc2))
    flat = {}
    _flatten(plan, 'plan', flat)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in flat.items()})


def _unflatten(z, prefix):
    import ast
    if prefix + '/__none__' in z:
        return None
    if prefix + '/__dict__' in z:
        return dict(ast.literal_eval(str(z[prefix + '/__dict__'])))
    if prefix + '/__len__' in z:
        return [_unflatten(z, '%s/%d' % (prefix, i)) for i in range(int(z[prefix + '/__len__']))]
    if prefix + '/__tuple__' in z:
        return tuple(ast.literal_eval(str(z[prefix + '/__tuple__'])))
    if prefix + '/__ns__' in z:
        from types import SimpleNamespace
        names = [f for f in str(z[prefix + '/__ns__']).split(',') if f]
        return SimpleNamespace(**{f: _unflatten(z, prefix + '/' + f) for f in names})
    if prefix + '/__class__' in z:
        cls = _PLAN_CLASSES[str(z[prefix + '/__class__'])]
        import dataclasses
        have = lambda f: any(k == prefix + '/' + f or k.startswith(prefix + '/' + f + '/') for k in z)
        return cls(**{f.name: _unflatten(z, prefix + '/' + f.name) for f in dataclasses.fields(cls)
                      if have(f.name) or f.default is dataclasses.MISSING})    # fields newer than the file keep defaults
    a = z[prefix]
    if a.ndim == 0:
        return a.item()
    if a.dtype.kind in 'iu' and a.ndim == 1 and prefix.endswith('bit_config'):
        return tuple(int(v) for v in a)
    return torch.from_numpy(a.copy())


def load_plan(path):
    """Read a VitPlan written by `save_plan`."""
    import numpy as np
    with np.load(path) as z:
        return _unflatten({k: z[k] for k in z.files}, 'plan')
