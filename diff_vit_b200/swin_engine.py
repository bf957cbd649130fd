"""Integer execution of the quantized Swin Transformer (BASELINE config 5) on the sm_100a kernels.

``build_swin_plan`` turns the calibrated state (``swin_quant.extract_swin_state``) into integer weight codes plus the
per-channel vectors of every fused epilogue, the window permutations / shift-mask regions / relative-position bias of
every attention layer and the 2 x 2 gather of every PatchMerging - plain CPU tensors, executable on the host by
tests/hostmath for the CPU test-suite.  ``SwinIntegerEngine`` uploads a plan, describes it as a ``p2v_swin_desc`` and
runs it with ONE C call per forward (``p2v_swin_forward``, include/p2v.h; captured into a CUDA graph per batch size):
tcgen05 GEMMs with fused re-quantisation / GELU / residual epilogues, the integer LayerNorm, the window attention
kernel (csrc/p2v_swin.cu).  ``forward_dump`` issues the same launches one by one from Python (``_run``) with every
intermediate code tensor copied out - the path the parity tests compare with the oracle, and with ``forward``.  Activations are int8 codes in token order from the patch embedding to
the head; the cyclic shift and the window partition are index permutations inside the attention kernel.

Per block (models/swin_quant.py:345-399, WindowAttention.forward :177-221, Mlp models/layers_quant.py:304-346):
  LN1 + qact1  ->  qkv GEMM + attn.qact1  ->  window attention (.. qact3)  ->  proj GEMM + qact4 + shortcut + qact2
  ->  LN2 + qact3 + / SmoothQuant scale + mlp.qact0 (one launch)  ->  fc1 GEMM + GELU + qact1  ->  fc2 GEMM + qact2 + residual + qact4

Scope: symmetric quantizers on power-of-two grids for the layer-wise activations (the minmax observer of config 5);
anything else raises NotImplementedError and the model keeps its per-module path.
"""
import ctypes as C
import math
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import _cabi
from .plan import LayerNormPlan, _Builder, _expand, _scalar, is_pot
from .ptq import QIntSoftmax


def num_linear_layers(arch):
    n = 1
    for i, d in enumerate(arch['depths']):
        n += 4 * d + (1 if i < len(arch['depths']) - 1 else 0)
    return n + 1


def _window_partition(x, ws):
    B, H, W, Cc = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, Cc)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, Cc)


def window_permutation(res, ws, shift):
    """perm[w * n + i] = token (row-major in the H x W grid) that roll(-shift) + window_partition put at row i of
    window w (models/swin_quant.py:362-369); window_reverse + roll(+shift) is its inverse (:380-385)."""
    H, W = res
    tok = torch.arange(H * W, dtype=torch.int32).view(1, H, W, 1)
    if shift > 0:
        tok = torch.roll(tok, shifts=(-shift, -shift), dims=(1, 2))
    return _window_partition(tok, ws).reshape(-1).contiguous()


def shift_regions(res, ws, shift):
    """Region id of every window row under the shifted-window mask (models/swin_quant.py:317-340): two rows of a
    window attend to each other iff their ids agree.  None without a shift."""
    if shift <= 0:
        return None
    H, W = res
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    return _window_partition(img, ws).reshape(-1).to(torch.uint8).contiguous()


def swin_exp_lut(scale):
    """Integer exp of the log-int-softmax for every distance d = rowmax - x, d = 0 .. n * |x0_int| (beyond that the
    reference clamps x_int, models/ptq/layers.py:345), with the reference's own fp32 tensor expressions.  The shift
    mask's -100 puts masked keys far beyond int8 distances, hence the longer table than the ViT one."""
    x0_int, b_int, c_int = QIntSoftmax.exp_constants(scale)
    n = QIntSoftmax.EXP_BITS
    period = int(-float(x0_int))
    if period < 1 or n * period + 1 > (1 << 16):
        raise NotImplementedError('score scale %g: integer-exp table of %d entries' % (float(scale), n * period + 1))
    x_int = -torch.arange(n * period + 1, dtype=torch.float32)
    x_int = torch.max(x_int, n * x0_int)
    q = torch.floor(x_int / x0_int)
    r = x_int - x0_int * q
    z = r + b_int
    z = r * z
    z = z + c_int
    e = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
    if not bool((e > 0).all()) or float(e.max()) >= 2.0 ** 62 or float(z.max()) >= 2.0 ** 24:
        raise NotImplementedError('score scale %g puts the integer exp outside the exact range' % float(scale))
    return e.contiguous()


class _SwinBuilder(_Builder):

    def __init__(self, state, bit_config):
        self.s = state
        self.P = dict(state['params'])
        self.arch = state['arch']
        n = num_linear_layers(self.arch)
        self.bits = [8] * n if bit_config is None else list(bit_config)
        if len(self.bits) < n:
            raise IndexError('bit_config has %d entries, the model has %d quantized layers' % (len(self.bits), n))
        for name, (_, zp, _, _) in state['act'].items():
            if float(torch.as_tensor(zp).abs().max()) != 0.0:
                raise NotImplementedError('%s: the Swin integer engine expects symmetric quantizers (minmax)' % name)

    def pot_scalar(self, name):
        s = _scalar(self.act(name)[0], name)
        if not is_pot(s):
            raise NotImplementedError('%s: scale %g is not a power of two' % (name, float(s)))
        return float(s)

    def ln_plan(self, in_scale, gamma, beta, out_act):
        d = gamma.numel()
        if d % 4 or d > 2048:
            raise NotImplementedError('LayerNorm over %d channels (p2v_layernorm_int: multiples of 4 up to 2048)' % d)
        in_scale = _expand(in_scale, d)
        in_scale1 = in_scale.min()
        out_s = _expand(self.act(out_act)[0], d)
        pot = is_pot(out_s)
        # the QAct behind the LayerNorm has the LayerNorm's own output grid: re-gridding factor 1
        return LayerNormPlan(in_mask=(in_scale / in_scale1).round().contiguous(), gamma=gamma.clone(), beta=beta.clone(),
                             ln_out_scale=out_s, ln_out_rscale=1.0 / out_s, post_mul=torch.ones(d),
                             post_div1=out_s.clone(), post_div2=1.0, post_zp=0.0, in_scale1=float(in_scale1), pot=int(pot))

    def window_attention(self, pre, heads, res, ws, shift):
        P = self.P
        n = ws * ws
        if n > 64:
            raise NotImplementedError('window of %d tokens' % n)
        dim = P[pre + '.qkv.weight'].shape[1]
        if dim != heads * 32:
            raise NotImplementedError('head dimension %d (the window attention kernel is built for 32)' % (dim // heads))
        s1, sa, s2, s3 = (self.pot_scalar(pre + k) for k in ('.qact1', '.qact_attn1', '.qact2', '.qact3'))
        st, _, tlo, thi = self.act(pre + '.qact_table')
        st = _scalar(st, pre)
        table = P[pre + '.relative_position_bias_table']
        tq = (table / st).round().clamp(tlo, thi)
        idx = P[pre + '.relative_position_index'].long().reshape(n, n)
        bias = (tq * st)[idx.reshape(-1)].reshape(n, n, heads)           # [row, key, head]
        qscale = float(torch.tensor((dim // heads) ** -0.5, dtype=torch.float32))
        qshift = 23 - math.floor(math.log2(qscale))
        codes = np.arange(-128, 128, dtype=np.float32)
        m = (codes * np.float32(qscale)).astype(np.float32).astype(np.float64) * 2.0 ** qshift
        if not (np.all(m == np.rint(m)) and np.abs(m).max() < 2.0 ** 31):
            raise NotImplementedError('q scaling %g does not fit the fixed-point product' % qscale)
        lut = swin_exp_lut(torch.tensor(s2))
        r = 1.0 / (3.0 * lut.double())
        r3 = torch.stack([(r * (1.0 - 2.0 ** -20)).float(), (r * (1.0 + 2.0 ** -20)).float()], -1).contiguous()
        if qscale * s1 * s1 / sa > 4.0:
            raise NotImplementedError('%s: qact_attn1 grid %g is too fine for scores of scale %g' % (pre, sa, qscale * s1 * s1))
        mask = 100.0 / s2
        if mask != int(mask):
            raise NotImplementedError('%s: 100 / qact2 scale is not an integer' % pre)
        return NS(perm=window_permutation(res, ws, shift), region=shift_regions(res, ws, shift),
                  bias=bias.permute(2, 1, 0).contiguous(),               # [head][key][row]
                  table_codes=tq.to(torch.int32), exp_lut=lut, exp_lut64=lut.double().contiguous(), r3=r3,
                  qk_scale=qscale * s1 * s1, err_mul=float(np.nextafter(np.float32(2.0 ** -24 * qscale * 128.0 * s1 * s1 / sa), np.float32(np.inf))),
                  n=n, heads=heads, windows=(res[0] // ws) * (res[1] // ws), tokens=res[0] * res[1], channels=dim,
                  qshift=qshift, qscale=qscale, acc_scale=s1 * s1 * 2.0 ** -qshift, a1_scale=sa, a1_rscale=1.0 / sa,
                  a2_rscale=1.0 / s2, mask_int=int(mask), out_unit=2.0 ** -15 * s1, out_rscale=1.0 / s3,
                  levels=2 ** self.arch['softmax_bits'])

    def block(self, pre, in_name, heads, res, ws, shift, bits):
        P = self.P
        cs = self.s['cs'][pre + '.mlp']
        s3 = self.pot_scalar(pre + '.qact3')
        s0 = self.pot_scalar(pre + '.mlp.qact0')
        d = cs.numel()
        # LN2 + qact3 + (/ channel_scale) + mlp.qact0 as ONE LayerNorm launch: the LN code is clamped to int8 (qact3 has
        # the LayerNorm's own grid) and re-gridded by s3 / (cs[c] * s0) - a power of two when the SmoothQuant scales are
        # (multiply, exact); otherwise the kernel's division path divides by cs[c] and by s0 like the reference
        norm2 = self.ln_plan(self.act(pre + '.qact2')[0], P[pre + '.norm2.weight'], P[pre + '.norm2.bias'], pre + '.qact3')
        norm2.pre_clamp = 1
        norm2.post_mul = (s3 / (cs.reshape(-1) * s0)).contiguous()
        norm2.post_div1, norm2.post_div2 = cs.reshape(-1).contiguous().clone(), s0
        norm2.pot = int(bool(norm2.pot) and is_pot(cs))
        return NS(
            norm1=self.ln_plan(self.act(in_name)[0], P[pre + '.norm1.weight'], P[pre + '.norm1.bias'], pre + '.qact1'),
            qkv=self.linear(pre + '.attn.qkv', P[pre + '.attn.qkv.weight'], bits[0], pre + '.qact1', pre + '.attn.qact1'),
            attn=self.window_attention(pre + '.attn', heads, res, ws, shift),
            proj=self.linear(pre + '.attn.proj', P[pre + '.attn.proj.weight'], bits[1], pre + '.attn.qact3',
                             pre + '.attn.qact4', residual=(in_name, pre + '.qact2')),
            norm2=norm2,
            fc1=self.linear(pre + '.mlp.fc1', P[pre + '.mlp.fc1.weight'] * cs.reshape(1, -1), bits[2], pre + '.mlp.qact0',
                            pre + '.mlp.qact1', gelu=True),
            fc2=self.linear(pre + '.mlp.fc2', P[pre + '.mlp.fc2.weight'], bits[3], pre + '.mlp.qact1', pre + '.mlp.qact2',
                            residual=(pre + '.qact2', pre + '.qact4')))

    def merge(self, pre, in_name, res, bit):
        P = self.P
        H, W = res
        tok = torch.arange(H * W, dtype=torch.int32).view(H, W)
        idx = torch.stack([tok[0::2, 0::2], tok[1::2, 0::2], tok[0::2, 1::2], tok[1::2, 1::2]], -1).reshape(-1).contiguous()
        in_scale = _expand(self.act(in_name)[0], P[pre + '.norm.weight'].numel() // 4).repeat(4)
        self.P[pre + '.reduction.bias'] = torch.zeros(P[pre + '.reduction.weight'].shape[0])
        return NS(idx=idx, norm=self.ln_plan(in_scale, P[pre + '.norm.weight'], P[pre + '.norm.bias'], pre + '.qact1'),
                  reduction=self.linear(pre + '.reduction', P[pre + '.reduction.weight'], bit, pre + '.qact1', pre + '.qact2'))

    def build(self):
        P, arch, bits = self.P, self.arch, self.bits
        in_s, in_z, _, _ = self.act('qact_input')
        plan = NS(arch=arch, bit_config=tuple(bits), input_scale=float(_scalar(in_s, 'qact_input')), input_zp=0.0,
                  patch_embed=self.linear('patch_embed.proj', P['patch_embed.proj.weight'], bits[0], 'qact_input',
                                          'patch_embed.qact_before_norm'),
                  pe_norm=self.ln_plan(self.act('patch_embed.qact_before_norm')[0], P['patch_embed.norm.weight'],
                                       P['patch_embed.norm.bias'], 'patch_embed.qact'),
                  stages=[])
        in_name = 'patch_embed.qact'
        grid = arch['img_size'] // arch['patch_size']
        pos = 1
        for i, depth in enumerate(arch['depths']):
            res = (grid // 2 ** i, grid // 2 ** i)
            st = NS(res=res, dim=arch['embed_dim'] * 2 ** i, heads=arch['num_heads'][i], blocks=[], merge=None)
            for j in range(depth):
                pre = 'layers.%d.blocks.%d' % (i, j)
                ws, shift = arch['window_size'], (0 if j % 2 == 0 else arch['window_size'] // 2)
                if min(res) <= arch['window_size']:
                    ws, shift = min(res), 0
                st.blocks.append(self.block(pre, in_name, st.heads, res, ws, shift, bits[pos:pos + 4]))
                in_name = pre + '.qact4'
                pos += 4
            if i < len(arch['depths']) - 1:
                pre = 'layers.%d.downsample' % i
                st.merge = self.merge(pre, in_name, res, bits[pos])
                in_name = pre + '.qact2'
                pos += 1
            plan.stages.append(st)
        plan.norm = self.ln_plan(self.act(in_name)[0], P['norm.weight'], P['norm.bias'], 'qact2')
        plan.pool_in_scale = self.pot_scalar('qact2')
        plan.pool_out_scale = float(_scalar(self.act('qact3')[0], 'qact3'))
        plan.head = self.linear('head', P['head.weight'], bits[-1], 'qact3', 'act_out')
        return plan


def build_swin_plan(state, bit_config=None):
    """Integer plan of the quantized Swin forward for one bit_config (layout: [patch embed] + per stage (per block
    [qkv, proj, fc1, fc2] ... + [reduction] where the stage downsamples) + [head]; None = all 8 bits)."""
    return _SwinBuilder(state, bit_config).build()


# ---- serialised plans (SURVEY section 5 "checkpoint / resume": the reference keeps calibrated scales only as Python
# attributes; an integer plan is self-contained and runs without the float model) ------------------------------------
def _plan_linears(plan):
    yield plan, 'patch_embed'
    for st in plan.stages:
        for b in st.blocks:
            for f in ('qkv', 'proj', 'fc1', 'fc2'):
                yield b, f
        if st.merge is not None:
            yield st.merge, 'reduction'
    yield plan, 'head'


def save_swin_plan(plan, path, pack4=True):
    """Write a Swin plan to a compressed .npz file (same format as plan.save_plan); 4-bit layers int4-packed."""
    import copy
    from .plan import _flatten
    if pack4:
        plan = copy.deepcopy(plan)
        for owner, f in _plan_linears(plan):
            setattr(owner, f, getattr(owner, f).pack())
    flat = {}
    _flatten(plan, 'plan', flat)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in flat.items()})


def load_swin_plan(path):
    """Read a plan written by `save_swin_plan`."""
    from .plan import _unflatten
    with np.load(path) as z:
        return _unflatten({k: z[k] for k in z.files}, 'plan')


# ---- execution ----------------------------------------------------------------------------------------------------
class _Bound:
    """A plan's tensors on one device plus the C descriptors that point at them."""

    def __init__(self, plan, device):
        self.plan, self.device, self.keep = plan, device, []
        self.patch_embed = self.linear(plan.patch_embed)
        self.pe_norm = self.ln(plan.pe_norm)
        self.stages = []
        for st in plan.stages:
            blocks = [NS(norm1=self.ln(b.norm1), qkv=self.linear(b.qkv), attn=self.attn(b.attn), proj=self.linear(b.proj),
                         norm2=self.ln(b.norm2),
                         fc1=self.linear(b.fc1), fc2=self.linear(b.fc2), plan=b) for b in st.blocks]
            merge = None
            if st.merge is not None:
                merge = NS(idx=self.up(st.merge.idx), norm=self.ln(st.merge.norm), reduction=self.linear(st.merge.reduction))
            self.stages.append(NS(res=st.res, dim=st.dim, heads=st.heads, blocks=blocks, merge=merge))
        self.norm = self.ln(plan.norm)
        self.head = self.linear(plan.head)
        self.desc = self._descriptor()

    def _descriptor(self):
        """The plan as a p2v_swin_desc (include/p2v.h): what p2v_swin_forward walks."""
        a = self.plan.arch
        ld = lambda lin: _cabi.LinearDesc(w=lin.w.data_ptr(), n=lin.n, k=lin.k, epi=lin.epi)
        self._c_blocks, stages = [], (_cabi.SwinStageDesc * len(self.stages))()
        for i, st in enumerate(self.stages):
            blocks = (_cabi.SwinBlockDesc * len(st.blocks))()
            for j, b in enumerate(st.blocks):
                blocks[j].norm1, blocks[j].norm2 = b.norm1, b.norm2
                blocks[j].qkv, blocks[j].proj, blocks[j].fc1, blocks[j].fc2 = ld(b.qkv), ld(b.proj), ld(b.fc1), ld(b.fc2)
                blocks[j].attn = b.attn
            self._c_blocks.append(blocks)
            s = stages[i]
            s.height, s.width, s.dim, s.depth = st.res[0], st.res[1], st.dim, len(st.blocks)
            s.blocks = C.cast(blocks, C.POINTER(_cabi.SwinBlockDesc))
            s.has_merge = int(st.merge is not None)
            if st.merge is not None:
                s.merge_idx, s.merge_norm, s.reduction = st.merge.idx.data_ptr(), st.merge.norm, ld(st.merge.reduction)
        self._c_stages = stages
        d = _cabi.SwinDesc()
        d.img_size, d.patch_size, d.in_chans = a['img_size'], a['patch_size'], a['in_chans']
        d.embed_dim, d.num_stages, d.num_classes = a['embed_dim'], len(self.stages), a['num_classes']
        d.input_scale = self.plan.input_scale
        d.patch_embed, d.pe_norm = ld(self.patch_embed), self.pe_norm
        d.stages = C.cast(stages, C.POINTER(_cabi.SwinStageDesc))
        d.norm, d.head = self.norm, ld(self.head)
        d.pool_in_scale, d.pool_out_scale = self.plan.pool_in_scale, self.plan.pool_out_scale
        return d

    def up(self, t):
        if t is None:
            return None
        d = t.to(self.device).contiguous()
        self.keep.append(d)
        return d

    def p(self, t):
        d = self.up(t)
        return None if d is None else d.data_ptr()

    def linear(self, lp):
        e = _cabi.Epilogue()
        e.acc_scale, e.bias = self.p(lp.acc_scale), self.p(lp.bias)
        e.out_scale, e.out_rscale = self.p(lp.out_scale), self.p(lp.out_rscale)
        e.res_scale, e.out2_scale = self.p(lp.res_scale), self.p(lp.out2_scale)
        e.out_zp, e.flags = lp.out_zp, lp.flags
        w = self.up(lp.codes())
        return NS(w=w, n=w.shape[0], k=w.shape[1], epi=e, residual=lp.res_scale is not None)

    def ln(self, p):
        d = _cabi.LayerNorm()
        d.in_mask, d.gamma, d.beta = self.p(p.in_mask), self.p(p.gamma), self.p(p.beta)
        d.ln_out_scale, d.ln_out_rscale = self.p(p.ln_out_scale), self.p(p.ln_out_rscale)
        d.post_mul, d.post_div1 = self.p(p.post_mul), self.p(p.post_div1)
        d.post_div2, d.post_zp, d.in_scale1, d.pot = p.post_div2, p.post_zp, p.in_scale1, p.pot
        d.pre_clamp = int(getattr(p, 'pre_clamp', 0))
        return d

    def attn(self, a):
        d = _cabi.WindowAttention()
        d.perm, d.region, d.bias, d.exp_lut = self.p(a.perm), self.p(a.region), self.p(a.bias), self.p(a.exp_lut)
        d.r3, d.exp_lut64, d.qk_scale, d.err_mul = self.p(a.r3), self.p(a.exp_lut64), a.qk_scale, a.err_mul
        d.lut_n = a.exp_lut.numel()
        d.n, d.heads, d.windows, d.tokens, d.channels = a.n, a.heads, a.windows, a.tokens, a.channels
        d.qshift, d.qscale, d.acc_scale = a.qshift, a.qscale, a.acc_scale
        d.a1_scale, d.a1_rscale, d.a2_rscale, d.mask_int = a.a1_scale, a.a1_rscale, a.a2_rscale, a.mask_int
        d.out_unit, d.out_rscale, d.softmax_levels = a.out_unit, a.out_rscale, a.levels
        return d


class SwinIntegerEngine:
    """Quantized forward of one calibrated SwinTransformer on one CUDA device."""

    def __init__(self, model=None, state=None, device=None, max_plans=4, plans=()):
        """`plans`: ready-made plans (`load_swin_plan`), served for their bit_config without the float model."""
        plans = list(plans)
        if state is None and model is not None:
            from .swin_quant import extract_swin_state
            state = extract_swin_state(model)
        if state is None and not plans:
            raise ValueError('SwinIntegerEngine needs a calibrated model, its state, or serialised plans')
        if device is None:
            device = next(model.parameters()).device if model is not None else torch.device('cuda')
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _cabi.P2VError('the integer engine runs on a CUDA device (sm_100a); got %s' % self.device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        _cabi.lib()          # fails loudly when the library is missing: there is no other implementation
        self.state = state
        self.arch = state['arch'] if state is not None else plans[0].arch
        self.max_plans = max(max_plans, len(plans))
        self._ready = {tuple(int(b) for b in p.bit_config): p for p in plans}
        self._bound = {}      # bit_config tuple -> _Bound
        self._buf = {}        # (name, shape) -> tensor
        self._graphs = {}     # (bit_config, batch) -> (graph, x_static, logits_static)
        self.launches = 0

    # -- plans ---------------------------------------------------------------------------------------------------------
    def bound(self, bit_config=None):
        n = num_linear_layers(self.arch)
        key = tuple([8] * n if bit_config is None else [int(b) for b in bit_config][:n])
        if key not in self._bound:
            while len(self._bound) >= self.max_plans:
                old = next(iter(self._bound))
                del self._bound[old]
                self._graphs = {k: v for k, v in self._graphs.items() if k[0] != old}
            if key in self._ready:
                plan = self._ready[key]
            elif self.state is not None:
                plan = build_swin_plan(self.state, key)
            else:
                raise KeyError('no serialised plan for bit_config %s and no calibrated state to build one from' % (key,))
            self._bound[key] = _Bound(plan, self.device)
        return key, self._bound[key]

    def buf(self, name, *shape, dtype=torch.int8):
        key = (name, shape, dtype)
        if key not in self._buf:
            self._buf[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return self._buf[key]

    # -- launches ------------------------------------------------------------------------------------------------------
    def _gemm(self, a, lin, out, m, residual=None, aux=None, out_f32=None):
        e = lin.epi
        flags = e.flags & ~(_cabi.EPI_RESIDUAL | _cabi.EPI_OUT_F32)
        e.residual = residual.data_ptr() if residual is not None else None
        e.aux_codes = aux.data_ptr() if aux is not None else None
        e.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        e.flags = flags | (_cabi.EPI_RESIDUAL if residual is not None else 0) | (_cabi.EPI_OUT_F32 if out_f32 is not None else 0)
        _cabi.check(self._lib.p2v_gemm_i8(a.data_ptr(), lin.k, lin.w.data_ptr(), out.data_ptr(), lin.n, m, lin.n, lin.k,
                                          C.byref(e), self._stream))
        self.launches += 1

    def _ln(self, x, desc, out, rows, d, ln_codes=None):
        _cabi.check(self._lib.p2v_layernorm_int(x.data_ptr(), d, out.data_ptr(), _cabi.ptr(ln_codes), rows, d,
                                                C.byref(desc), self._stream))
        self.launches += 1

    def _run(self, b, x, logits, logit_codes, dump=None):
        """The launch sequence of one forward of `b` images (x fp32 [b, c, h, w] on the device)."""
        lib = self._lib = _cabi.lib()
        st = self._stream = _cabi.current_stream(self.device)
        self.launches = 0
        bound = self._cur
        a = self.arch
        P, cin = a['patch_size'], a['in_chans']
        grid = a['img_size'] // P
        L, Cd = grid * grid, a['embed_dim']
        put = (lambda k, t: dump.__setitem__(k, t.clone())) if dump is not None else (lambda k, t: None)
        i32 = lambda name, *shape: self.buf(name, *shape, dtype=torch.int32) if dump is not None else None

        patches = self.buf('patches', b * L, cin * P * P)
        _cabi.check(lib.p2v_quant_patchify(x.data_ptr(), patches.data_ptr(), b, cin, a['img_size'], a['img_size'], P,
                                           bound.plan.input_scale, 0.0, st))
        self.launches += 1
        if dump is not None:
            dump['act/qact_input'] = patches.view(b, grid, grid, cin, P, P).permute(0, 3, 1, 4, 2, 5).reshape(
                b, cin, a['img_size'], a['img_size']).clone()
        pe = self.buf('pe', b * L, Cd)
        self._gemm(patches, bound.patch_embed, pe, b * L)
        put('act/patch_embed.qact_before_norm', pe.view(b, L, Cd))
        xs = self.buf('stream0', b * L, Cd)
        lnc = i32('lnc', b * L, Cd)
        self._ln(pe, bound.pe_norm, xs, b * L, Cd, lnc)
        if dump is not None:
            put('ln/patch_embed.norm', lnc.view(b, L, Cd))
            put('act/patch_embed.qact', xs.view(b, L, Cd))

        for si, stg in enumerate(bound.stages):
            H, W = stg.res
            L, Cd = H * W, stg.dim
            rows = b * L
            for bi, blk in enumerate(stg.blocks):
                pre = 'layers.%d.blocks.%d' % (si, bi)
                ap = blk.plan.attn
                y = self.buf('ln_out', rows, Cd)
                lnc = i32('lnc', rows, Cd)
                self._ln(xs, blk.norm1, y, rows, Cd, lnc)
                qkv = self.buf('qkv', rows, 3 * Cd)
                self._gemm(y, blk.qkv, qkv, rows)
                att = self.buf('att', rows, Cd)
                wa = blk.attn
                nw = b * ap.windows
                d1 = d2 = d3 = None
                if dump is not None:
                    d1 = self.buf('dump_a1', nw, ap.heads, ap.n, ap.n)
                    d2 = self.buf('dump_a2', nw, ap.heads, ap.n, ap.n)
                    d3 = self.buf('dump_sm', nw, ap.heads, ap.n, ap.n, dtype=torch.uint8)
                wa.dump_a1, wa.dump_a2, wa.dump_softmax = _cabi.ptr(d1), _cabi.ptr(d2), _cabi.ptr(d3)
                _cabi.check(lib.p2v_window_attention_int(qkv.data_ptr(), att.data_ptr(), b, C.byref(wa), st))
                self.launches += 1
                x1 = self.buf('stream_mid', rows, Cd)
                aux = self.buf('aux', rows, Cd) if dump is not None else None
                self._gemm(att, blk.proj, x1, rows, residual=xs, aux=aux)
                if dump is not None:
                    perm = ap.perm.to(self.device, torch.long)
                    win = lambda t, c: t.view(b, L, c)[:, perm, :].reshape(nw, ap.n, c)
                    put('ln/' + pre + '.norm1', lnc.view(b, L, Cd))
                    put('act/' + pre + '.qact1', y.view(b, L, Cd))
                    put('act/' + pre + '.attn.qact1', win(qkv, 3 * Cd))
                    put('act/' + pre + '.attn.qact_attn1', d1)
                    dump['act/' + pre + '.attn.qact_table'] = ap.table_codes.clone()
                    put('act/' + pre + '.attn.qact2', d2)
                    put('softmax/' + pre + '.attn.log_int_softmax', d3)
                    put('act/' + pre + '.attn.qact3', win(att, Cd))
                    put('act/' + pre + '.attn.qact4', win(aux, Cd))
                    put('act/' + pre + '.qact2', x1.view(b, L, Cd))
                m0 = self.buf('ln_out', rows, Cd)
                self._ln(x1, blk.norm2, m0, rows, Cd, lnc)      # LN2 + qact3 + / channel scale + mlp.qact0
                hid = self.buf('hidden', rows, blk.fc1.n)
                self._gemm(m0, blk.fc1, hid, rows)
                xn = self.buf('stream%d' % ((bi + 1) % 2), rows, Cd)
                self._gemm(hid, blk.fc2, xn, rows, residual=x1, aux=aux)
                if dump is not None:
                    put('ln/' + pre + '.norm2', lnc.view(b, L, Cd))
                    put('act/' + pre + '.qact3', lnc.view(b, L, Cd).clamp(-128, 127))
                    put('act/' + pre + '.mlp.qact0', m0.view(b, L, Cd))
                    put('act/' + pre + '.mlp.qact1', hid.view(b, L, blk.fc1.n))
                    put('act/' + pre + '.mlp.qact2', aux.view(b, L, Cd))
                    put('act/' + pre + '.qact4', xn.view(b, L, Cd))
                xs = xn
            if stg.merge is not None:
                pre = 'layers.%d.downsample' % si
                Lo = L // 4
                cat = self.buf('merge_cat', b * Lo, 4 * Cd)
                _cabi.check(lib.p2v_gather_row_segments(xs.data_ptr(), cat.data_ptr(), stg.merge.idx.data_ptr(), b, L, Lo,
                                                        4, Cd, st))
                self.launches += 1
                y = self.buf('merge_ln', b * Lo, 4 * Cd)
                lnc = i32('lnc', b * Lo, 4 * Cd)
                self._ln(cat, stg.merge.norm, y, b * Lo, 4 * Cd, lnc)
                xn = self.buf('stream0', b * Lo, 2 * Cd)
                self._gemm(y, stg.merge.reduction, xn, b * Lo)
                if dump is not None:
                    put('ln/' + pre + '.norm', lnc.view(b, Lo, 4 * Cd))
                    put('act/' + pre + '.qact1', y.view(b, Lo, 4 * Cd))
                    put('act/' + pre + '.qact2', xn.view(b, Lo, 2 * Cd))
                xs = xn
        rows = b * L
        y = self.buf('ln_out', rows, Cd)
        lnc = i32('lnc', rows, Cd)
        self._ln(xs, bound.norm, y, rows, Cd, lnc)
        pooled = self.buf('pooled', b, Cd)
        _cabi.check(lib.p2v_avgpool_requant(y.data_ptr(), pooled.data_ptr(), b, L, Cd, bound.plan.pool_in_scale,
                                            bound.plan.pool_out_scale, 0.0, st))
        self.launches += 1
        self._gemm(pooled, bound.head, logit_codes, b, out_f32=logits)
        if dump is not None:
            put('ln/norm', lnc.view(b, L, Cd))
            put('act/qact2', y.view(b, L, Cd))
            put('act/qact3', pooled.view(b, Cd, 1))
            put('act/act_out', logit_codes)

    # -- public --------------------------------------------------------------------------------------------------------
    def _prepare(self, x):
        if not (torch.is_tensor(x) and x.is_cuda):
            raise _cabi.P2VError('SwinIntegerEngine.forward expects a CUDA tensor (there is no CPU implementation)')
        a = self.arch
        if tuple(x.shape[1:]) != (a['in_chans'], a['img_size'], a['img_size']):
            raise ValueError('expected [b, %d, %d, %d] images, got %s' % (a['in_chans'], a['img_size'], a['img_size'],
                                                                          tuple(x.shape)))
        return x.to(self.device, torch.float32).contiguous()

    def forward(self, x, bit_config=None, graph=True):
        """logits fp32 [b, classes] of the quantized forward.  Launches are captured into a CUDA graph per
        (bit_config, batch) after one eager run and replayed from then on."""
        x = self._prepare(x)
        b = x.shape[0]
        key, bound = self.bound(bit_config)
        self._cur = bound
        nc = self.arch['num_classes']
        with torch.cuda.device(self.device):
            gk = (key, b)
            if graph and gk in self._graphs:
                g, xs, logits = self._graphs[gk]
                xs.copy_(x)
                g.replay()
                return logits.clone()
            logits = torch.empty(b, nc, dtype=torch.float32, device=self.device)
            codes = self.buf('logit_codes', b, nc)
            self._c_forward(bound, b, x, logits, codes)
            if graph:
                xs = x.clone()
                out = torch.empty_like(logits)
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._c_forward(bound, b, xs, out, codes)
                self._graphs[gk] = (g, xs, out)
            return logits

    def forward_u8(self, x_u8, mean, std, bit_config=None, graph=True):
        """The forward from 8-bit pixels [b, c, h, w] (CUDA uint8) with the loader's normalisation constants: the device
        evaluates (pixel / 255 - mean[c]) / std[c] op for op (p2v_swin_forward_u8), so the logits equal `forward` on the
        normalised fp32 tensor, for a quarter of the input traffic."""
        a = self.arch
        if not (torch.is_tensor(x_u8) and x_u8.is_cuda and x_u8.dtype == torch.uint8):
            raise _cabi.P2VError('SwinIntegerEngine.forward_u8 expects a CUDA uint8 tensor')
        if tuple(x_u8.shape[1:]) != (a['in_chans'], a['img_size'], a['img_size']):
            raise ValueError('expected [b, %d, %d, %d] images, got %s' % (a['in_chans'], a['img_size'], a['img_size'],
                                                                          tuple(x_u8.shape)))
        x_u8 = x_u8.to(self.device).contiguous()
        b, c = x_u8.shape[0], x_u8.shape[1]
        key, bound = self.bound(bit_config)
        norm = ((C.c_float * c)(*[float(v) for v in mean]), (C.c_float * c)(*[float(v) for v in std]))
        nc = a['num_classes']
        with torch.cuda.device(self.device):
            gk = (key, b, 'u8', tuple(float(v) for v in mean), tuple(float(v) for v in std))
            if graph and gk in self._graphs:
                g, xs, logits = self._graphs[gk]
                xs.copy_(x_u8)
                g.replay()
                return logits.clone()
            logits = torch.empty(b, nc, dtype=torch.float32, device=self.device)
            codes = self.buf('logit_codes', b, nc)
            self._c_forward(bound, b, x_u8, logits, codes, norm)
            if graph:
                xs = x_u8.clone()
                out = torch.empty_like(logits)
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._c_forward(bound, b, xs, out, codes, norm)
                self._graphs[gk] = (g, xs, out)
            return logits

    def _c_forward(self, bound, b, x, logits, codes, norm=None):
        """One p2v_swin_forward call (csrc/p2v_swin.cu): the launch sequence of `_run` without the dumps, on the
        current stream, over one workspace per batch size."""
        lib = _cabi.lib()
        nbytes = lib.p2v_swin_workspace_bytes(C.byref(bound.desc), b)
        if nbytes < 0:
            _cabi.check(-1)
        ws = self.buf('workspace', nbytes + 1024, dtype=torch.uint8)
        ptr = ws.data_ptr() + (-ws.data_ptr()) % 1024
        st = _cabi.current_stream(self.device)
        if norm is not None:       # 8-bit pixels + (mean, std) as host float arrays
            _cabi.check(lib.p2v_swin_forward_u8(C.byref(bound.desc), x.data_ptr(), norm[0], norm[1], logits.data_ptr(),
                                                codes.data_ptr(), b, ptr, st))
        else:
            _cabi.check(lib.p2v_swin_forward(C.byref(bound.desc), x.data_ptr(), logits.data_ptr(), codes.data_ptr(), b, ptr, st))
        self.launches = lib.p2v_swin_launches_per_forward(C.byref(bound.desc))

    def forward_dump(self, x, bit_config=None):
        """(logits, {golden-style key: integer codes on the CPU}) - every quantizer's codes in the layout of
        oracle/swin_fakequant_forward.py (window order for the attention internals)."""
        x = self._prepare(x)
        b = x.shape[0]
        key, bound = self.bound(bit_config)
        self._cur = bound
        dump = {}
        with torch.cuda.device(self.device):
            logits = torch.empty(b, self.arch['num_classes'], dtype=torch.float32, device=self.device)
            codes = self.buf('logit_codes', b, self.arch['num_classes'])
            self._run(b, x, logits, codes, dump)
            torch.cuda.current_stream(self.device).synchronize()
        return logits, {k: v.to('cpu', torch.int32) for k, v in dump.items()}
