"""CPU restatement of the reference's quantized ("fake-quant") forward.  TEST INFRASTRUCTURE ONLY.

What it is: a functional, dependency-free (torch CPU only) restatement of what
`VisionTransformer.forward(x, bit_config)` of LeSN-Lab/diff-ViT computes after calibration and
`model_quant()` (reference: models/vit_fquant.py:700-799), i.e. fp32 tensors that are rounded to an
integer grid and multiplied back, fp32 `F.linear`/`F.conv2d`/`@`, the fp32 "integer" LayerNorm and the
log-int-softmax.  It uses the same ATen ops in the same order as the reference, so on the same CPU it
reproduces the reference bit for bit; it additionally returns the integer codes of every quantizer.

Pinning: the reference has no tests or golden vectors of its own, so this oracle
is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).

Input: a "quant state" dict (plain tensors, see `diff_vit_b200.plan.extract_state`):
  arch   : img_size, patch_size, in_chans, embed_dim, depth, num_heads, num_classes, attn_scale, softmax_bits
  params : float parameters under the reference's state_dict names
  act    : {qact name: (scale, zero_point, qmin, qmax)}
  weight : {layer name: {'int8': (scale, zero_point), 'int4': ...}}
  cs     : {'blocks.i.attn' | 'blocks.i.mlp': SmoothQuant channel scale}
"""
import torch
import torch.nn.functional as F

_W_RANGE = {4: (-8, 7), 8: (-128, 127)}

# Accumulation mode of every contraction and row reduction (F.linear, F.conv2d, q @ k^T, attn @ v, the LayerNorm
# row sums, the softmax row sum).  'fp32' is the reference as it runs (ATen's fp32 kernels, summation order of the
# host BLAS).  'fp64' evaluates the SAME fp32 operands with fp64 products and sums and rounds the result to fp32
# once: the reference's expression without the accumulation noise of one particular BLAS.  On power-of-two grids
# both modes give identical codes (the sums are exact either way); with float scales (percentile / omse / ema) the
# fp32 sums are inexact, and the 'fp64' mode is what an exact-accumulation integer implementation can be held to
# (tests/test_gpu_model.py config 3 measures oracle-fp32 vs oracle-fp64 next to kernel vs oracle-fp64).
_ACCUM = 'fp32'


def _wide():
    return _ACCUM == 'fp64'


def _linear(x, w, b):
    if _wide():
        return (x.double() @ w.double().t()).float() + b
    return F.linear(x, w, b)


def _conv2d(x, w, b, stride):
    if _wide():
        return F.conv2d(x.double(), w.double(), None, stride=stride).float() + b.reshape(1, -1, 1, 1)
    return F.conv2d(x, w, b, stride=stride)


def _matmul(a, b):
    if _wide():
        return (a.double() @ b.double()).float()
    return a @ b


def _rowsum(x):
    if _wide():
        return x.double().sum(dim=-1).float()
    return x.sum(dim=-1)


def _bshape(x, weight_kind=None):
    """Broadcast shape of a per-channel scale (reference: models/ptq/quantizer/base.py:14-33)."""
    if weight_kind == 'conv':
        return (-1, 1, 1, 1)
    if weight_kind == 'linear':
        return (-1, 1)
    return {2: (1, -1), 3: (1, 1, -1), 4: (1, -1, 1, 1)}[x.dim()]


def fake_quant(x, scale, zero_point, qmin, qmax, weight_kind=None):
    """codes = clamp(RNE(x/s + zp)); value = (codes - zp) * s  (models/ptq/quantizer/uniform.py:82-88,123-127)."""
    shape = _bshape(x, weight_kind)
    s = scale.reshape(shape)
    z = zero_point.reshape(shape)
    q = (x / s + z).round().clamp(qmin, qmax)
    return q, (q - z) * s


class Trace:
    """Collects the integer codes of every quantizer, keyed like the golden fixtures.

    override ("teacher forcing"): {key: codes}.  The oracle still computes and records ITS codes of such a layer from
    the inputs it was given, but continues with the supplied ones.  With the codes of an implementation under test
    supplied for every layer, each layer of the oracle is evaluated on exactly the inputs the implementation's layer
    saw, so a per-layer comparison stays meaningful over the whole depth: without it one admissible rounding-tie flip
    (device erf vs ATen erf in a GELU, one step of a float-scale division) is amplified by the following random-init
    blocks until the two runs have nothing to do with each other - the reference's own CPU and GPU runs diverge the
    same way."""

    def __init__(self, enabled, override=None):
        self.enabled = enabled
        self.codes = {}
        self.override = override or {}

    def put(self, key, q):
        if self.enabled:
            self.codes[key] = q.to(torch.int32)
        if key in self.override:
            o = torch.as_tensor(self.override[key]).to(q.dtype)
            if o.numel() == q.numel():
                return o.reshape(q.shape)
            if key == 'ln/norm':          # an implementation may normalise the CLS rows only
                q = q.clone()
                q[:, 0] = o.reshape(q.shape[0], q.shape[-1])
                return q
            raise ValueError('override %s: %d elements for a layer of %d' % (key, o.numel(), q.numel()))
        return q


def _qact(state, name, x, trace):
    scale, zp, qmin, qmax = state['act'][name]
    q, _ = fake_quant(x, scale, zp, qmin, qmax)
    q = trace.put('act/' + name, q)
    return (q - zp.reshape(_bshape(x))) * scale.reshape(_bshape(x))


def _qweight(state, name, w, bits, kind):
    """The reference re-quantizes the fp32 weight on every call (models/ptq/layers.py:86,177)."""
    scale, zp = state['weight'][name]['int%d' % bits]
    lo, hi = _W_RANGE[bits]
    return fake_quant(w, scale, zp, lo, hi, kind)[1]


def int_layernorm(x, in_scale, out_scale, gamma, beta):
    """fp32 'integer' LayerNorm (models/ptq/layers.py:255-289).  Returns (codes on the out grid, value)."""
    channel_nums = x.shape[-1]
    in_scale = in_scale.reshape(1, 1, -1)
    out_scale = out_scale.reshape(1, 1, -1)
    x_q = (x / in_scale).round()
    in_scale1 = in_scale.min()
    in_scale_mask = (in_scale / in_scale1).round()
    x_q = x_q * in_scale_mask
    mean_x_q = (_rowsum(x_q) / channel_nums if _wide() else x_q.mean(dim=-1)) * in_scale1
    std_x_q = (in_scale1 / channel_nums) * torch.sqrt(
        channel_nums * _rowsum(x_q ** 2) - _rowsum(x_q) ** 2)
    A = (in_scale1 / std_x_q).unsqueeze(-1) * gamma.reshape(1, 1, -1) / out_scale
    A_sign = A.sign()
    # dyadic approximation A ~ M / 2^N (get_MN, layers.py:234-238)
    A_abs = A.abs()
    N = torch.clamp(7 - torch.floor(torch.log2(A_abs)), 0, 31)
    M = torch.clamp(torch.floor(A_abs * torch.pow(2, N)), 0, 255)
    B = ((beta.reshape(1, 1, -1) - (mean_x_q / std_x_q).unsqueeze(-1) * gamma.reshape(1, 1, -1)) /
         out_scale * torch.pow(2, N)).round()
    x_q = ((A_sign * M * x_q + B) / torch.pow(2, N)).round()
    return x_q, x_q * out_scale


def softmax_exp_constants(scale):
    """x0_int, b_int, c_int of the I-BERT integer exp (layers.py:334-352)."""
    a, b, c = 0.35815147, 0.96963238, 1.
    b /= a
    c /= a
    return torch.floor(-0.6931 / scale), torch.floor(b / scale), torch.floor(c / scale ** 2)


def log_int_softmax(x, scale, bits):
    """log-int-softmax (layers.py:323-376).  Returns (log2 codes with 2^bits meaning 'zero', value)."""
    n = 32
    x0_int, b_int, c_int = softmax_exp_constants(scale)
    x_int = x / scale
    x_int = x_int - x_int.max(dim=-1, keepdim=True).values
    x_int = torch.max(x_int, n * x0_int)
    q = torch.floor(x_int / x0_int)
    r = x_int - x0_int * q
    z = r + b_int
    z = r * z
    z = z + c_int
    exp_int = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
    exp_int_sum = _rowsum(exp_int).unsqueeze(-1)
    softmax_out = torch.round(exp_int_sum / exp_int)
    big = softmax_out.log2().floor()
    extra = (softmax_out - 2 ** big) >= 2 ** (big - 1)
    big[extra] = big[extra] + 1
    levels = 2 ** bits
    mask = big >= levels
    qlog = torch.clamp(big, 0, levels - 1)
    deq = 2 ** (-qlog)
    deq[mask] = 0
    codes = qlog.clone()
    codes[mask] = levels
    return codes, deq


def _attention(state, i, x, bits_qkv, bits_proj, trace):
    """models/vit_fquant.py:281-338 (calibrated branch)."""
    pre = 'blocks.%d.attn' % i
    P = state['params']
    arch = state['arch']
    B, N, C = x.shape
    H = arch['num_heads']
    cs = state['cs'][pre]
    x = _qact(state, pre + '.qact0', x / cs.reshape(1, 1, -1), trace)
    w = _qweight(state, pre + '.qkv', P[pre + '.qkv.weight'] * cs.reshape(1, -1), bits_qkv, 'linear')
    x = _linear(x, w, P[pre + '.qkv.bias'])
    x = _qact(state, pre + '.qact1', x, trace)
    qkv = x.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = _matmul(q, k.transpose(-2, -1)) * arch['attn_scale']
    attn = _qact(state, pre + '.qact_attn1', attn, trace)
    codes, attn = log_int_softmax(attn, state['act'][pre + '.qact_attn1'][0], arch['softmax_bits'])
    used = trace.put('softmax/' + pre + '.log_int_softmax', codes)
    if used is not codes:
        attn = torch.where(used >= 2 ** arch['softmax_bits'], torch.zeros_like(used), 2 ** (-used))
    x = _matmul(attn, v).transpose(1, 2).reshape(B, N, C)
    x = _qact(state, pre + '.qact2', x, trace)
    w = _qweight(state, pre + '.proj', P[pre + '.proj.weight'], bits_proj, 'linear')
    x = _linear(x, w, P[pre + '.proj.bias'])
    return _qact(state, pre + '.qact3', x, trace)


def _mlp(state, i, x, bits_fc1, bits_fc2, trace):
    """models/layers_quant.py:304-346 (calibrated branch)."""
    pre = 'blocks.%d.mlp' % i
    P = state['params']
    cs = state['cs'][pre]
    x = _qact(state, pre + '.qact0', x / cs.reshape(1, 1, -1), trace)
    w = _qweight(state, pre + '.fc1', P[pre + '.fc1.weight'] * cs.reshape(1, -1), bits_fc1, 'linear')
    x = _linear(x, w, P[pre + '.fc1.bias'])
    x = F.gelu(x)
    x = _qact(state, pre + '.qact1', x, trace)
    w = _qweight(state, pre + '.fc2', P[pre + '.fc2.weight'], bits_fc2, 'linear')
    x = _linear(x, w, P[pre + '.fc2.bias'])
    return _qact(state, pre + '.qact2', x, trace)


def forward(state, x, bit_config, capture=False, accum='fp32', override=None):
    """Quantized forward.  Returns (logits fp32 [B, classes], {key: int32 codes}).

    accum: 'fp32' (the reference as it runs) or 'fp64' (same operands, exact-to-fp64 sums; see _ACCUM).
    override: teacher forcing, see Trace.

    bit_config index map: 0 = patch-embed conv, 1+4i..4+4i = block i qkv/proj/fc1/fc2,
    -1 = head."""
    global _ACCUM
    assert accum in ('fp32', 'fp64')
    prev, _ACCUM = _ACCUM, accum
    try:
        return _forward(state, x, bit_config, capture, override)
    finally:
        _ACCUM = prev


def _forward(state, x, bit_config, capture, override=None):
    arch, P = state['arch'], state['params']
    trace = Trace(capture, override)
    with torch.no_grad():
        B = x.shape[0]
        x = _qact(state, 'qact_input', x, trace)
        w = _qweight(state, 'patch_embed.proj', P['patch_embed.proj.weight'], bit_config[0], 'conv')
        x = _conv2d(x, w, P['patch_embed.proj.bias'], arch['patch_size'])
        x = x.flatten(2).transpose(1, 2)
        x = _qact(state, 'patch_embed.qact', x, trace)
        x = torch.cat((P['cls_token'].expand(B, -1, -1), x), dim=1)
        x = _qact(state, 'qact_embed', x, trace)
        x = x + _qact(state, 'qact_pos', P['pos_embed'], trace)
        x = _qact(state, 'qact1', x, trace)
        in_name = 'qact1'
        for i in range(arch['depth']):
            b = bit_config[4 * i + 1:4 * i + 5]
            pre = 'blocks.%d' % i
            cs_attn = state['cs'][pre + '.attn']
            out_scale = state['act'][pre + '.attn.qact0'][0] * cs_attn
            codes, y = int_layernorm(x, state['act'][in_name][0], out_scale, P[pre + '.norm1.weight'],
                                     P[pre + '.norm1.bias'])
            y = trace.put('ln/' + pre + '.norm1', codes) * out_scale.reshape(1, 1, -1)
            y = _attention(state, i, y, b[0], b[1], trace)
            x = _qact(state, pre + '.qact2', x + y, trace)
            # norm2 is given the ATTENTION SmoothQuant scale (vit_fquant.py:464)
            out_scale = state['act'][pre + '.mlp.qact0'][0] * cs_attn
            codes, y = int_layernorm(x, state['act'][pre + '.qact2'][0], out_scale, P[pre + '.norm2.weight'],
                                     P[pre + '.norm2.bias'])
            y = trace.put('ln/' + pre + '.norm2', codes) * out_scale.reshape(1, 1, -1)
            y = _mlp(state, i, y, b[2], b[3], trace)
            x = _qact(state, pre + '.qact4', x + y, trace)
            in_name = pre + '.qact4'
        codes, x = int_layernorm(x, state['act'][in_name][0], state['act']['qact2'][0], P['norm.weight'],
                                 P['norm.bias'])
        x = trace.put('ln/norm', codes) * state['act']['qact2'][0].reshape(1, 1, -1)
        x = _qact(state, 'qact2', x[:, 0], trace)
        w = _qweight(state, 'head', P['head.weight'], bit_config[-1], 'linear')
        x = _linear(x, w, P['head.bias'])
        x = _qact(state, 'act_out', x, trace)
    return x, trace.codes
