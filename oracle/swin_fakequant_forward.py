"""CPU restatement of the reference's quantized Swin forward (BASELINE config 5).  TEST INFRASTRUCTURE ONLY.

Follows models/swin_quant.py (WindowAttention.forward :177-221, SwinTransformerBlock.forward :345-399,
PatchMerging.forward :445-467, SwinTransformer.forward_features/forward :790-817) with the quantizer / integer
LayerNorm / log-int-softmax arithmetic shared with the ViT oracle (oracle/fakequant_forward.py), same ATen ops in
the same order.  The reference file does not run as shipped (four stale call sites, see
tests/golden/make_golden_swin.py); this restatement is pinned against the reference run with those call-plumbing shims:
tests/golden/swin_micro.npz, tests/test_swin_golden.py.

Input: the "quant state" of diff_vit_b200.swin_quant.extract_swin_state (plain tensors):
  arch   : img_size, patch_size, num_classes, embed_dim, depths, num_heads, window_size, softmax_bits
  params : float parameters / buffers under the reference's state_dict names
  act    : {qact name: (scale, zero_point, qmin, qmax)}
  weight : {layer name: {'int8': (scale, zero_point), ...}}
  cs     : {'layers.i.blocks.j.mlp': SmoothQuant channel scale}
"""
import torch
import torch.nn.functional as F

from . import fakequant_forward as base
from .fakequant_forward import Trace, _qact, _qweight, int_layernorm, log_int_softmax, _linear, _conv2d, _matmul


def _window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _window_reverse(windows, ws, H, W):
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _ln(state, trace, key, x, in_name, out_name, weight, bias, expand=1):
    """QIntLayerNorm in 'int' mode (ptq/layers.py:255-289); `expand` tiles the input scale (PatchMerging)."""
    in_scale = state['act'][in_name][0]
    if expand != 1:
        in_scale = in_scale.unsqueeze(-1).expand(-1, expand).T.reshape(-1)
    out_scale = state['act'][out_name][0]
    codes, _ = int_layernorm(x, in_scale, out_scale, weight, bias)
    return trace.put('ln/' + key, codes) * out_scale.reshape(1, 1, -1)


def _window_attention(state, pre, x, mask, heads, ws, bits, trace):
    """swin_quant.py:177-221"""
    P = state['params']
    B_, N, C = x.shape
    w = _qweight(state, pre + '.qkv', P[pre + '.qkv.weight'], bits[0], 'linear')
    x = _linear(x, w, P[pre + '.qkv.bias'])
    x = _qact(state, pre + '.qact1', x, trace)
    qkv = x.reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * ((C // heads) ** -0.5)
    attn = _matmul(q, k.transpose(-2, -1))
    attn = _qact(state, pre + '.qact_attn1', attn, trace)
    table_q = _qact(state, pre + '.qact_table', P[pre + '.relative_position_bias_table'], trace)
    idx = P[pre + '.relative_position_index'].long()
    bias = table_q[idx.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    attn = _qact(state, pre + '.qact2', attn, trace)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    codes, attn_p = log_int_softmax(attn, state['act'][pre + '.qact2'][0], state['arch']['softmax_bits'])
    used = trace.put('softmax/' + pre + '.log_int_softmax', codes)
    if used is not codes:
        attn_p = torch.where(used >= 2 ** state['arch']['softmax_bits'], torch.zeros_like(used), 2 ** (-used))
    x = _matmul(attn_p, v).transpose(1, 2).reshape(B_, N, C)
    x = _qact(state, pre + '.qact3', x, trace)
    w = _qweight(state, pre + '.proj', P[pre + '.proj.weight'], bits[1], 'linear')
    x = _linear(x, w, P[pre + '.proj.bias'])
    return _qact(state, pre + '.qact4', x, trace)


def _mlp(state, pre, x, bits, trace):
    """models/layers_quant.py:304-346 (calibrated branch), as in the ViT oracle."""
    P = state['params']
    cs = state['cs'][pre]
    x = _qact(state, pre + '.qact0', x / cs.reshape(1, 1, -1), trace)
    w = _qweight(state, pre + '.fc1', P[pre + '.fc1.weight'] * cs.reshape(1, -1), bits[0], 'linear')
    x = _linear(x, w, P[pre + '.fc1.bias'])
    x = F.gelu(x)
    x = _qact(state, pre + '.qact1', x, trace)
    w = _qweight(state, pre + '.fc2', P[pre + '.fc2.weight'], bits[1], 'linear')
    x = _linear(x, w, P[pre + '.fc2.bias'])
    return _qact(state, pre + '.qact2', x, trace)


def _block(state, pre, x, in_name, res, heads, ws, shift, bits, trace):
    """swin_quant.py:345-399"""
    P = state['params']
    H, W = res
    B, L, C = x.shape
    shortcut = x
    x = _ln(state, trace, pre + '.norm1', x, in_name, pre + '.qact1', P[pre + '.norm1.weight'], P[pre + '.norm1.bias'])
    x = _qact(state, pre + '.qact1', x, trace)
    x = x.view(B, H, W, C)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = _window_partition(x, ws).view(-1, ws * ws, C)
    mask = P.get(pre + '.attn_mask') if shift > 0 else None
    aw = _window_attention(state, pre + '.attn', xw, mask, heads, ws, bits[0:2], trace)
    x = _window_reverse(aw.view(-1, ws, ws, C), ws, H, W)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x.view(B, H * W, C)
    x = _qact(state, pre + '.qact2', shortcut + x, trace)
    y = _ln(state, trace, pre + '.norm2', x, pre + '.qact2', pre + '.qact3', P[pre + '.norm2.weight'], P[pre + '.norm2.bias'])
    y = _qact(state, pre + '.qact3', y, trace)
    y = _mlp(state, pre + '.mlp', y, bits[2:4], trace)
    return _qact(state, pre + '.qact4', x + y, trace)


def _patch_merging(state, pre, x, in_name, res, bit, trace):
    """swin_quant.py:445-467"""
    P = state['params']
    H, W = res
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    x = x.view(B, -1, 4 * C)
    x = _ln(state, trace, pre + '.norm', x, in_name, pre + '.qact1', P[pre + '.norm.weight'], P[pre + '.norm.bias'], 4)
    x = _qact(state, pre + '.qact1', x, trace)
    w = _qweight(state, pre + '.reduction', P[pre + '.reduction.weight'], bit, 'linear')
    x = _linear(x, w, torch.zeros(w.shape[0]))
    return _qact(state, pre + '.qact2', x, trace)


def num_linear_layers(arch):
    n = 1
    for i, d in enumerate(arch['depths']):
        n += 4 * d + (1 if i < len(arch['depths']) - 1 else 0)
    return n + 1


def forward(state, x, bit_config=None, capture=False, accum='fp32', override=None):
    """Quantized Swin forward.  Returns (logits fp32 [B, classes], {key: int32 codes}).  bit_config: None = all 8
    bits; layout [patch embed] + per stage (per block qkv, proj, fc1, fc2; then the reduction) + [head]."""
    assert accum in ('fp32', 'fp64')
    prev, base._ACCUM = base._ACCUM, accum
    try:
        return _forward(state, x, bit_config, capture, override)
    finally:
        base._ACCUM = prev


def _forward(state, x, bit_config, capture, override):
    arch, P = state['arch'], state['params']
    bits = [8] * num_linear_layers(arch) if bit_config is None else list(bit_config)
    trace = Trace(capture, override)
    ws_cfg = arch['window_size']
    with torch.no_grad():
        x = _qact(state, 'qact_input', x, trace)
        w = _qweight(state, 'patch_embed.proj', P['patch_embed.proj.weight'], bits[0], 'conv')
        x = _conv2d(x, w, P['patch_embed.proj.bias'], arch['patch_size'])
        x = x.flatten(2).transpose(1, 2)
        x = _qact(state, 'patch_embed.qact_before_norm', x, trace)
        x = _ln(state, trace, 'patch_embed.norm', x, 'patch_embed.qact_before_norm', 'patch_embed.qact',
                P['patch_embed.norm.weight'], P['patch_embed.norm.bias'])
        x = _qact(state, 'patch_embed.qact', x, trace)
        in_name = 'patch_embed.qact'
        grid = arch['img_size'] // arch['patch_size']
        pos = 1
        for i, depth in enumerate(arch['depths']):
            res = (grid // 2 ** i, grid // 2 ** i)
            heads = arch['num_heads'][i]
            for j in range(depth):
                pre = 'layers.%d.blocks.%d' % (i, j)
                ws, shift = ws_cfg, (0 if j % 2 == 0 else ws_cfg // 2)
                if min(res) <= ws_cfg:
                    ws, shift = min(res), 0
                x = _block(state, pre, x, in_name, res, heads, ws, shift, bits[pos:pos + 4], trace)
                in_name = pre + '.qact4'
                pos += 4
            if i < len(arch['depths']) - 1:
                pre = 'layers.%d.downsample' % i
                x = _patch_merging(state, pre, x, in_name, res, bits[pos], trace)
                in_name = pre + '.qact2'
                pos += 1
        x = _ln(state, trace, 'norm', x, in_name, 'qact2', P['norm.weight'], P['norm.bias'])
        x = _qact(state, 'qact2', x, trace)
        x = F.adaptive_avg_pool1d(x.transpose(1, 2), 1)
        x = _qact(state, 'qact3', x, trace)
        x = torch.flatten(x, 1)
        w = _qweight(state, 'head', P['head.weight'], bits[-1], 'linear')
        x = _linear(x, w, P['head.bias'])
        x = _qact(state, 'act_out', x, trace)
    return x, trace.codes
