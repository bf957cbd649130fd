"""Per-module quantized execution (a Q-module called on its own after ``model_quant()``).

The fast path of this package is the whole-graph integer engine; this module keeps the individual
operators usable the way the reference's analysis scripts use them (forward hooks, hand-built graphs,
a bit_config that leaves some layers in fp32): fp32 tensors in, fp32 dequantized tensors out, the
quantization arithmetic in the sm_100a kernels of csrc/p2v_modules.cu and p2v_rowops.cu.  The fp32 matrix
products of QLinear / QConv2d stay library calls, exactly the products the reference itself forms.
"""
import torch
from torch.nn import functional as F

from . import _cabi


def _per_channel(t, n, device):
    t = torch.as_tensor(t, dtype=torch.float32, device=device).reshape(-1)
    return (t.expand(n) if t.numel() == 1 else t).contiguous()


def qact(module, x):
    """QAct.forward in quantized mode (reference: models/ptq/layers.py:217-220)."""
    q = module.quantizer
    x = x.contiguous().float()
    if x.dim() == 4:
        channels, inner = x.shape[1], x.shape[2] * x.shape[3]
    elif x.dim() in (2, 3):
        channels, inner = x.shape[-1], 1
    else:
        raise NotImplementedError
    outer = x.numel() // (channels * inner)
    scale = _per_channel(q.scale, channels, x.device)
    zp = _per_channel(q.zero_point, channels, x.device)
    out = torch.empty_like(x)
    _cabi.check(_cabi.lib().p2v_fake_quant_f32(x.data_ptr(), out.data_ptr(), None, outer, channels, inner,
                                              scale.data_ptr(), zp.data_ptr(), q.bit_type.lower_bound,
                                              q.bit_type.upper_bound, _cabi.current_stream()))
    return out


def _fake_quant_weight(module, weight):
    """quantizer(weight) of QLinear / QConv2d (models/ptq/layers.py:95,176): per-out-channel (or scalar)
    scale of the currently selected bit type, fake-quantized by the sm_100a kernel."""
    q = module.quantizer
    name = q.bit_type.name
    w = weight.detach().contiguous().float()
    channels = w.shape[0]
    inner = w.numel() // channels
    scale = _per_channel(q.dic_scale[name], channels, w.device)
    zp = _per_channel(q.dic_zero_point[name], channels, w.device)
    out = torch.empty_like(w)
    _cabi.check(_cabi.lib().p2v_fake_quant_f32(w.data_ptr(), out.data_ptr(), None, 1, channels, inner,
                                              scale.data_ptr(), zp.data_ptr(), q.bit_type.lower_bound,
                                              q.bit_type.upper_bound, _cabi.current_stream()))
    return out


def qlinear(module, x, weight_smoothed):
    """QLinear.forward in quantized mode (models/ptq/layers.py:172-178): weights fake-quantized by the
    library kernel, then the fp32 product exactly as the reference forms it (a plain library GEMM; the
    int8 tensor-core product lives in the whole-model engine, where the input codes are known)."""
    return F.linear(x, _fake_quant_weight(module, weight_smoothed), module.bias)


def qconv2d(module, x):
    """QConv2d.forward in quantized mode (models/ptq/layers.py:93-97)."""
    w = _fake_quant_weight(module, module.weight)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        return F.conv2d(x, w, module.bias, module.stride, module.padding, module.dilation, module.groups)


def qint_layernorm(module, x, in_quantizer, out_quantizer, out_quantizer_scale, in_scale_expand):
    """QIntLayerNorm.forward, mode 'int' (models/ptq/layers.py:255-289)."""
    in_scale = in_quantizer.scale
    if in_scale_expand != 1:
        in_scale = in_scale.unsqueeze(-1).expand(-1, in_scale_expand).T.reshape(-1)
    out_scale = out_quantizer.scale
    assert in_scale is not None and out_scale is not None
    if out_quantizer_scale is not None:
        out_scale = out_scale * out_quantizer_scale.to(out_scale.device)
    d = x.shape[-1]
    x = x.contiguous().float()
    in_scale = _per_channel(in_scale, d, x.device)
    out_scale = _per_channel(out_scale, d, x.device)
    in_scale1 = in_scale.min()
    in_mask = (in_scale / in_scale1).round().contiguous()
    gamma = module.weight.detach().float().contiguous()
    beta = module.bias.detach().float().contiguous()
    out = torch.empty_like(x)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    _cabi.check(_cabi.lib().p2v_layernorm_int_f32(x.data_ptr(), out.data_ptr(), x.numel() // d, d, in_scale.data_ptr(),
                                                 in_mask.data_ptr(), float(in_scale1), gamma.data_ptr(),
                                                 beta.data_ptr(), out_scale.data_ptr(), flag.data_ptr(),
                                                 _cabi.current_stream()))
    if int(flag) != 0:
        raise ValueError('QIntLayerNorm: input is not on the grid of in_quantizer (|x / scale| >= 2^20)')
    return out


def qint_softmax(module, x, scale):
    """QIntSoftmax.forward with log_i_softmax (models/ptq/layers.py:323-376): fp32 scores on the grid
    `scale` in, dequantized probabilities 2^-k out."""
    scale = torch.as_tensor(scale, dtype=torch.float32).reshape(-1)
    if scale.numel() != 1:
        raise NotImplementedError('QIntSoftmax: expected a layer-wise (scalar) score scale')
    x0_int, b_int, c_int = module.exp_constants(scale.cpu())
    x = x.contiguous().float()
    n = x.shape[-1]
    out = torch.empty_like(x)
    _cabi.check(_cabi.lib().p2v_softmax_log_int_f32(x.data_ptr(), out.data_ptr(), None, x.numel() // n, n, float(scale),
                                                   float(x0_int), float(b_int), float(c_int), module.EXP_BITS,
                                                   2 ** module.bit_type.bits, _cabi.current_stream()))
    return out
