"""Per-module quantized execution (a Q-module called on its own after ``model_quant()``).

The fast path of this package is the whole-graph integer engine; this module keeps the individual
operators usable the way the reference's analysis scripts use them (forward hooks, hand-built graphs):
fp32 tensors in, fp32 dequantized tensors out, arithmetic in the sm_100a kernels.
"""
import torch

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


def _unsupported(what):
    raise NotImplementedError(
        '%s cannot be run on its own in quantized mode yet: its integer kernel needs the scales of the '
        'neighbouring quantizers, which only the model-level forward (IntegerEngine) knows' % what)


def qlinear(module, x, weight_smoothed):
    _unsupported('QLinear')


def qconv2d(module, x):
    _unsupported('QConv2d')


def qint_layernorm(module, x, in_quantizer, out_quantizer, out_quantizer_scale, in_scale_expand):
    _unsupported('QIntLayerNorm')


def qint_softmax(module, x, scale):
    _unsupported('QIntSoftmax')
