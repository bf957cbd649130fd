"""Quantizers: the objects that hold calibrated (scale, zero_point) state and turn tensors into codes
and back (reference interfaces: models/ptq/quantizer/{base,uniform,log2,build}.py).

`UniformQuantizer` keeps one pair for activations and one pair per calibrated bit width for weights
(`dic_scale[bit_name]`); `Log2Quantizer` is the softmax code the reference constructs for QIntSoftmax but
never calls on the active path.  The reference pins operands to CUDA with `.cuda()`; these follow the input.
"""
import torch
import torch.nn as nn

# broadcast shape of a per-channel parameter, keyed by (module kind, tensor rank)
_PARAM_SHAPE = {
    ('conv_weight', None): (-1, 1, 1, 1),
    ('linear_weight', None): (-1, 1),
    ('activation', 2): (1, -1),
    ('activation', 3): (1, 1, -1),
    ('activation', 4): (1, -1, 1, 1),
}


def map_state(fn, value):
    """Apply a Module._apply function (device / dtype move) to calibration state held as plain tensors,
    or as lists / dicts of them.  Integer tensors (zero points) only change device."""
    if isinstance(value, torch.Tensor):
        moved = fn(value)
        return moved if value.is_floating_point() else value.to(moved.device)
    if isinstance(value, list):
        return [map_state(fn, v) for v in value]
    if isinstance(value, dict):
        return {k: map_state(fn, v) for k, v in value.items()}
    return value


class BaseQuantizer(nn.Module):

    def __init__(self, bit_type, observer, module_type):
        super().__init__()
        self.bit_type, self.observer, self.module_type = bit_type, observer, module_type

    def get_reshape_range(self, inputs):
        """Out-channel first for weights, innermost dim for [.., C] activations, dim 1 for NCHW."""
        key = (self.module_type, inputs.dim() if self.module_type == 'activation' else None)
        if key not in _PARAM_SHAPE:
            raise NotImplementedError(key)
        return _PARAM_SHAPE[key]

    def update_quantization_params(self, *args, **kwargs):
        """Nothing to calibrate by default."""

    def quant(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def dequantize(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def forward(self, inputs):
        return self.dequantize(self.quant(inputs))


class UniformQuantizer(BaseQuantizer):
    """codes = clamp(RNE(x / s + zp)); value = (codes - zp) * s."""

    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.scale = self.zero_point = None          # activations
        self.dic_scale, self.dic_zero_point = {}, {}  # weights, per bit-type name

    @property
    def _is_act(self):
        return self.module_type == 'activation'

    def _apply(self, fn, *args, **kwargs):
        """The calibrated pairs are plain attributes (as in the reference); make them follow .cuda()/.to()."""
        super()._apply(fn, *args, **kwargs)
        for name in ('scale', 'zero_point', 'dic_scale', 'dic_zero_point'):
            setattr(self, name, map_state(fn, getattr(self, name)))
        return self

    def update_quantization_params(self, *args, **kwargs):
        scale, zero_point = self.observer.get_quantization_params(*args, **kwargs)
        if self._is_act:
            self.scale, self.zero_point = scale, zero_point
        else:
            name = self.bit_type.name
            self.dic_scale[name], self.dic_zero_point[name] = scale, zero_point

    def _broadcast(self, inputs, scale, zero_point):
        if scale is None:
            scale = self.scale if self._is_act else self.dic_scale[self.bit_type.name]
        if zero_point is None:
            zero_point = self.zero_point if self._is_act else self.dic_zero_point[self.bit_type.name]
        shape = self.get_reshape_range(inputs)
        return scale.reshape(shape).to(inputs.device), zero_point.reshape(shape).to(inputs.device)

    def quant(self, inputs, scale=None, zero_point=None):
        s, z = self._broadcast(inputs, scale, zero_point)
        return (inputs / s + z).round().clamp(self.bit_type.lower_bound, self.bit_type.upper_bound)

    def dequantize(self, inputs, scale=None, zero_point=None):
        s, z = self._broadcast(inputs, scale, zero_point)
        return (inputs - z) * s


class Log2Quantizer(BaseQuantizer):
    """k = clamp(RNE(-log2 p)) with everything at or beyond 2^bits flushed to probability 0."""

    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.softmax_mask = None

    def quant(self, inputs):
        levels = 1 << self.bit_type.bits
        k = (-inputs.log2()).round()
        self.softmax_mask = k >= levels
        return k.clamp(0, levels - 1)

    def dequantize(self, inputs):
        p = torch.pow(2.0, -inputs)
        p[self.softmax_mask] = 0
        return p


str2quantizer = {'uniform': UniformQuantizer, 'log2': Log2Quantizer}


def build_quantizer(quantizer_str, bit_type, observer, module_type):
    return str2quantizer[quantizer_str](bit_type, observer, module_type)
