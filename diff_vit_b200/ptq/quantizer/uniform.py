"""Uniform affine quantizer (reference: models/ptq/quantizer/uniform.py:8-127).

Activations keep one (scale, zero_point); weights keep one pair per calibrated bit type in
``dic_scale`` / ``dic_zero_point`` keyed by the bit-type name.  The reference pins every operand
to CUDA with ``.cuda()``; this mirror follows the input's device instead."""
from .base import BaseQuantizer


class UniformQuantizer(BaseQuantizer):

    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.scale = None
        self.zero_point = None
        self.dic_scale = {}
        self.dic_zero_point = {}

    def update_quantization_params(self, *args, **kwargs):
        scale, zero_point = self.observer.get_quantization_params(*args, **kwargs)
        if self.module_type == 'activation':
            self.scale, self.zero_point = scale, zero_point
        else:
            self.dic_scale[self.bit_type.name] = scale
            self.dic_zero_point[self.bit_type.name] = zero_point

    def _params(self, inputs, scale, zero_point):
        if scale is None:
            scale = self.scale if self.module_type == 'activation' else self.dic_scale[self.bit_type.name]
        if zero_point is None:
            zero_point = (self.zero_point if self.module_type == 'activation'
                          else self.dic_zero_point[self.bit_type.name])
        shape = self.get_reshape_range(inputs)
        return (scale.reshape(shape).to(inputs.device), zero_point.reshape(shape).to(inputs.device))

    def quant(self, inputs, scale=None, zero_point=None):
        scale, zero_point = self._params(inputs, scale, zero_point)
        outputs = inputs / scale + zero_point
        return outputs.round().clamp(self.bit_type.lower_bound, self.bit_type.upper_bound)

    def dequantize(self, inputs, scale=None, zero_point=None):
        scale, zero_point = self._params(inputs, scale, zero_point)
        return (inputs - zero_point) * scale
