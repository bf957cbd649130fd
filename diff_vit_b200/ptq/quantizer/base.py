"""Quantizer base (reference: models/ptq/quantizer/base.py:6-45)."""
import torch.nn as nn

_WEIGHT_RANGE = {'conv_weight': (-1, 1, 1, 1), 'linear_weight': (-1, 1)}
_ACT_RANGE = {2: (1, -1), 3: (1, 1, -1), 4: (1, -1, 1, 1)}


class BaseQuantizer(nn.Module):

    def __init__(self, bit_type, observer, module_type):
        super().__init__()
        self.bit_type = bit_type
        self.observer = observer
        self.module_type = module_type

    def get_reshape_range(self, inputs):
        """Broadcast shape of a per-channel scale: out-channel first for weights, last dim for
        [.., C] activations, dim 1 for NCHW."""
        if self.module_type in _WEIGHT_RANGE:
            return _WEIGHT_RANGE[self.module_type]
        if self.module_type == 'activation' and inputs.dim() in _ACT_RANGE:
            return _ACT_RANGE[inputs.dim()]
        raise NotImplementedError

    def update_quantization_params(self, *args, **kwargs):
        pass

    def quant(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def dequantize(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def forward(self, inputs):
        return self.dequantize(self.quant(inputs))
