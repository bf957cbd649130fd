"""Quantization flag bundle consumed by the model factories (reference interface: config.py:4-51).

The reference spells every flag out as an attribute assignment; here the flag set is a table so that the
two switches (`ptf`, `lis`) read as what they are: overrides of three groups of fields.
"""
from .ptq.bit_type import BIT_TYPE_DICT

# field -> value shared by every configuration: int8 activations, int4 channel-wise minmax weights
# (the per-call `bit_config` list overrides the weight width layer by layer)
_COMMON = {
    'BIT_TYPE_W': BIT_TYPE_DICT['int4'],
    'BIT_TYPE_A': BIT_TYPE_DICT['int8'],
    'OBSERVER_W': 'minmax',
    'QUANTIZER_W': 'uniform',
    'QUANTIZER_A': 'uniform',
    'QUANTIZER_A_LN': 'uniform',
    'CALIBRATION_MODE_W': 'channel_wise',
    'CALIBRATION_MODE_A': 'layer_wise',
    'CALIBRATION_MODE_S': 'layer_wise',
}


class Config:
    """``Config(ptf, lis, quant_method)``

    quant_method  activation observer: 'minmax' (P2-ViT power-of-two scale search) or the FQ-ViT float-scale
                  observers 'ema' / 'percentile' / 'omse'.
    lis           log-int-softmax: 4-bit log2 codes from an integer exp (else a float softmax in 8 bits).
    ptf           power-of-two-factor, per-channel activation scales in front of the integer LayerNorm
                  (else the plain activation observer and a float LayerNorm).
    """

    def __init__(self, ptf=True, lis=True, quant_method='minmax'):
        fields = dict(_COMMON, OBSERVER_A=quant_method)
        fields.update(self._softmax_fields(lis, quant_method))
        fields.update(self._layernorm_fields(ptf, quant_method))
        for name, value in fields.items():
            setattr(self, name, value)

    @staticmethod
    def _softmax_fields(lis, quant_method):
        if lis:
            return dict(INT_SOFTMAX=True, BIT_TYPE_S=BIT_TYPE_DICT['uint4'], OBSERVER_S='minmax', QUANTIZER_S='log2')
        return dict(INT_SOFTMAX=False, BIT_TYPE_S=BIT_TYPE_DICT['uint8'], OBSERVER_S=quant_method,
                    QUANTIZER_S=_COMMON['QUANTIZER_A'])

    @staticmethod
    def _layernorm_fields(ptf, quant_method):
        if ptf:
            return dict(INT_NORM=True, OBSERVER_A_LN='ptf', CALIBRATION_MODE_A_LN='channel_wise')
        return dict(INT_NORM=False, OBSERVER_A_LN=quant_method, CALIBRATION_MODE_A_LN=_COMMON['CALIBRATION_MODE_A'])
