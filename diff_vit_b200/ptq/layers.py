"""The five quantized operators of P2-ViT (reference: models/ptq/layers.py:12-395).

Same constructors, attributes (`quant`, `calibrate`, `last_calibrate`, `bit_type`, `observer`,
`quantizer`, `module_type`) and `forward` signatures as the reference, so the model graphs, the
calibration flow and third-party forward hooks work unchanged.

Execution model (this is where the mirror departs from the reference):
  * calibration / float mode (``quant=False``) is plain PyTorch on whatever device the input lives on;
  * quantized mode (``quant=True``) never runs in PyTorch.  The model-level forward hands the whole
    graph to the integer engine (``diff_vit_b200.engine``); a module called on its own in quantized
    mode runs its sm_100a kernel through the C-ABI library and raises if the library or a CUDA
    device is missing.  There is no CPU fallback for quantized arithmetic.
"""
import torch
import torch.nn as nn
from torch.nn import functional as F

from .bit_type import BIT_TYPE_DICT, BIT_TYPE_LIST
from .observer import build_observer, utils
from .quantizer import build_quantizer


def _standalone_kernels(x, what):
    """Gate for per-module quantized execution: CUDA tensor + loaded C-ABI library, or an error."""
    if not x.is_cuda:
        raise RuntimeError(
            '%s: quantized execution runs only in the sm_100a kernels of libp2vit_b200.so; '
            'got a %s tensor and there is no CPU fallback' % (what, x.device))
    from .. import standalone
    return standalone


class _QWeightMixin:
    """Calibration loop shared by QConv2d and QLinear (layers.py:57-73,148-170): every bit type
    except uint8 is calibrated in BIT_TYPE_LIST order; int8 is layer-wise, the rest channel-wise."""

    def _init_quant(self, quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str,
                    quantizer_str, module_type):
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = module_type
        self.observer = build_observer(observer_str, module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, module_type)

    def _calibration_bit_types(self):
        for bit_type in BIT_TYPE_LIST:
            if bit_type is BIT_TYPE_DICT['uint8']:
                continue
            self.quantizer.bit_type = bit_type
            self.observer.bit_type = bit_type
            self.observer.calibration_mode = 'layer_wise' if bit_type is BIT_TYPE_DICT['int8'] else 'channel_wise'
            yield bit_type

    @staticmethod
    def _bit_type_of(bits):
        return BIT_TYPE_DICT['int' + str(bits)]

    def _select_bits(self, bit_config):
        if bit_config:
            bit_type = BIT_TYPE_DICT['int' + str(bit_config)]  # KeyError for unsupported widths, as the reference
            self.quantizer.bit_type = bit_type
            self.observer.bit_type = bit_type


class QConv2d(nn.Conv2d, _QWeightMixin):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, quant=False, calibrate=False, last_calibrate=False,
                 bit_type=BIT_TYPE_DICT['int8'], calibration_mode='layer_wise', observer_str='minmax',
                 quantizer_str='uniform'):
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                         stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias)
        self._init_quant(quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str,
                         quantizer_str, 'conv_weight')

    def forward(self, x, bit_config):
        if self.calibrate:
            for _ in self._calibration_bit_types():
                self.quantizer.observer.update(self.weight)
                if self.last_calibrate:
                    self.quantizer.update_quantization_params(
                        x, others=[self.bias, self.stride, self.padding, self.dilation, self.groups])
        if not self.quant or bit_config == -1:
            return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        self._select_bits(bit_config)
        return _standalone_kernels(x, 'QConv2d').qconv2d(self, x)


class QLinear(nn.Linear, _QWeightMixin):

    def __init__(self, in_features, out_features, bias=True, quant=False, calibrate=False,
                 last_calibrate=False, bit_type=BIT_TYPE_DICT['int8'], calibration_mode='layer_wise',
                 observer_str='minmax', quantizer_str='uniform'):
        super().__init__(in_features, out_features, bias)
        self._init_quant(quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str,
                         quantizer_str, 'linear_weight')

    def forward(self, x, global_distance=[], bit_config=None, weight_smoothed=None, attn=False, attn_para=None):
        # The mutable default for global_distance is part of the reference signature (layers.py:133).
        if weight_smoothed is None:
            weight_smoothed = self.weight
        float_path = not self.quant or bit_config == -1
        if float_path:
            y = F.linear(x, weight_smoothed, self.bias)
        if self.calibrate:
            distance = []
            for _ in self._calibration_bit_types():
                self.quantizer.observer.update(weight_smoothed)
                self.quantizer.update_quantization_params(x, others=[self.bias], attn=attn, attn_para=attn_para)
                distance.append(utils.lp_loss(weight_smoothed, self.quantizer(weight_smoothed), p=2.0,
                                              reduction='all'))
            global_distance.append(distance)
        if float_path:
            return y
        self._select_bits(bit_config)
        return _standalone_kernels(x, 'QLinear').qlinear(self, x, weight_smoothed)


class QAct(nn.Module):

    def __init__(self, quant=False, calibrate=False, last_calibrate=False, bit_type=BIT_TYPE_DICT['int8'],
                 calibration_mode='layer_wise', observer_str='minmax', quantizer_str='uniform'):
        super().__init__()
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = 'activation'
        self.observer = build_observer(observer_str, self.module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, self.module_type)

    def forward(self, x, asymmetric=False, attn=False, attn_para=None):
        if self.calibrate:
            if asymmetric:
                self.quantizer.bit_type = BIT_TYPE_DICT['uint8']
                self.observer.bit_type = BIT_TYPE_DICT['uint8']
                self.observer.symmetric = False
            self.quantizer.observer.update(x)
            if self.last_calibrate:
                self.quantizer.update_quantization_params(x, attn=attn, attn_para=attn_para)
        if not self.quant:
            return x
        return _standalone_kernels(x, 'QAct').qact(self, x)


class QIntLayerNorm(nn.LayerNorm):
    """Float LayerNorm while calibrating (mode 'ln'); dyadic integer LayerNorm once quantized
    (mode 'int', reference: layers.py:255-289)."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__(normalized_shape, eps, elementwise_affine)
        assert isinstance(normalized_shape, int)
        self.mode = 'ln'

    def get_MN(self, x):
        """8-bit dyadic multiplier: x ~ M / 2^N with N in [0, 31], M in [0, 255]."""
        bit = 7
        N = torch.clamp(bit - torch.floor(torch.log2(x)), 0, 31)
        M = torch.clamp(torch.floor(x * torch.pow(2, N)), 0, 2 ** (bit + 1) - 1)
        return M, N

    def forward(self, x, in_quantizer=None, out_quantizer=None, out_quantizer_scale=None, in_scale_expand=1):
        if self.mode == 'ln':
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if self.mode == 'int':
            return _standalone_kernels(x, 'QIntLayerNorm').qint_layernorm(
                self, x, in_quantizer, out_quantizer, out_quantizer_scale, in_scale_expand)
        raise NotImplementedError(self.mode)


class QIntSoftmax(nn.Module):
    """log-int-softmax: I-BERT integer exp, integer row sum, 4-bit log2 code of sum/exp
    (reference: layers.py:295-395)."""

    EXP_BITS = 32  # 'n' of the integer exp (layers.py:350)

    def __init__(self, log_i_softmax=False, quant=False, calibrate=False, last_calibrate=False,
                 bit_type=BIT_TYPE_DICT['int8'], calibration_mode='layer_wise', observer_str='minmax',
                 quantizer_str='uniform'):
        super().__init__()
        self.log_i_softmax = log_i_softmax
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = 'activation'
        self.observer = build_observer(observer_str, self.module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, self.module_type)

    @staticmethod
    def log_round(x):
        """floor(log2 x), +1 when the mantissa is >= 1.5."""
        big = x.log2().floor()
        extra = (x - 2 ** big) >= 2 ** (big - 1)
        big[extra] = big[extra] + 1
        return big

    @staticmethod
    def exp_constants(scaling_factor):
        """(x0_int, b_int, c_int) of the integer exp for a given input scale, evaluated with the
        reference's own fp32 tensor expressions (layers.py:334-352) so host and device agree."""
        a, b, c = 0.35815147, 0.96963238, 1.
        b /= a
        c /= a
        x0_int = torch.floor(-0.6931 / scaling_factor)
        b_int = torch.floor(b / scaling_factor)
        c_int = torch.floor(c / scaling_factor ** 2)
        return x0_int, b_int, c_int

    @staticmethod
    def int_softmax(x, scaling_factor):
        n = QIntSoftmax.EXP_BITS
        x0_int, b_int, c_int = QIntSoftmax.exp_constants(scaling_factor)
        x_int = x / scaling_factor
        x_int = x_int - x_int.max(dim=-1, keepdim=True).values
        x_int = torch.max(x_int, n * x0_int)
        q = torch.floor(x_int / x0_int)
        r = x_int - x0_int * q
        z = r + b_int
        z = r * z
        z = z + c_int
        exp_int = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
        return exp_int, exp_int.sum(dim=-1, keepdim=True)

    def forward(self, x, scale):
        if self.log_i_softmax and scale is not None:
            if self.quant:
                return _standalone_kernels(x, 'QIntSoftmax').qint_softmax(self, x, scale)
            # calibration pass: the same integer pipeline evaluated on not-yet-quantized scores
            exp_int, exp_int_sum = self.int_softmax(x, scale)
            rounds = self.log_round(torch.round(exp_int_sum / exp_int))
            levels = 2 ** self.bit_type.bits
            mask = rounds >= levels
            deq = 2 ** (-torch.clamp(rounds, 0, levels - 1))
            deq[mask] = 0
            return deq
        return x.softmax(dim=-1)
