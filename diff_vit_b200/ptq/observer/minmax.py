"""P2-ViT power-of-two scale search (reference: models/ptq/observer/minmax.py:9-272).

The reference walks a Python loop over every output channel and issues five GEMV-sized
``F.linear`` calls per channel.  Here the four candidate exponents of ALL channels are scored with
four full-size GEMMs and a per-channel reduction, which is the same arithmetic per channel
(column j of ``F.linear(x, W, b)`` is ``F.linear(x, W[j:j+1], b[j:j+1])``) at ~1/out_channels of
the cost.  Layer-wise searches are op-for-op identical to the reference.
"""
import torch
from torch.nn import functional as F

from .base import BaseObserver
from .utils import ln2_floor


class MinmaxObserver(BaseObserver):

    def __init__(self, module_type, bit_type, calibration_mode):
        super().__init__(module_type, bit_type, calibration_mode)
        # Fixed at construction: later bit_type swaps (QLinear's calibration loop) do not
        # change it, so uint3/uint4 weights are "calibrated" with the symmetric formula too.
        self.symmetric = self.bit_type.signed

    def update(self, v):
        self.v = v
        self._track_minmax(v)

    # -- layer output used to score a candidate (minmax.py:119-176) ---------------------------
    def _project(self, w):
        if self.module_type == 'activation':
            return w
        bias = self.others[0] if self.others else None
        if self.module_type == 'linear_weight':
            return F.linear(self.input, w, bias)
        if self.module_type == 'conv_weight':
            _, stride, padding, dilation, groups = self.others
            return F.conv2d(self.input, w, bias, stride, padding, dilation, groups)
        raise NotImplementedError(self.module_type)

    def _channel_mse(self, a, b):
        from ... import dist as _dist
        d = (a - b).abs().pow(2.0)
        if self.calibration_mode == 'layer_wise':
            return _dist.global_mean(d).reshape(1)
        if self.module_type == 'conv_weight':
            return _dist.global_mean(d, (0, 2, 3))
        return _dist.global_mean(d.reshape(-1, d.shape[-1]), 0)

    def _search_exponent(self, scale, zero_point=None):
        """Pick alpha in {af-1, af, af+1, af+2}, af = floor(log2 scale), minimising the MSE of the
        layer output (weights) or of the tensor itself (activations); first minimum wins
        (minmax.py:180-242)."""
        qmin, qmax = self.bit_type.lower_bound, self.bit_type.upper_bound
        alpha_floor = ln2_floor(scale)
        layer_wise = self.calibration_mode == 'layer_wise'
        if self.module_type == 'activation':
            if not layer_wise:
                raise NotImplementedError('channel-wise PoT search is only defined for weights')
            target = self.input
        else:
            target = self.v
        if self.attn and layer_wise:
            raise NotImplementedError('attn-output calibration objective is never enabled by the callers')
        from . import gpu_stats
        if self.module_type == 'activation' and zero_point is None and gpu_stats.usable(target):
            # fused kernel: the four candidate errors in one pass over the activation
            steps = [float(2.0 ** (float(alpha_floor.reshape(-1)[0]) - 1 + k)) for k in range(4)]
            sums, count = gpu_stats.scale_sse(target, steps, qmin, qmax, per_channel=False)
            return alpha_floor - 1 + gpu_stats.first_argmin_mean(sums, count).reshape(1)
        zp = torch.zeros(1, device=target.device) if zero_point is None else zero_point
        view = (-1,) + (1,) * (target.dim() - 1)
        ref_out = self._project(target)
        best = None
        choice = torch.zeros_like(alpha_floor)
        for k in range(4):
            step = 2 ** (alpha_floor - 1 + k)
            step = step[0] if layer_wise else step.reshape(view)
            cand = ((target / step + zp).round().clamp(qmin, qmax) - zp) * step
            score = self._channel_mse(ref_out, self._project(cand))
            if best is None:
                best = score
            else:
                better = score < best
                best = torch.where(better, score, best)
                choice = torch.where(better, torch.full_like(choice, float(k)), choice)
        return alpha_floor - 1 + choice

    def get_quantization_params(self, x, others=None, attn=False, attn_para=None, *args, **kwargs):
        max_val, min_val = self.max_val, self.min_val
        self.input, self.others, self.attn, self.attn_para = x, others, attn, attn_para
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        if self.symmetric:
            zero_point = torch.zeros_like(max_val, dtype=torch.int64)
            max_val = torch.max(-min_val, max_val)
            scale = max_val / (float(qmax - qmin) / 2)
            scale = 2 ** self._search_exponent(scale)
        else:
            scale = (max_val - min_val) / float(qmax - qmin)
            zero_point = qmin - torch.round(min_val / scale)
            zero_point.clamp_(qmin, qmax)
            zp = None if not bool((zero_point != 0).any()) else zero_point
            scale = 2 ** self._search_exponent(scale, zp)
        scale.clamp_(self.eps)
        self.input = None
        return scale, zero_point
