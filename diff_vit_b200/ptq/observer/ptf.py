"""Power-of-two-factor observer for LayerNorm inputs (reference: models/ptq/observer/ptf.py:8-134).

One float base scale for the tensor plus a per-channel factor 2^m, m in {0,1,2,3}; the reference's
Python loop over channels is replaced by one vectorised pass per candidate factor."""
import torch

from ... import dist as _dist
from .base import BaseObserver


class PtfObserver(BaseObserver):

    def __init__(self, module_type, bit_type, calibration_mode):
        super().__init__(module_type, bit_type, calibration_mode)

    def update(self, v):
        self.v = v
        self._track_minmax(v)

    def get_quantization_params(self, inputs, *args, **kwargs):
        max_val, min_val = self.max_val, self.min_val
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        max_val_t = torch.max(-min_val.min(), max_val.max())
        scale8 = 2 * max_val_t / float(qmax - qmin)
        scale8.clamp_(self.eps)
        scale4 = scale8 / 2
        scale2 = scale4 / 2
        scale1 = scale2 / 2
        zero_point = torch.zeros_like(max_val.max(), dtype=torch.int64)
        if inputs.dim() != 3:
            raise NotImplementedError('PTF calibration expects a [B, N, C] activation')
        from . import gpu_stats
        if gpu_stats.usable(inputs):
            # fused kernel: per-channel error of the four power-of-two factors in one pass
            sums, count = gpu_stats.scale_sse(inputs, [float(s) for s in (scale1, scale2, scale4, scale8)], qmin, qmax,
                                              per_channel=True)
            self.scale_mask = 2 ** gpu_stats.first_argmin_mean(sums, count)
            return scale1 * self.scale_mask, zero_point
        best = None
        choice = torch.zeros_like(max_val)
        for m, s in enumerate((scale1, scale2, scale4, scale8)):
            q = ((inputs / s + zero_point).round().clamp(qmin, qmax) - zero_point) * s
            score = _dist.global_mean((inputs - q).abs().pow(2.0).reshape(-1, inputs.shape[-1]), 0)
            if best is None:
                best = score
            else:
                better = score < best
                best = torch.where(better, score, best)
                choice = torch.where(better, torch.full_like(choice, float(m)), choice)
        self.scale_mask = 2 ** choice
        return scale1 * self.scale_mask, zero_point
