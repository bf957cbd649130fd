"""FQ-ViT float-scale activation observers: EMA, percentile and OMSE
(reference: models/ptq/observer/ema.py:7-58, percentile.py:9-71, omse.py:8-56)."""
import numpy as np
import torch

from ... import dist as _dist
from . import gpu_stats
from .base import BaseObserver


def _affine_params(obs, max_val, min_val):
    qmax, qmin = obs.bit_type.upper_bound, obs.bit_type.lower_bound
    if obs.symmetric:
        max_val = torch.max(-min_val, max_val)
        scale = max_val / (float(qmax - qmin) / 2)
        scale.clamp_(obs.eps)
        zero_point = torch.zeros_like(max_val, dtype=torch.int64)
    else:
        scale = (max_val - min_val) / float(qmax - qmin)
        scale.clamp_(obs.eps)
        zero_point = qmin - torch.round(min_val / scale)
        zero_point.clamp_(qmin, qmax)
    return scale, zero_point


class EmaObserver(BaseObserver):

    def __init__(self, module_type, bit_type, calibration_mode, ema_sigma=0.01):
        super().__init__(module_type, bit_type, calibration_mode)
        self.ema_sigma = ema_sigma
        self.symmetric = self.bit_type.signed

    def update(self, v):
        v = self.reshape_tensor(v)
        cur_max = _dist.reduce_max_(v.max(axis=1).values)
        cur_min = _dist.reduce_min_(v.min(axis=1).values)
        if self.max_val is None:
            self.max_val, self.min_val = cur_max, cur_min
        else:
            self.max_val = self.max_val + self.ema_sigma * (cur_max - self.max_val)
            self.min_val = self.min_val + self.ema_sigma * (cur_min - self.min_val)
        if self.calibration_mode == 'layer_wise':
            self.max_val = self.max_val.max()
            self.min_val = self.min_val.min()

    def get_quantization_params(self, *args, **kwargs):
        return _affine_params(self, self.max_val, self.min_val)


class PercentileObserver(BaseObserver):
    """The constructor arguments are accepted and ignored, as in the reference, which hard-codes
    sigma = 0.01 and alpha = 0.99999 (percentile.py:19-20)."""

    def __init__(self, module_type, bit_type, calibration_mode, percentile_sigma=0.01,
                 percentile_alpha=0.99999):
        super().__init__(module_type, bit_type, calibration_mode)
        self.percentile_sigma = 0.01
        self.percentile_alpha = 0.99999
        self.symmetric = self.bit_type.signed

    def update(self, v):
        assert self.calibration_mode == 'layer_wise'
        flat = self.reshape_tensor(v).reshape(-1)
        if _dist.is_active() or gpu_stats.usable(flat):
            # exact radix select in the sm_100a histogram kernel (summed over ranks when calibrating
            # data-parallel) instead of a full sort; same interpolation as the reference's call
            total = flat.numel()
            if _dist.is_active():
                total = int(_dist.reduce_sum_(torch.tensor([total], dtype=torch.int64, device=flat.device)))
            cur_max, cur_min = gpu_stats.quantile_pair(flat, self.percentile_alpha, total)
        else:
            try:
                cur_max = torch.quantile(flat, self.percentile_alpha)
                cur_min = torch.quantile(flat, 1.0 - self.percentile_alpha)
            except RuntimeError:  # torch.quantile refuses > 16M elements
                host = flat.cpu()
                cur_max = torch.tensor(np.percentile(host, self.percentile_alpha * 100),
                                       device=v.device, dtype=torch.float32)
                cur_min = torch.tensor(np.percentile(host, (1 - self.percentile_alpha) * 100),
                                       device=v.device, dtype=torch.float32)
        if self.max_val is None:
            self.max_val, self.min_val = cur_max, cur_min
        else:
            self.max_val = self.max_val + self.percentile_sigma * (cur_max - self.max_val)
            self.min_val = self.min_val + self.percentile_sigma * (cur_min - self.min_val)

    def get_quantization_params(self, *args, **kwargs):
        return _affine_params(self, self.max_val, self.min_val)


class OmseObserver(BaseObserver):
    """90-step range shrink minimising the fake-quant MSE with an asymmetric zero point.

    The reference signature ``get_quantization_params(self, inputs)`` rejects the ``attn=`` /
    ``attn_para=`` keywords QAct always passes (layers.py:216) and so raises TypeError as shipped;
    this mirror accepts and ignores them."""

    def __init__(self, module_type, bit_type, calibration_mode):
        super().__init__(module_type, bit_type, calibration_mode)

    def update(self, v):
        self._track_minmax(v)

    def get_quantization_params(self, inputs, *args, **kwargs):
        max_val, min_val = self.max_val, self.min_val
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        best_score = 1e+10
        scale = zero_point = None
        for i in range(90):
            new_max = max_val * (1.0 - (i * 0.01))
            new_min = min_val * (1.0 - (i * 0.01))
            new_scale = (new_max - new_min) / float(qmax - qmin)
            new_scale.clamp_(self.eps)
            new_zero_point = qmin - torch.round(new_min / new_scale)
            new_zero_point.clamp_(qmin, qmax)
            inputs_q = ((inputs / new_scale + new_zero_point).round().clamp(qmin, qmax) -
                        new_zero_point) * new_scale
            score = _dist.global_mean((inputs - inputs_q).abs().pow(2.0))
            if score < best_score:
                best_score = score
                self.max_val, self.min_val = new_max, new_min
                scale, zero_point = new_scale, new_zero_point
        return scale, zero_point
