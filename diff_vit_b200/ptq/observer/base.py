"""Observer base class (reference: models/ptq/observer/base.py:5-37)."""
import torch


class BaseObserver:

    def __init__(self, module_type, bit_type, calibration_mode):
        self.module_type = module_type
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.max_val = None
        self.min_val = None
        self.eps = torch.finfo(torch.float32).eps

    def reshape_tensor(self, v):
        """Weights -> [out, -1]; activations -> [channels, -1] (NCHW is permuted channel-last first)."""
        if not isinstance(v, torch.Tensor):
            v = torch.tensor(v)
        v = v.detach()
        if self.module_type in ('conv_weight', 'linear_weight'):
            return v.reshape(v.shape[0], -1)
        if self.module_type == 'activation':
            if v.dim() == 4:
                v = v.permute(0, 2, 3, 1)
            return v.reshape(-1, v.shape[-1]).transpose(0, 1)
        raise NotImplementedError(self.module_type)

    def _track_minmax(self, v):
        """Running per-channel min/max, collapsed to scalars for layer-wise calibration."""
        from . import gpu_stats
        if self.module_type == 'activation' and gpu_stats.usable(v) and \
                (self.calibration_mode == 'layer_wise' or v.dim() in (2, 3)):
            # fused range kernel (csrc/p2v_observe.cu); the scalar case does not need the channel split at all
            cur_min, cur_max = gpu_stats.minmax(v, per_channel=self.calibration_mode != 'layer_wise')
        else:
            v = self.reshape_tensor(v)
            cur_max = v.max(axis=1).values
            cur_min = v.min(axis=1).values
        if self.module_type == 'activation':   # weights are replicated; activations are sharded over ranks
            from ... import dist as _dist
            _dist.reduce_max_(cur_max)
            _dist.reduce_min_(cur_min)
        self.max_val = cur_max if self.max_val is None else torch.max(cur_max, self.max_val)
        self.min_val = cur_min if self.min_val is None else torch.min(cur_min, self.min_val)
        if self.calibration_mode == 'layer_wise':
            self.max_val = self.max_val.max()
            self.min_val = self.min_val.min()

    def update(self, v):
        raise NotImplementedError

    def get_quantization_params(self, *args, **kwargs):
        raise NotImplementedError
