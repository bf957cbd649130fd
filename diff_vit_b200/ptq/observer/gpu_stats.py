"""Calibration statistics on the GPU: the observers' range and candidate-scale error reductions run in the
fused sm_100a kernels of csrc/p2v_observe.cu whenever the activation lives on a CUDA device.  On CPU tensors
(unit tests, gloo) the observers keep their PyTorch reductions, which implement the same arithmetic.
"""
import ctypes as C

import torch


def usable(x):
    """CUDA fp32 activation and a loadable C-ABI library."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.numel() > 0):
        return False
    try:
        from ... import _cabi
        _cabi.lib()
        return True
    except Exception:
        return False


def _stream():
    return torch.cuda.current_stream().cuda_stream


def minmax(x, per_channel):
    """(min, max) of the whole tensor (0-dim) or per innermost channel ([C])."""
    from ... import _cabi
    x = x.detach().contiguous()
    c = x.shape[-1]
    rows = x.numel() // c
    n = c if per_channel else 1
    lo = torch.full((n,), float('inf'), dtype=torch.float32, device=x.device)
    hi = torch.full((n,), float('-inf'), dtype=torch.float32, device=x.device)
    _cabi.check(_cabi.lib().p2v_observe_minmax(x.data_ptr(), rows, c, 1 if per_channel else 0, lo.data_ptr(),
                                              hi.data_ptr(), _stream()))
    return (lo, hi) if per_channel else (lo[0], hi[0])


def scale_sse(x, scales, qmin, qmax, per_channel):
    """fp64 sums of squared fake-quant error for each candidate scale: [K] or [K, C]; plus the element count
    each sum runs over."""
    from ... import _cabi
    x = x.detach().contiguous()
    c = x.shape[-1]
    rows = x.numel() // c
    k = len(scales)
    out = torch.zeros((k, c) if per_channel else (k,), dtype=torch.float64, device=x.device)
    arr = (C.c_float * k)(*[float(s) for s in scales])
    _cabi.check(_cabi.lib().p2v_observe_scale_sse(x.data_ptr(), rows, c, 1 if per_channel else 0, arr, k, float(qmin),
                                                 float(qmax), out.data_ptr(), _stream()))
    return out, (rows if per_channel else rows * c)


def first_argmin_mean(sums, count):
    """Index of the first minimum of the (globally averaged) scores along dim 0, as a float tensor."""
    from ... import dist as _dist
    if _dist.is_active():
        import torch.distributed as tdist
        cnt = torch.tensor([float(count)], dtype=torch.float64, device=sums.device)
        tdist.all_reduce(sums, op=tdist.ReduceOp.SUM, group=_dist._GROUP)
        tdist.all_reduce(cnt, op=tdist.ReduceOp.SUM, group=_dist._GROUP)
        count = float(cnt.item())
    score = (sums / count).float()
    best = score[0]
    choice = torch.zeros_like(best)
    for k in range(1, score.shape[0]):
        better = score[k] < best
        best = torch.where(better, score[k], best)
        choice = torch.where(better, torch.full_like(choice, float(k)), choice)
    return choice
