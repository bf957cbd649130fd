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


# ---- exact order statistics (percentile observer) -------------------------------------------------------
_SELECT_PASSES = (21, 10, 0)   # 11-bit digits of the 32-bit order-preserving key, most significant first
_SELECT_BINS = 2048


def _digit_histogram(flat, prefix, mask, shift):
    """int64[2048] histogram of key bits [shift, shift+11) over the elements matching prefix under mask."""
    if flat.is_cuda:
        from ... import _cabi
        hist = torch.zeros(_SELECT_BINS, dtype=torch.int64, device=flat.device)
        _cabi.check(_cabi.lib().p2v_select_histogram(flat.data_ptr(), flat.numel(), prefix, mask, shift,
                                                    hist.data_ptr(), _stream()))
        return hist
    # CPU tensors (gloo runs, unit tests): the same digit histogram with torch ops
    bits = flat.view(torch.int32).to(torch.int64) & 0xffffffff
    key = torch.where(bits >= 0x80000000, bits ^ 0xffffffff, bits | 0x80000000)
    sel = key[(key & mask) == prefix]
    return torch.bincount((sel >> shift) & (_SELECT_BINS - 1), minlength=_SELECT_BINS)


def order_statistics(x, ranks):
    """Exact ascending order statistics x_(r), r 0-based, of a fp32 tensor; under data-parallel calibration
    the ranks refer to the union of every process's shard (digit histograms are summed over the group).
    Three streaming passes per requested rank; replaces the full sort behind torch.quantile / np.percentile
    (models/ptq/observer/percentile.py:27-38)."""
    from ... import dist as _dist
    flat = x.detach().contiguous().reshape(-1).float()
    out = []
    for r in ranks:
        prefix = mask = 0
        remaining = int(r)
        for shift in _SELECT_PASSES:
            hist = _digit_histogram(flat, prefix, mask, shift)
            if _dist.is_active():
                import torch.distributed as tdist
                tdist.all_reduce(hist, op=tdist.ReduceOp.SUM, group=_dist._GROUP)
            cum = torch.cumsum(hist.cpu(), 0)
            b = int(torch.searchsorted(cum, torch.tensor(remaining), right=True))
            if b >= _SELECT_BINS:
                raise IndexError('order statistic %d is beyond the %d elements observed' % (r, int(cum[-1])))
            if b > 0:
                remaining -= int(cum[b - 1])
            prefix |= b << shift
            mask |= ((_SELECT_BINS - 1) << shift) & 0xffffffff
        bits = prefix ^ 0x80000000 if prefix & 0x80000000 else prefix ^ 0xffffffff
        out.append(torch.tensor([bits], dtype=torch.int64).to(torch.int32).view(torch.float32)[0])
    return torch.stack(out).to(x.device)


def quantile_pair(x, alpha, total):
    """(quantile(alpha), quantile(1 - alpha)) over `total` elements (all ranks) with the interpolation rules of
    the call the reference would have made: torch.quantile up to 16M elements, np.percentile beyond
    (percentile.py:27-38).  Only the two neighbouring order statistics of each are ever materialised."""
    import numpy as np
    res = []
    for q in (alpha, 1.0 - alpha):
        if total <= 16_000_000:
            # ATen quantile: ranks = q * (n - 1) in the tensor dtype, lerp(below, above, frac)
            pos = torch.tensor(q, dtype=torch.float32) * (total - 1)
            lo, hi = int(pos.floor()), int(pos.ceil())
            v = order_statistics(x, [lo, hi])
            res.append(torch.lerp(v[0], v[1], (pos - pos.floor()).to(v.device)))
        else:
            # numpy 'linear' method on a float32 array: the quantile and the virtual index follow numpy's promotion
            qq = np.asanyarray(np.true_divide(q * 100, np.float32(100)))
            vi = np.asanyarray((total - 1) * qq)
            lo = int(np.floor(vi))
            hi = min(lo + 1, total - 1)
            v = order_statistics(x, [lo, hi]).cpu().numpy()
            gamma = np.asanyarray(vi - lo, dtype=vi.dtype)
            diff = np.subtract(v[1], v[0])
            val = np.add(v[0], diff * gamma)
            if gamma >= 0.5:
                val = np.subtract(v[1], diff * (1 - gamma))
            res.append(torch.tensor(val, device=x.device, dtype=torch.float32))
    return res[0], res[1]
