"""Calibration helpers shared by the observers."""
import torch


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """L_p distance (reference: models/ptq/observer/utils.py:2-9): per-row sum then mean for
    'none', plain mean over all elements otherwise."""
    d = (pred - tgt).abs().pow(p)
    return d.sum(1).mean() if reduction == 'none' else d.mean()


def ln2_floor(x):
    """floor(ln(x)/ln(2)) evaluated in fp32 exactly as the reference does (minmax.py:65-73):
    the quotient of two rounded logs, NOT an exact exponent extraction."""
    two = torch.tensor([2.0], dtype=torch.float32, device=x.device)
    return torch.floor(torch.div(torch.log(x), torch.log(two)))


def ln2_round(x):
    """Nearest power of two in the LINEAR domain, ties down (minmax.py:69-73)."""
    y = ln2_floor(x)
    return torch.gt(x - 2 ** y, 2 ** (y + 1) - x) + y


def fake_quant(x, scale, zero_point, qmin, qmax):
    """round-half-even(x / s + zp) -> clamp -> (q - zp) * s, clamp after round
    (reference: models/ptq/quantizer/uniform.py:82-88,123-127)."""
    return ((x / scale + zero_point).round().clamp(qmin, qmax) - zero_point) * scale
