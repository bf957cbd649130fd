"""Calibration observers (reference: models/ptq/observer/*)."""
from .base import BaseObserver
from .float_scale import EmaObserver, OmseObserver, PercentileObserver
from .minmax import MinmaxObserver
from .ptf import PtfObserver
from . import utils

str2observer = {
    'minmax': MinmaxObserver,
    'ema': EmaObserver,
    'omse': OmseObserver,
    'percentile': PercentileObserver,
    'ptf': PtfObserver,
}


def build_observer(observer_str, module_type, bit_type, calibration_mode):
    """reference: models/ptq/observer/build.py:17-22"""
    return str2observer[observer_str](module_type, bit_type, calibration_mode)
