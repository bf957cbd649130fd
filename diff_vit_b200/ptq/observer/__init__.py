"""Calibration observers: running range statistics plus the search that turns them into (scale, zero_point)
(reference interfaces: models/ptq/observer/*).  `build_observer(name, ...)` is the factory the Q-modules use."""
from . import utils
from .base import BaseObserver
from .float_scale import EmaObserver, OmseObserver, PercentileObserver
from .minmax import MinmaxObserver
from .ptf import PtfObserver

str2observer = {cls.__name__[:-len('Observer')].lower(): cls
                for cls in (MinmaxObserver, EmaObserver, OmseObserver, PercentileObserver, PtfObserver)}


def build_observer(observer_str, module_type, bit_type, calibration_mode):
    try:
        cls = str2observer[observer_str]
    except KeyError:
        raise KeyError('unknown observer %r (have: %s)' % (observer_str, ', '.join(sorted(str2observer)))) from None
    return cls(module_type, bit_type, calibration_mode)
