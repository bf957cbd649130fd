"""Quantizers holding the calibrated scale / zero-point state (reference: models/ptq/quantizer/*)."""
from .base import BaseQuantizer
from .log2 import Log2Quantizer
from .uniform import UniformQuantizer

str2quantizer = {'uniform': UniformQuantizer, 'log2': Log2Quantizer}


def build_quantizer(quantizer_str, bit_type, observer, module_type):
    """reference: models/ptq/quantizer/build.py:8-10"""
    return str2quantizer[quantizer_str](bit_type, observer, module_type)
