from .core import BaseQuantizer, Log2Quantizer, UniformQuantizer, build_quantizer, str2quantizer
