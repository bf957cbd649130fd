"""Quantized operator surface (reference: models/ptq/__init__.py:2-3)."""
from .bit_type import BIT_TYPE_DICT, BIT_TYPE_LIST, BitType
from .layers import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear
