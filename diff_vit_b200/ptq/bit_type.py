"""Integer code ranges used by the P2-ViT quantizers.

Mirrors the reference's bit-type registry (reference: models/ptq/bit_type.py:7-57): the live list is
uint3, uint4, int4, int8, uint8, looked up by name through ``BIT_TYPE_DICT``.
"""


class BitType:
    """A (bits, signed) pair with its closed code range [lower_bound, upper_bound]."""

    __slots__ = ('bits', 'signed', 'name')

    def __init__(self, bits, signed, name=None):
        self.bits = int(bits)
        self.signed = bool(signed)
        self.name = name if name is not None else ('int' if signed else 'uint') + str(bits)

    @property
    def upper_bound(self):
        return (1 << (self.bits - 1)) - 1 if self.signed else (1 << self.bits) - 1

    @property
    def lower_bound(self):
        return -(1 << (self.bits - 1)) if self.signed else 0

    @property
    def range(self):
        return 1 << self.bits

    def update_name(self):
        self.name = ('int' if self.signed else 'uint') + str(self.bits)

    def __repr__(self):
        return 'BitType(%s)' % self.name


# Order matters: QLinear/QConv2d calibrate every entry except uint8 in this order, and the
# per-layer `global_distance` rows follow it (reference: models/ptq/layers.py:148-170).
BIT_TYPE_LIST = [
    BitType(3, False, 'uint3'),
    BitType(4, False, 'uint4'),
    BitType(4, True, 'int4'),
    BitType(8, True, 'int8'),
    BitType(8, False, 'uint8'),
]
BIT_TYPE_DICT = {b.name: b for b in BIT_TYPE_LIST}
