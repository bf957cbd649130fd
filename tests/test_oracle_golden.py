"""Pins the CPU oracle (oracle/fakequant_forward.py) to outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import checksum
from oracle import fakequant_forward as orc


@pytest.mark.parametrize('tag', ['w8', 'w4', 'mixed'])
def test_micro_every_layer_bit_exact(micro_state, micro_golden, tag):
    z = micro_golden
    bc = {'w8': [8] * 10, 'w4': [4] * 10, 'mixed': list(z['mixed/bit_config'])}[tag]
    logits, codes = orc.forward(micro_state, torch.from_numpy(z['x_eval']), bc, capture=True)
    assert len(codes) == 34
    for k, v in codes.items():
        g = z['%s/%s' % (tag, k)]
        np.testing.assert_array_equal(g.astype(np.int64), v.numpy().astype(np.int64).reshape(g.shape), err_msg=k)
    np.testing.assert_array_equal(z[tag + '/logits'], logits.numpy())


def test_deit_tiny_c1_checksums(tiny_state, tiny_model, tiny_golden):
    z = tiny_golden
    logits, codes = orc.forward(tiny_state, tiny_model._c1_batch, [8] * 50, capture=True)
    np.testing.assert_array_equal(z['w8/logits'], logits.numpy())
    for k, v in codes.items():
        np.testing.assert_array_equal(z['w8/sum/' + k], checksum(v.numpy()), err_msg=k)
    for k in z.files:
        if k.startswith('w8/act/') or k.startswith('w8/softmax/'):
            g = z[k]
            got = codes[k[3:]].numpy()[:g.shape[0]]
            np.testing.assert_array_equal(g.astype(np.int64), got.astype(np.int64).reshape(g.shape), err_msg=k)


def test_softmax_extreme_rows():
    """Codes 0..15 and the 'zero' code, which random-init attention never produces."""
    s = torch.tensor([2.0 ** -4])
    x = torch.zeros(1, 1, 4, 64)
    x[0, 0, 0, 0] = 127 * s                      # one dominant score, everything else far below
    x[0, 0, 0, 1:] = -128 * s
    x[0, 0, 1, :] = torch.arange(64) * 2 * s      # a ramp
    x[0, 0, 2, :] = 5 * s                         # uniform
    x[0, 0, 3, ::2] = 100 * s
    codes, val = orc.log_int_softmax(x, s, 4)
    assert codes[0, 0, 0, 0] == 0 and codes[0, 0, 0, 1] == 16 and val[0, 0, 0, 1] == 0
    assert codes[0, 0, 2].unique().tolist() == [6]          # 1/64 -> 2^-6
    assert len(set(codes[0, 0, 1].tolist())) >= 8   # the ramp spans many codes
    assert codes.max() == 16 and codes.min() == 0


def test_fp64_accumulation_mode_equals_reference_on_power_of_two_grids(micro_state, micro_golden):
    """The oracle's 'fp64' accumulation mode (what the exact-accumulation kernels are held to when scales are
    floats, BASELINE config 3) changes nothing where the reference's fp32 sums are exact: on the power-of-two
    grids of the minmax observer it reproduces the reference's golden codes bit for bit."""
    z = micro_golden
    logits, codes = orc.forward(micro_state, torch.from_numpy(z['x_eval']), [8] * 10, capture=True, accum='fp64')
    for k, v in codes.items():
        g = z['w8/' + k]
        np.testing.assert_array_equal(g.astype(np.int64), v.numpy().astype(np.int64).reshape(g.shape), err_msg=k)
    np.testing.assert_array_equal(z['w8/logits'], logits.numpy())


def test_teacher_forcing_is_transparent_and_local(micro_state, micro_golden):
    """The oracle's `override` (teacher forcing, used by the GPU parity tests to compare every layer on identical
    inputs): feeding the oracle its own codes changes nothing; feeding it altered codes of one layer leaves that
    layer's RECORDED codes (computed from the layer's true inputs) untouched and moves only what is downstream."""
    z = micro_golden
    x = torch.from_numpy(z['x_eval'])
    logits, codes = orc.forward(micro_state, x, [8] * 10, capture=True)
    own = {k: v.numpy() for k, v in codes.items()}
    l2, c2 = orc.forward(micro_state, x, [8] * 10, capture=True, override=own)
    assert torch.equal(l2, logits) and all(torch.equal(c2[k], codes[k]) for k in codes)
    bumped = dict(own)
    key = 'act/blocks.0.mlp.qact1'
    bumped[key] = np.clip(own[key] + 1, -128, 127)
    l3, c3 = orc.forward(micro_state, x, [8] * 10, capture=True, override={key: bumped[key]})
    keys = list(codes)
    at = keys.index(key)
    assert all(torch.equal(c3[k], codes[k]) for k in keys[:at + 1])         # recorded = own evaluation
    assert not torch.equal(c3['act/blocks.0.mlp.qact2'], codes['act/blocks.0.mlp.qact2'])
    # the CLS-only form of the final LayerNorm codes is accepted
    l4, _ = orc.forward(micro_state, x, [8] * 10, override={'ln/norm': own['ln/norm'][:, 0]})
    assert torch.equal(l4, logits)
