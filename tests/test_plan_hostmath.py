"""The integer plan (diff_vit_b200.plan) executed with the kernels' own scalar arithmetic compiled for
the host (tests/hostmath) must reproduce the reference's integer codes.  This pins the plan builder and
the fp32 operation order of csrc/p2v_math.cuh on a CPU-only box; the GPU tests then only have to prove
that the kernels index, tile and accumulate correctly."""
import numpy as np
import pytest
import torch

import hostmath
from diff_vit_b200.plan import build_plan, is_pot
from oracle import fakequant_forward as orc


@pytest.mark.parametrize('tag', ['w8', 'w4', 'mixed'])
def test_micro_plan_bit_exact(micro_state, micro_golden, tag):
    z = micro_golden
    bc = {'w8': [8] * 10, 'w4': [4] * 10, 'mixed': list(z['mixed/bit_config'])}[tag]
    plan = build_plan(micro_state, bc)
    logits, codes = hostmath.run_plan(plan, z['x_eval'])
    assert len(codes) == 32
    for k, v in codes.items():
        g = z['%s/%s' % (tag, k)].astype(np.int64)
        if k == 'ln/norm':
            g = g[:, 0]
        np.testing.assert_array_equal(g, v.astype(np.int64).reshape(g.shape), err_msg=k)
    np.testing.assert_array_equal(z[tag + '/logits'], logits)


def test_deit_tiny_plan_vs_oracle(tiny_state, tiny_model):
    """DeiT-T, 4 images: every layer of the integer plan against the oracle (<= 1 LSB on <= 0.1 %)."""
    x = tiny_model._c1_batch[:4]
    plan = build_plan(tiny_state, [8] * 50)
    assert all(b.norm1.pot and b.norm2.pot and (b.qkv.flags & 4) and not (b.proj.flags & 4) for b in plan.blocks)
    logits, codes = hostmath.run_plan(plan, x.numpy())
    ref_logits, ref = orc.forward(tiny_state, x, [8] * 50, capture=True)
    total = bad = 0
    for k, v in codes.items():
        g = ref[k].numpy().astype(np.int64)
        if k == 'ln/norm':
            g = g[:, 0]
        d = np.abs(g - v.astype(np.int64).reshape(g.shape))
        assert d.max() <= 1, k
        total += d.size
        bad += int((d != 0).sum())
    assert bad <= 1e-3 * total
    assert np.abs(ref_logits.numpy() - logits).max() <= float(tiny_state['act']['act_out'][0])


def test_plan_rejects_unsupported(micro_state):
    with pytest.raises(KeyError):
        build_plan(micro_state, [6] * 10)        # reference: BIT_TYPE_DICT['int6'] KeyError
    with pytest.raises(IndexError):
        build_plan(micro_state, [8] * 3)
    assert is_pot(torch.tensor([0.5, 2.0, 2.0 ** -20])) and not is_pot(torch.tensor([0.3]))
