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


@pytest.mark.parametrize('method', ['ema', 'percentile', 'omse'])
def test_float_scale_observers_through_the_integer_plan(micro_golden, method):
    """BASELINE config 3 style: FQ-ViT float-scale activation observers.  The activation scales are no longer
    powers of two, so the reference's fp32 GEMM is not an exact integer product and codes can only agree
    within the allowance: each layer differs from the oracle by a few codes at rounding ties, and those
    propagate.  The plan must take its general (IEEE division) paths and stay close end to end.  `omse` adds
    asymmetric zero points to every plain activation quantizer: GEMM inputs fold z * sum_k w_nk into the bias, the
    attention products are completed with row / key sums (the reference's own omse observer raises TypeError as
    shipped; the mirror accepts the keywords QAct passes)."""
    import diff_vit_b200 as dv
    from conftest import build_micro
    from diff_vit_b200.plan import extract_state
    z = micro_golden
    model = build_micro(z)
    model.cfg = dv.Config(True, True, method)
    fresh = dv.VisionTransformer(img_size=48, patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                 qkv_bias=True, norm_layer=__import__('functools').partial(dv.QIntLayerNorm, eps=1e-6),
                                 input_quant=True, cfg=dv.Config(True, True, method), num_classes=16).eval()
    fresh.load_state_dict(model.state_dict())
    dv.calibrate_model(fresh, [torch.from_numpy(z['x_calib'])])
    state = extract_state(fresh)
    plan = build_plan(state, [8] * 10)
    assert not plan.blocks[0].norm1.pot and not (plan.blocks[0].qkv.flags & 4)     # general paths
    if method == 'omse':
        assert plan.blocks[0].attn.in_zp != 0.0 and plan.blocks[0].attn.score_zp != 0.0 and plan.blocks[0].fc1.out_zp != 0.0
    x = torch.from_numpy(z['x_eval'])
    ref_logits, ref = orc.forward(state, x, [8] * 10, capture=True)
    logits, codes = hostmath.run_plan(plan, z['x_eval'])
    first = ['act/qact_input', 'act/patch_embed.qact', 'act/qact1', 'ln/blocks.0.norm1', 'act/blocks.0.attn.qact0']
    for k in first:      # before any inexact fp32 accumulation the codes are still bit-exact
        np.testing.assert_array_equal(ref[k].numpy().astype(np.int64).reshape(codes[k].shape), codes[k].astype(np.int64), err_msg=k)
    total = bad = 0
    for k, v in codes.items():
        g = ref[k].numpy().astype(np.int64)
        g = g[:, 0] if k == 'ln/norm' else g
        d = g - v.astype(np.int64).reshape(g.shape)
        total += d.size
        bad += int((d != 0).sum())
    assert bad / total < 0.02
    lsb = float(state['act']['act_out'][0])
    assert np.abs(ref_logits.numpy() - logits).max() <= 8 * lsb


def test_plan_round_trips_through_npz(micro_state, micro_golden, tmp_path):
    from diff_vit_b200.plan import load_plan, save_plan
    plan = build_plan(micro_state, [int(v) for v in micro_golden['mixed/bit_config']])
    path = str(tmp_path / 'plan.npz')
    save_plan(plan, path, pack4=False)
    back = load_plan(path)
    assert back.bit_config == plan.bit_config and back.arch == plan.arch
    a, b = list(plan.tensors()), list(back.tensors())
    assert len(a) == len(b) > 50 and all(x.dtype == y.dtype and torch.equal(x, y) for x, y in zip(a, b))
    # default: 4-bit layers travel int4-packed (two codes per byte) and are expanded again when used
    from diff_vit_b200.plan import _linear_plans, pack_int4, unpack_int4
    save_plan(plan, path)
    back = load_plan(path)
    packed = [(o, l) for o, l in zip(_linear_plans(plan), _linear_plans(back)) if o.bits == 4]
    assert len(packed) >= 2 and all(l.w is None and l.w4.dtype == torch.uint8 and l.w4.shape[1] * 2 == o.w.shape[1]
                                    and torch.equal(l.codes(), o.w) for o, l in packed)
    assert all(l.w4 is None and torch.equal(l.w, o.w) for o, l in zip(_linear_plans(plan), _linear_plans(back)) if o.bits == 8)
    assert all(o.w is not None for o in _linear_plans(plan))                       # the caller's plan is untouched
    edge = torch.tensor([[-8, 7, -1, 0], [3, -4, 5, -6]], dtype=torch.int8)
    assert torch.equal(unpack_int4(pack_int4(edge)), edge) and pack_int4(edge).tolist() == [[0x78, 0x0F], [0xC3, 0xA5]]
    with pytest.raises(ValueError):
        pack_int4(torch.tensor([[8, 0]], dtype=torch.int8))
    assert back.blocks[1].attn.out_mul == plan.blocks[1].attn.out_mul and back.blocks[0].norm2.pot == plan.blocks[0].norm2.pot
    logits, _ = hostmath.run_plan(back, micro_golden['x_eval'])
    np.testing.assert_array_equal(logits, micro_golden['mixed/logits'])
