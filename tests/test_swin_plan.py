"""The Swin integer plan (diff_vit_b200.swin_engine.build_swin_plan) executed on the host with the kernels' own scalar
arithmetic (tests/hostmath: GEMM epilogues and LayerNorm from csrc/p2v_math.cuh, the window attention kernel's integer
formulation restated in numpy) against the reference's golden codes of the micro Swin and against the oracle.  Pins the
plan builder - window permutations, shift-mask regions, relative-position bias, the exact fixed-point q scaling, the
2 x 2 merge gather - on a CPU-only box; the GPU tests then prove the kernels."""
import numpy as np
import pytest
import torch

import hostmath
from diff_vit_b200.swin_engine import build_swin_plan, num_linear_layers, window_permutation, shift_regions
from diff_vit_b200.swin_quant import extract_swin_state, shift_attention_mask
from oracle import swin_fakequant_forward as sorc
from test_swin_golden import swin_golden, swin_model  # noqa: F401  (fixtures)


def test_window_permutation_and_regions_match_the_reference_ops():
    for res, ws, shift in (((14, 14), 7, 0), ((14, 14), 7, 3), ((56, 56), 7, 3), ((7, 7), 7, 0)):
        perm = window_permutation(res, ws, shift).long()
        x = torch.arange(res[0] * res[1], dtype=torch.float32).view(1, res[0], res[1], 1)
        if shift:
            x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        want = sorc._window_partition(x, ws).reshape(-1)
        assert torch.equal(perm.float(), want)
        assert sorted(perm.tolist()) == list(range(res[0] * res[1]))
        reg = shift_regions(res, ws, shift)
        mask = shift_attention_mask(res, ws, shift)
        if shift == 0:
            assert reg is None and mask is None
        else:
            r = reg.view(-1, ws * ws).long()
            assert torch.equal((r[:, :, None] != r[:, None, :]).float() * -100.0, mask.permute(0, 2, 1))
            assert torch.equal((r[:, None, :] != r[:, :, None]).float() * -100.0, mask)


@pytest.mark.parametrize('tag', ['w8', 'w4'])
def test_swin_micro_plan_on_the_host_vs_reference_golden_and_oracle(swin_model, swin_golden, tag):
    z = swin_golden
    state = extract_swin_state(swin_model)
    n = num_linear_layers(state['arch'])
    bits = [8] * n if tag == 'w8' else [4] * n
    plan = build_swin_plan(state, bits)
    x = z['x_eval']
    logits, codes = hostmath.run_swin_plan(plan, x)
    # layer by layer on identical inputs, against the oracle with exact accumulation (the q * 32^-1/2 product is not
    # on an integer grid: an fp32 BLAS adds summation-order noise the integer engine does not have)
    want, ref = sorc.forward(state, torch.from_numpy(x), bits, capture=True, accum='fp64',
                             override={k: v for k, v in codes.items()})
    assert len(ref) >= 70
    total = 0
    for k, r in ref.items():
        assert k in codes, k
        g = codes[k].astype(np.int64).reshape(r.shape)
        np.testing.assert_array_equal(g, r.numpy().astype(np.int64), err_msg=k)
        total += g.size
    assert total > 500000
    np.testing.assert_array_equal(want.numpy(), logits)
    if tag == 'w8' and 'w8/logits' in z.files:
        # free-running against the reference's own run (fp32 BLAS accumulation): the first block and the logits
        for k in ('act/qact_input', 'act/patch_embed.qact', 'ln/layers.0.blocks.0.norm1', 'act/layers.0.blocks.0.attn.qact1',
                  'act/layers.0.blocks.0.attn.qact_attn1', 'act/layers.0.blocks.0.attn.qact2',
                  'softmax/layers.0.blocks.0.attn.log_int_softmax', 'act/layers.0.blocks.0.attn.qact3',
                  'act/layers.0.blocks.0.qact2', 'act/layers.0.blocks.0.qact4'):
            g = z['w8/' + k].astype(np.int64)
            d = np.abs(codes[k].astype(np.int64).reshape(g.shape) - g)
            assert d.max() <= 1 and (d != 0).mean() <= 1e-3, k
        lsb = float(state['act']['act_out'][0])
        assert np.abs(logits - z['w8/logits']).max() <= 2 * lsb


def test_swin_flops_list_and_engine_dispatch_rules(swin_golden):
    """`flops()` (what the engine path returns) equals the list a per-module forward accumulates; the engine is only
    considered for CUDA tensors of a quantized model, and is dropped by everything that can change its plan."""
    import diff_vit_b200 as dv
    from test_swin_golden import build_swin_micro
    model = build_swin_micro(swin_golden)
    x = torch.from_numpy(swin_golden['x_eval'])
    n = model.num_linear_layers()
    with torch.no_grad():
        _, flops, _ = model(x, [8] * n, False)                 # float graph
    assert flops == model.flops() and len(flops) == n
    assert model._integer_forward(x, [8] * n) is None          # not quantized, CPU tensor
    model._engine = object()
    model.model_quant()
    assert model._engine is None and model._int_active
    assert model._integer_forward(x, [8] * n) is None          # CPU tensor: never the engine
    model._engine = object()
    model.model_dequant()
    assert model._engine is None and not model._int_active
    model._engine = object()
    model.load_state_dict(model.state_dict())
    assert model._engine is None


def test_swin_plan_scope_is_enforced(swin_model):
    """What the Swin integer plan refuses (the model then keeps its per-module path): asymmetric quantizers, float
    activation grids where a power of two is assumed, a qact2 grid that makes the shift mask's -100 a non-integer."""
    import copy
    state = extract_swin_state(swin_model)
    n = num_linear_layers(state['arch'])
    build_swin_plan(state, [8] * n)                                        # the calibrated state itself is inside
    with pytest.raises(IndexError):
        build_swin_plan(state, [8] * 3)
    with pytest.raises(KeyError):
        build_swin_plan(state, [6] * n)                                    # the reference: BIT_TYPE_DICT['int6']
    pre = 'layers.0.blocks.1.attn'
    s = copy.deepcopy(state)
    sc, zp, lo, hi = s['act'][pre + '.qact2']
    s['act'][pre + '.qact2'] = (sc, zp + 3.0, lo, hi)
    with pytest.raises(NotImplementedError, match='symmetric'):
        build_swin_plan(s, [8] * n)
    s = copy.deepcopy(state)
    s['act'][pre + '.qact_attn1'] = (sc * 0.3, zp, lo, hi)
    with pytest.raises(NotImplementedError, match='power of two'):
        build_swin_plan(s, [8] * n)
    s = copy.deepcopy(state)
    s['act'][pre + '.qact2'] = (torch.full_like(sc, 8.0), zp, lo, hi)      # 100 / 8 = 12.5
    with pytest.raises(NotImplementedError):                               # (the integer exp or the mask check trips)
        build_swin_plan(s, [8] * n)


def test_swin_plan_round_trips_through_a_file(swin_model, swin_golden, tmp_path):
    """save_swin_plan / load_swin_plan: the integer plan on disk (4-bit layers int4-packed) executes to the same codes."""
    from diff_vit_b200.swin_engine import load_swin_plan, save_swin_plan
    state = extract_swin_state(swin_model)
    n = num_linear_layers(state['arch'])
    bits = [4 if i % 2 else 8 for i in range(n)]
    plan = build_swin_plan(state, bits)
    path = str(tmp_path / 'swin_plan.npz')
    save_swin_plan(plan, path)
    back = load_swin_plan(path)
    assert back.bit_config == tuple(bits) and back.stages[0].res == plan.stages[0].res
    assert back.stages[0].blocks[0].qkv.w is None and back.stages[0].blocks[0].qkv.w4 is not None     # 4-bit: packed nibbles
    assert back.patch_embed.w is not None and back.patch_embed.w4 is None                             # 8-bit: int8 codes
    x = swin_golden['x_eval']
    want_logits, want = hostmath.run_swin_plan(plan, x)
    got_logits, got = hostmath.run_swin_plan(back, x)
    assert set(got) == set(want)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    np.testing.assert_array_equal(got_logits, want_logits)
