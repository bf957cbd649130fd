"""The package's vectorised calibration must reproduce the reference's calibrated state exactly
(scales, zero points, SmoothQuant channel scales, per-bit-width weight scales, layer distances)."""
import hashlib

import numpy as np
import torch

import diff_vit_b200 as dv


def _compare_scales(model, z):
    n = 0
    for name, m in model.named_modules():
        if isinstance(m, dv.QAct) and m.quantizer.scale is not None:
            np.testing.assert_array_equal(z['scale/' + name], m.quantizer.scale.numpy(), err_msg=name)
            np.testing.assert_array_equal(z['zp/' + name], m.quantizer.zero_point.numpy(), err_msg=name)
            n += 1
        if isinstance(m, (dv.QLinear, dv.QConv2d)):
            assert sorted(m.quantizer.dic_scale) == ['int4', 'int8', 'uint3', 'uint4']
            for bit, s in m.quantizer.dic_scale.items():
                np.testing.assert_array_equal(z['wscale/%s/%s' % (name, bit)].reshape(-1), s.numpy().reshape(-1),
                                              err_msg='%s %s' % (name, bit))
                n += 1
        if isinstance(m, (dv.Attention, dv.Mlp)):
            np.testing.assert_array_equal(z['cs/' + name], m.channel_scale.numpy(), err_msg=name)
            n += 1
    return n


def test_micro_calibration_matches_reference(micro_model, micro_golden):
    assert _compare_scales(micro_model, micro_golden) == 71


def test_micro_calibration_outputs(micro_golden):
    from conftest import build_micro
    model = build_micro(micro_golden)
    model.model_open_calibrate()
    model.model_open_last_calibrate()
    with torch.no_grad():
        out, flops, gd = model(torch.from_numpy(micro_golden['x_calib']), plot=False)
    np.testing.assert_array_equal(np.asarray(flops), micro_golden['calib/flops'])
    assert len(gd) == 9 and all(len(r) == 4 for r in gd)   # [uint3, uint4, int4, int8] per layer, head excluded
    got = np.asarray([[float(d) for d in r] for r in gd], dtype=np.float32)
    np.testing.assert_array_equal(got, micro_golden['calib/global_distance'])
    assert model.flops() == list(micro_golden['calib/flops'])


def test_deit_tiny_weights_and_calibration_match_reference(tiny_model, tiny_golden):
    h = hashlib.sha256()
    torch.manual_seed(0)
    fresh = dv.deit_tiny_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax'))
    sd = fresh.state_dict()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    assert h.hexdigest() == str(tiny_golden['sd_hash'])      # same RNG stream as the reference factory
    assert hashlib.sha256(tiny_model._c1_batch.numpy().tobytes()).hexdigest() == str(tiny_golden['x_hash'])
    assert _compare_scales(tiny_model, tiny_golden) == 351


def test_quantized_forward_requires_bit_config_and_gpu(micro_model, micro_golden):
    import pytest
    x = torch.from_numpy(micro_golden['x_eval'])
    with pytest.raises(ValueError):
        micro_model(x)                      # reference: bit_pool.index(None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            micro_model(x, [8] * 10, False)  # no CPU fallback for quantized arithmetic
