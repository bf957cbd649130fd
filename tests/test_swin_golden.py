"""BASELINE config 5 (Swin) on the CPU: the oracle restatement (oracle/swin_fakequant_forward.py) and the package's
graph + calibration (diff_vit_b200/swin_quant.py) against outputs of the reference's own Swin modules
(tests/golden/swin_micro.npz, produced by tests/golden/make_golden_swin.py with four call-plumbing shims)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import swin_fakequant_forward as sorc


@pytest.fixture(scope='module')
def swin_golden():
    return np.load(os.path.join(GOLDEN, 'swin_micro.npz'))


def build_swin_micro(z):
    import diff_vit_b200 as dv
    a = [int(v) for v in z['arch']]
    model = dv.SwinTransformer(img_size=a[0], patch_size=a[1], num_classes=a[2], embed_dim=a[3], window_size=a[4],
                               depths=tuple(a[5:7]), num_heads=tuple(a[7:9]), norm_layer=dv.QIntLayerNorm,
                               input_quant=True, cfg=dv.Config(True, True, 'minmax')).eval()
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    model.load_state_dict(sd, strict=True)       # same parameter / buffer names as the reference
    return model


@pytest.fixture(scope='module')
def swin_model(swin_golden):
    import diff_vit_b200 as dv
    model = build_swin_micro(swin_golden)
    dv.calibrate_model(model, [torch.from_numpy(swin_golden['x_calib'])])
    return model


def test_swin_calibration_reproduces_the_reference_scales(swin_model, swin_golden):
    import diff_vit_b200 as dv
    z = swin_golden
    bad = []
    for name, m in swin_model.named_modules():
        if isinstance(m, dv.QAct) and m.quantizer.scale is not None and 'mlp.qact0' not in name:
            if not np.array_equal(z['scale/' + name].reshape(-1), m.quantizer.scale.numpy().reshape(-1)):
                bad.append(name)
        if isinstance(m, (dv.QLinear, dv.QConv2d)) and not name.endswith('mlp.fc1'):
            for bit in ('int4', 'int8'):
                if not np.array_equal(z['wscale/%s/%s' % (name, bit)].reshape(-1),
                                      m.quantizer.dic_scale[bit].numpy().reshape(-1)):
                    bad.append(name + '/' + bit)
        if isinstance(m, dv.Mlp) and not np.array_equal(z['cs/' + name], m.best_scale[-1].numpy()):
            bad.append(name + '/cs')
    assert bad == [], bad
    assert sum(1 for k in z.files if k.startswith('scale/')) >= 60


def test_swin_oracle_every_layer_bit_exact(swin_model, swin_golden):
    from diff_vit_b200.swin_quant import extract_swin_state
    z = swin_golden
    state = extract_swin_state(swin_model)
    logits, codes = sorc.forward(state, torch.from_numpy(z['x_eval']), capture=True)
    keys = [k[3:] for k in z.files if k.startswith('w8/') and k != 'w8/logits']
    assert len(keys) >= 70 and set(keys) <= set(codes), sorted(set(keys) - set(codes))
    for k in keys:
        g = z['w8/' + k].astype(np.int64)
        np.testing.assert_array_equal(g, codes[k].numpy().astype(np.int64).reshape(g.shape), err_msg=k)
    np.testing.assert_array_equal(z['w8/logits'], logits.numpy())
    sm = np.concatenate([z['w8/' + k].reshape(-1) for k in keys if k.startswith('softmax/')])
    assert sm.min() == 0 and sm.max() == 16          # shifted windows: masked keys get probability 0
