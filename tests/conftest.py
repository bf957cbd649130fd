import os
import sys
from functools import partial

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def build_micro(z):
    """The tiny 2-block / 2-head ViT of tests/golden/micro_minmax.npz with the fixture's weights."""
    import diff_vit_b200 as dv
    model = dv.VisionTransformer(img_size=48, patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                 qkv_bias=True, norm_layer=partial(dv.QIntLayerNorm, eps=1e-6), input_quant=True,
                                 cfg=dv.Config(True, True, 'minmax'), num_classes=16).eval()
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    model.load_state_dict(sd, strict=True)
    return model


@pytest.fixture(scope='session')
def micro_golden():
    return np.load(os.path.join(GOLDEN, 'micro_minmax.npz'))


@pytest.fixture(scope='session')
def tiny_golden():
    return np.load(os.path.join(GOLDEN, 'deit_tiny_c1.npz'))


@pytest.fixture(scope='session')
def micro_model(micro_golden):
    """Micro model calibrated on the CPU with the package's own (vectorised) calibration."""
    import diff_vit_b200 as dv
    model = build_micro(micro_golden)
    dv.calibrate_model(model, [torch.from_numpy(micro_golden['x_calib'])])
    return model


@pytest.fixture(scope='session')
def micro_state(micro_model):
    from diff_vit_b200.plan import extract_state
    return extract_state(micro_model)


@pytest.fixture(scope='session')
def tiny_model():
    """Config C1: deit_tiny, random init under seed 0, calibrated on randn(32,3,224,224) (CPU)."""
    import diff_vit_b200 as dv
    torch.manual_seed(0)
    model = dv.deit_tiny_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval()
    x = torch.randn(32, 3, 224, 224)
    dv.calibrate_model(model, [x])
    model._c1_batch = x
    return model


@pytest.fixture(scope='session')
def tiny_state(tiny_model):
    from diff_vit_b200.plan import extract_state
    return extract_state(tiny_model)


def checksum(a):
    """The int64 checksum pair tests/golden/make_golden.py stores for every code tensor."""
    flat = np.asarray(a).astype(np.int64).reshape(-1)
    w = (np.arange(flat.size, dtype=np.int64) % 65521) + 1
    return np.asarray([flat.sum(), (flat * w).sum()], dtype=np.int64)
