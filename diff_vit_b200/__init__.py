"""diff_vit_b200: B200-native integer inference path for P2-ViT quantized Vision Transformers.

Drop-in for the reference's ``from models import *`` / ``from config import Config`` surface
(reference: models/__init__.py:2-5): the ptq operators, the DeiT/ViT factories, ``Config`` and
``str2model``.
"""
from .config import Config
from .ptq import BIT_TYPE_DICT, BIT_TYPE_LIST, BitType, QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear
from .vit_fquant import (Attention, Block, VisionTransformer, deit_base_patch16_224, deit_small_patch16_224,
                         deit_tiny_patch16_224, vit_base_patch16_224, vit_large_patch16_224)
from .layers_quant import Mlp, PatchEmbed
from .weights import build_model, load_weights_from_npz
from .swin_quant import (SwinTransformer, swin_base_patch4_window7_224, swin_small_patch4_window7_224,
                         swin_tiny_patch4_window7_224)

_MODELS = {
    'deit_tiny': deit_tiny_patch16_224,
    'deit_small': deit_small_patch16_224,
    'deit_base': deit_base_patch16_224,
    'vit_base': vit_base_patch16_224,
    'vit_large': vit_large_patch16_224,
    'swin_tiny': swin_tiny_patch4_window7_224,
    'swin_small': swin_small_patch4_window7_224,
    'swin_base': swin_base_patch4_window7_224,
}


def str2model(name):
    """reference: test_quant.py:56-68."""
    return _MODELS[name]


def calibrate_model(model, batches):
    """One calibration pass per batch, quantization parameters fixed on the last one, then
    ``model_quant()`` (reference flow: test_quant.py:222-249, model_utility.py:121-175)."""
    import torch
    batches = list(batches)
    model.model_open_calibrate()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False        # the weight search scores fp32 conv outputs, as the reference does
    try:
        with torch.no_grad():
            for i, x in enumerate(batches):
                if i == len(batches) - 1:
                    model.model_open_last_calibrate()
                _, _, distance = model(x, plot=False)
                model.global_distance = distance   # per-layer weight distances of the last pass (search.omega)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    model.model_close_calibrate()
    model.model_quant()
    return model
