"""Quick per-kernel CUDA-event timing of the DeiT-S b256 forward pieces through the C ABI (no ncu).
Usage: python tools/time_kernels.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diff_vit_b200 as dv
from diff_vit_b200 import _cabi

torch.manual_seed(0)
dev = torch.device('cuda', 0)
model = dv.deit_small_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
dv.calibrate_model(model, [torch.randn(8, 3, 224, 224, device=dev, generator=g)])
x = torch.randn(256, 3, 224, 224, device=dev, generator=g)
eng = model.integer_engine()
bits = [8] * 50
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3):
        eng.forward_into(x, bits)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        eng.forward_into(x, bits)
    e1.record(st)
st.synchronize()
print('forward (graph) ms/step: %.3f' % (e0.elapsed_time(e1) / 10))

# attention alone on the engine's own qkv buffer
bp = eng.bound(bits)
lib = _cabi.lib()
plan = bp.plan
B, N, H, D = 256, 197, 6, 384
qkv = torch.randint(-60, 60, (B * N, 3 * D), dtype=torch.int8, device=dev)
out = torch.empty(B * N, D, dtype=torch.int8, device=dev)
att = bp.blocks[0].attn
with torch.cuda.stream(st):
    for _ in range(3):
        _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), B, N, H, C.byref(att), st.cuda_stream))
    e0.record(st)
    for _ in range(20):
        _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), B, N, H, C.byref(att), st.cuda_stream))
    e1.record(st)
st.synchronize()
print('attention us: %.1f' % (e0.elapsed_time(e1) / 20 * 1e3))
