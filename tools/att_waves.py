"""Attention kernel time vs number of CTAs (wave quantisation check)."""
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
eng = model.integer_engine()
bp = eng.bound([8] * 50)
lib = _cabi.lib()
att = bp.blocks[0].attn
st = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N, H, D = 197, 6, 384
for B in (74, 148, 222, 240, 256, 296, 370):
    qkv = torch.randint(-60, 60, (B * N, 3 * D), dtype=torch.int8, device=dev)
    out = torch.empty(B * N, D, dtype=torch.int8, device=dev)
    with torch.cuda.stream(st):
        for _ in range(3):
            _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), B, N, H, C.byref(att), st.cuda_stream))
        e0.record(st)
        for _ in range(20):
            _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), B, N, H, C.byref(att), st.cuda_stream))
        e1.record(st)
    st.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print('B=%d CTAs=%d waves=%.2f: %.1f us, %.4f us/CTA' % (B, B * H, B * H / 444.0, us, us / (B * H)))
