"""Profiling driver: calibrate DeiT-S on a small batch, then run a few quantized forwards (no CUDA graph)
so that ncu sees every kernel launch of a step.  Usage: python tools/profile_step.py [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diff_vit_b200 as dv

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
dev = torch.device('cuda', 0)
model = dv.deit_small_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
dv.calibrate_model(model, [torch.randn(8, 3, 224, 224, device=dev, generator=g)])
x = torch.randn(batch, 3, 224, 224, device=dev, generator=g)
eng = model.integer_engine()
for _ in range(steps):
    out = eng.forward_into(x, [8] * 50, use_graph=False)
torch.cuda.synchronize()
print('ok', float(out.abs().sum()))
