import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import diff_vit_b200 as dv
from diff_vit_b200.plan import extract_state
from oracle import fakequant_forward as orc
torch.manual_seed(0)
model = dv.deit_base_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().cuda()
g = torch.Generator(device='cuda').manual_seed(3)
dv.calibrate_model(model, [torch.randn(4, 3, 224, 224, device='cuda', generator=g)])
x = torch.randn(3, 3, 224, 224, device='cuda', generator=g)
state = extract_state(model)
got, dump = model.integer_engine().forward_dump(x, [8] * 50)
want, ref = orc.forward(state, x.cpu(), [8] * 50, capture=True)
n = 0
for k, v in ref.items():
    if k not in dump: continue
    r = v.numpy().astype(np.int64)
    if k == 'ln/norm': r = r[:, 0]
    d = np.abs(dump[k].astype(np.int64).reshape(r.shape) - r)
    if d.max() > 0:
        print('%-45s differ %8.4f%%  max %d' % (k, 100 * (d != 0).mean(), d.max()))
        n += 1
        if n > 14: break
