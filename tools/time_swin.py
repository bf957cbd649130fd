import sys, time, torch
sys.path.insert(0, '/root/repo')
import diff_vit_b200 as dv
torch.manual_seed(0)
model = dv.swin_tiny_patch4_window7_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().cuda()
g = torch.Generator(device='cuda').manual_seed(0)
dv.calibrate_model(model, [torch.randn(16, 3, 224, 224, device='cuda', generator=g)])
x = torch.randn(128, 3, 224, 224, device='cuda', generator=g)
with torch.no_grad():
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print('swin_tiny b128 quantized forward through the per-module operators: %.1f ms, %.0f img/s' % (dt * 1e3, 128 / dt))
