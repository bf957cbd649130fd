"""One eager quantized swin_tiny forward (batch from argv, default 128) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/profile_swin.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
import diff_vit_b200 as dv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
model = dv.swin_tiny_patch4_window7_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().cuda()
g = torch.Generator(device='cuda').manual_seed(0)
dv.calibrate_model(model, [torch.randn(8, 3, 224, 224, device='cuda', generator=g)])
x = torch.randn(B, 3, 224, 224, device='cuda', generator=g)
eng = model.integer_engine()
with torch.no_grad():
    eng.forward(x, graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.forward(x, graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print('done')
