"""swin_tiny, batch 128 (BASELINE config 5): the quantized forward on the integer engine (CUDA-graph replay, CUDA
events) and, for comparison, through the per-module operators.  python tools/time_swin.py [batch] [tiny|small|base]"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import diff_vit_b200 as dv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
NAME = 'swin_' + (sys.argv[2] if len(sys.argv) > 2 else 'tiny')
torch.manual_seed(0)
model = dv.str2model(NAME)(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().cuda()
g = torch.Generator(device='cuda').manual_seed(0)
dv.calibrate_model(model, [torch.randn(16, 3, 224, 224, device='cuda', generator=g)])
x = torch.randn(B, 3, 224, 224, device='cuda', generator=g)
with torch.no_grad():
    for _ in range(3):
        y = model(x)
    assert model._engine_off is None, model._engine_off
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(NAME + ' b%d quantized forward on the integer engine (%d launches, graph replay, incl. the input copy): '
          '%.2f ms, %.0f img/s' % (B, model.integer_engine().launches, ms, B / ms * 1e3))
    if NAME != 'swin_tiny':
        sys.exit(0)
    model.per_module = True
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print('swin_tiny b%d quantized forward through the per-module operators: %.1f ms, %.0f img/s' % (B, dt * 1e3, B / dt))
