import time, torch
x = torch.empty(256, 3, 224, 224).pin_memory()
d = torch.empty_like(x, device='cuda')
for nstreams in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    chunks_h = x.chunk(nstreams); chunks_d = d.chunk(nstreams)
    for _ in range(2):
        for s, h, dd in zip(streams, chunks_h, chunks_d):
            with torch.cuda.stream(s): dd.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        for s, h, dd in zip(streams, chunks_h, chunks_d):
            with torch.cuda.stream(s): dd.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print('streams %d: %.2f ms  %.1f GB/s' % (nstreams, dt * 1e3, x.numel() * 4 / dt / 1e9))
