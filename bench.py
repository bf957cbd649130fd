#!/usr/bin/env python
"""Headline benchmark: images/s of the W8A8 PoT DeiT-S quantized forward at batch 256 per B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU fake-quant path (port)

One "step" = one quantized forward of ONE batch of 256 synthetic images (calibration excluded, as in
BASELINE.md).  N > 1 is launched by torchrun, one rank per GPU: the 256-image batch is sharded (256 / N images per
GPU, `dist.shard`) and the step ends with the NCCL all-gather of the logits (`dist.gather_logits`), i.e. strong
scaling of BASELINE config 2.  The round-1 measurement (256 images per GPU, no data-path collective) is reported
beside it under "weak".  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = 'deit_small'
BATCH = 256
GOP_PER_IMAGE = 9.198      # 2 * 4 598 882 304 MAC incl. QK^T / AV
INT8_NOMINAL_TOPS = 4500.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default=MODEL)
    ap.add_argument('--batch', type=int, default=BATCH, help='global batch per step (sharded over the GPUs)')
    ap.add_argument('--calib-batch', type=int, default=32)
    ap.add_argument('--e2e-steps', type=int, default=40, help='batches of the host-buffer serving loop (its un-overlapped first upload is inside the timed region)')
    ap.add_argument('--cpu-sample', type=int, default=32, help='images in the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-swin', action='store_true', help='skip the config-5 (swin_tiny b128) side measurement at N = 1')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


def kernel_metrics():
    """{kernel: {dram_bytes, tensor_pipe_pct, ...}} from the newest profiles/r*_kernel_metrics.json (written by
    tools/ncu_to_metrics.py from `ncu --set full` captures), and the file name."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_kernel_metrics.json')))
    if not files:
        return {}, None
    with open(files[-1]) as f:
        return json.load(f), os.path.relpath(files[-1], ROOT)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = [r for t, r in self.rows if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.05)]
        if not rows:       # a timed region shorter than one sampling period: the nearest samples
            rows = [r for _, r in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def build_calibrated(args, device):
    import torch
    import diff_vit_b200 as dv
    torch.manual_seed(0)
    model = dv.str2model(args.model)(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().to(device)
    g = torch.Generator(device=device).manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    calib = torch.randn(args.calib_batch, 3, 224, 224, device=device, generator=g)
    t0 = time.time()
    dv.calibrate_model(model, [calib])
    return model, time.time() - t0


def cpu_port_rate(state, nbits, sample, steps, warmup, threads):
    """images/s of the oracle's fp32 fake-quant forward (the reference's CPU path, restated) on `sample` images."""
    import torch
    from oracle import fakequant_forward as orc
    torch.set_num_threads(threads)
    x = torch.randn(sample, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.forward(state, x, nbits)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sample * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    """The reference arm: the reference's own CPU fake-quant implementation of the path (oracle port; the
    reference is Python and /root/reference is not on the GPU box), all host threads, rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import torch
    from diff_vit_b200.plan import extract_state
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model, calib_s = build_calibrated(argparse.Namespace(**{**vars(args), 'calib_batch': 8}), torch.device('cpu'))
    state = extract_state(model)
    sample = 16
    rate, per_step = cpu_port_rate(state, [8] * 50, sample, args.steps, args.warmup, threads)
    print(json.dumps({
        'impl': 'reference', 'metric': 'images/sec W8A8 PoT DeiT-S b256', 'value': round(rate, 3), 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(per_step * 1e3, 3),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 fake-quant (int8 grid)',
        'data': 'synthetic',
        'config': {'workload': '%s W8A8 PoT minmax quantized forward, bit_config [8]*50' % args.model,
                   'per_step_sample': '%d images of the 256-image batch' % sample},
        'cpu_baseline': {'value': round(rate, 3), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d-image forward per step, oracle/fakequant_forward.py (restates '
                                   'models/vit_fquant.py:700-799)' % sample},
        'e2e': {'value': round(rate, 3), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'calibration_s': round(calib_s, 2)}))


def time_gemm(lib_mod, m, n, k, flags, iters, stream_obj):
    """Average device time (ms) of the fc1-shaped tensor-core GEMM with its GELU + re-quant epilogue."""
    import ctypes as C
    import torch
    dev = 'cuda'
    a = torch.randint(-128, 128, (m, k), dtype=torch.int8, device=dev)
    w = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=dev)
    out = torch.empty(m, n, dtype=torch.int8, device=dev)
    vec = lambda v: torch.full((n,), v, dtype=torch.float32, device=dev)
    # random int8 x int8 over k terms has std ~ 5400 sqrt(k): 2^-17 puts the pre-GELU values at std ~ 0.8, like the
    # model's fc1 outputs, so the epilogue sees a realistic mix of GELU arguments (not only saturated ones)
    acc, bias, osc, ors = vec(2.0 ** -17), vec(0.01), vec(2.0 ** -5), vec(2.0 ** 5)
    e = lib_mod.Epilogue()
    e.acc_scale, e.bias, e.out_scale, e.out_rscale = acc.data_ptr(), bias.data_ptr(), osc.data_ptr(), ors.data_ptr()
    e.flags = flags
    lib = lib_mod.lib()
    with torch.cuda.stream(stream_obj):
        for _ in range(3):
            lib_mod.check(lib.p2v_gemm_i8(a.data_ptr(), k, w.data_ptr(), out.data_ptr(), n, m, n, k, C.byref(e),
                                          stream_obj.cuda_stream))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream_obj)
        for _ in range(iters):
            lib_mod.check(lib.p2v_gemm_i8(a.data_ptr(), k, w.data_ptr(), out.data_ptr(), n, m, n, k, C.byref(e),
                                          stream_obj.cuda_stream))
        t1.record(stream_obj)
    stream_obj.synchronize()
    return t0.elapsed_time(t1) / iters


def time_attention(lib_mod, bound, qkv, b, n, heads, iters, stream_obj):
    """Average device time (ms) of the fused integer attention of block 0 on the model's own q/k/v codes."""
    import ctypes as C
    import torch
    out = torch.empty(b * n, heads * 64, dtype=torch.int8, device=qkv.device)
    att = bound.blocks[0].attn
    lib = lib_mod.lib()
    with torch.cuda.stream(stream_obj):
        for _ in range(3):
            lib_mod.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(att), stream_obj.cuda_stream))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream_obj)
        for _ in range(iters):
            lib_mod.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(att), stream_obj.cuda_stream))
        t1.record(stream_obj)
    stream_obj.synchronize()
    return t0.elapsed_time(t1) / iters


def time_layernorm(lib_mod, bound, rows, d, iters, stream_obj):
    """Average device time (ms) of block 0's integer LayerNorm (norm1) on random int8 residual codes."""
    import ctypes as C
    import torch
    x = torch.randint(-128, 128, (rows, d), dtype=torch.int8, device='cuda')
    out = torch.empty_like(x)
    ln = bound.blocks[0].norm1
    lib = lib_mod.lib()
    call = lambda: lib_mod.check(lib.p2v_layernorm_int(x.data_ptr(), d, out.data_ptr(), None, rows, d, C.byref(ln),
                                                       stream_obj.cuda_stream))
    with torch.cuda.stream(stream_obj):
        for _ in range(3):
            call()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream_obj)
        for _ in range(iters):
            call()
        t1.record(stream_obj)
    stream_obj.synchronize()
    return t0.elapsed_time(t1) / iters


def swin_side_measurement(device, iters=20):
    """BASELINE config 5 beside the headline (N = 1 only; not part of the timed region above): swin_tiny, W8A8 PoT,
    batch 128 synthetic images, random-init weights, on the Swin integer engine (diff_vit_b200/swin_engine.py) -
    CUDA-graph replay of its launch sequence, device-resident fp32 input, CUDA events."""
    import torch
    import diff_vit_b200 as dv
    torch.manual_seed(0)
    m = dv.swin_tiny_patch4_window7_224(cfg=dv.Config(True, True, 'minmax')).eval().to(device)
    g = torch.Generator(device=device).manual_seed(1)
    dv.calibrate_model(m, [torch.randn(8, 3, 224, 224, device=device, generator=g)])
    x = torch.randn(128, 3, 224, 224, device=device, generator=g)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        if m._engine_off is not None:
            return {'unavailable': m._engine_off}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        e0.record()
        for _ in range(iters):
            m(x)
        e1.record()
        torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / iters
    # integer operations per image: every linear over ALL its tokens (the reference's FLOPs list counts the window
    # attention's linears per window) plus the two products of the window attention
    pe = m.patch_embed
    macs = pe.proj.in_channels * pe.patch_size[0] ** 2 * m.embed_dim * pe.num_patches + m.num_features * m.num_classes
    for layer in m.layers:
        t, c = layer.input_resolution[0] * layer.input_resolution[1], layer.dim
        for blk in layer.blocks:
            macs += t * c * (3 * c + c + 2 * blk.mlp.fc1.out_features) + 2 * t * blk.window_size ** 2 * c
        if layer.downsample is not None:
            macs += (t // 4) * 4 * c * 2 * c
    gop = 2.0 * macs / 1e9
    return {'metric': 'images/sec W8A8 PoT swin_tiny b128 (BASELINE config 5)', 'value': round(128 / ms * 1e3, 1),
            'unit': 'images/s', 'ms_per_step': round(ms, 4), 'steps': iters, 'launches_per_step': m.integer_engine().launches,
            'gop_per_image': round(gop, 3), 'achieved_tops': round(128 / ms * gop, 1),
            'note': 'SwinTransformer.forward on the integer engine: graph replay incl. the device-side copy of the input '
                    'into the graph buffer and of the logits out of it; calibration (8 images) excluded'}


def int8_library_peak(device, n=8192, iters=10):
    """SURVEY 8d asks for three int8 denominators: nominal 4.5 POP/s, 2 x the measured bf16 burst, and - when the box
    offers it - the library's own int8 GEMM measured in the same run: torch._int_mm (cuBLASLt) on n^3.  A measurement
    of the denominator only; nothing on the product path calls it."""
    import torch
    try:
        a = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=device)
        b = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=device).t()      # column-major second operand
        for _ in range(3):
            torch._int_mm(a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        e0.record()
        for _ in range(iters):
            torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / iters
        return {'tops': round(2.0 * n ** 3 / (ms * 1e-3) / 1e12, 1), 'source': 'torch._int_mm %d^3 (cuBLASLt int8, s32 out), %d launches' % (n, iters)}
    except Exception as e:
        return {'tops': None, 'source': 'torch._int_mm unavailable: %s' % type(e).__name__}


def bind_to_gpu_numa_node(local):
    """One process per GPU: run (and first-touch the pinned staging buffers) on the CPUs NVML reports as local to
    this GPU, so that eight ranks do not pull their 154 MB batches across the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return 'cpus %d-%d' % (min(cpus), max(cpus))
    except Exception as exc:   # no NVML / restricted container: keep the inherited affinity
        return 'unchanged (%s)' % type(exc).__name__
    return 'unchanged'


MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)     # the loaders' normalisation (test_quant.py:96-110)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diff_vit_b200 import _cabi
    from diff_vit_b200 import dist as dvd
    from diff_vit_b200.plan import extract_state

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the quantized forward has no CPU fallback '
                         '(use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    _cabi.check(_cabi.lib().p2v_check_device(local))

    model, calib_s = build_calibrated(args, device)
    bits = [8] * (4 * model.depth + 2)
    eng = model.integer_engine()
    bound = eng.bound(bits)
    classes = model.num_classes
    # ONE global batch, identical on every rank (same seed), of which a rank keeps its shard
    g = torch.Generator(device=device).manual_seed(1)
    x_global = torch.randn(args.batch, 3, 224, 224, device=device, generator=g)   # 154 MB fp32 > 126 MB L2
    x = dvd.shard(x_global, rank, world).contiguous() if world > 1 else x_global
    per_rank = x.shape[0]
    assert per_rank * world == args.batch, 'the global batch must divide by the number of GPUs'
    gathered = torch.empty(args.batch, classes, dtype=torch.float32, device=device)
    stream = torch.cuda.Stream(device)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step(xin):
        """One step of the sharded job: this rank's forward, then the logits of all ranks on every rank."""
        out = eng.forward_into(xin, bits)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)        # NCCL, ordered after the forward on this stream
            return gathered
        return out

    # Timing hygiene: a 256-image fp32 batch (154 MB) is larger than the 126 MB L2, so back-to-back steps already start
    # cold.  A shard of it is not (32 images = 19 MB at 8 GPUs): then every step is timed on its own, with a write of
    # a 256 MB buffer (an L2 flush) between steps, outside the timed intervals.
    flush_needed = x.numel() * 4 < 140e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if flush_needed else None

    def timed(fn, steps, warmup, flush):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
        stream.synchronize()
        barrier()
        if not flush:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                t0.record(stream)
                for _ in range(steps):
                    out = fn()
                t1.record(stream)
            stream.synchronize()
            total = t0.elapsed_time(t1)
        else:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            with torch.cuda.stream(stream):
                for a, b in evs:
                    flush_buf.zero_()
                    a.record(stream)
                    out = fn()
                    b.record(stream)
            stream.synchronize()
            total = sum(a.elapsed_time(b) for a, b in evs)
        barrier()
        ms = torch.tensor([total], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(warm):
            step(x)
    stream.synchronize()
    barrier()
    sampler.mark_begin()
    total_ms, logits = timed(lambda: step(x), args.steps, 0, flush_needed)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    value = args.batch * args.steps / (total_ms * 1e-3)
    logits = logits.clone()
    if world > 1:       # the sharded job gives the rows of the unsharded forward of the same batch
        whole = eng.forward_into(x_global, bits)
        assert torch.equal(logits, whole), 'sharded + gathered logits differ from the single-GPU forward'

    # weak scaling beside it (round 1's number): every GPU its own full batch, no data-path collective
    weak = None
    if world > 1:
        weak_ms, _ = timed(lambda: eng.forward_into(x_global, bits), max(args.steps // 2, 5), 3, False)
        weak = {'value': round(world * args.batch * max(args.steps // 2, 5) / (weak_ms * 1e-3), 1), 'unit': 'images/s',
                'per_gpu_batch': args.batch, 'global_batch': args.batch * world, 'scaling': 'weak',
                'note': 'every rank forwards its own %d-image batch; no collective in the timed region' % args.batch}

    # ---- end to end through the public host-buffer serving call: every step uploads this rank's shard as 8-bit
    # pixels from pinned host memory (the device applies the loader's normalisation), runs the forward, gathers the
    # logits over NCCL and reads them back to the host; copies overlap the previous / next forward ----
    mean_t = torch.tensor(MEAN, device=device).reshape(1, 3, 1, 1)
    std_t = torch.tensor(STD, device=device).reshape(1, 3, 1, 1)
    pix = ((x * std_t + mean_t) * 255).round().clamp(0, 255).to(torch.uint8)
    x_norm = (pix.float().div(255) - mean_t) / std_t                       # what ToTensor + Normalize hand the model
    ref_logits = eng.forward_into(x_norm, bits).clone()
    x_hosts = [torch.empty(pix.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for xh in x_hosts:
        xh.copy_(pix)
    n_e2e = max(args.e2e_steps, 2)
    out_rows = args.batch if world > 1 else per_rank
    logits_hosts = [torch.empty(out_rows, classes, dtype=torch.float32).pin_memory() for _ in range(n_e2e)]
    batches = [x_hosts[i & 1] for i in range(n_e2e)]
    gather_bufs = [torch.empty(args.batch, classes, dtype=torch.float32, device=device) for _ in range(2)]

    def after(out, slot):
        if world == 1:
            return out
        dist.all_gather_into_tensor(gather_bufs[slot], out)
        return gather_bufs[slot]

    with torch.cuda.stream(stream):
        eng.forward_host_pipelined(batches[:2], logits_hosts[:2], bits, mean=MEAN, std=STD, after_forward=after)
    barrier()
    w0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.forward_host_pipelined(batches, logits_hosts, bits, mean=MEAN, std=STD, after_forward=after)
        e1.record(stream)
    stream.synchronize()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)], device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = args.batch * n_e2e / (float(e2e_ms.item()) * 1e-3)
    mine = slice(rank * per_rank, (rank + 1) * per_rank) if world > 1 else slice(None)
    for lh in logits_hosts:
        assert torch.equal(lh[mine].to(device), ref_logits), 'host-buffer call disagrees with the device-resident call'
    x_host, logits_host = x_hosts[0], logits_hosts[0]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines, timed live with CUDA events on the launch stream -----------------------------------------
    # Dominant kernel by share of the step (ncu launch list, profiles/): the fused integer attention.  Algorithmic work
    # per launch = QK^T + PV = 4 * B * H * N^2 * 64 int8 ops.  Second: the fc1 GEMM.  The kernels are timed at the
    # full batch of 256 (what one GPU runs at N = 1); DRAM traffic and tensor-pipe utilisation come from the ncu
    # captures summarised in profiles/r*_kernel_metrics.json (tools/ncu_to_metrics.py), not from literals.
    hbm, bf16, src = peaks()
    peak = 2.0 * bf16
    peak_note = ('2 x %s bf16 burst (MEASURED_PEAKS.json has no int8 entry); nominal dense int8 %.0f' % (src, INT8_NOMINAL_TOPS))
    km, km_file = kernel_metrics()
    ntok = model.patch_embed.num_patches + 1
    m = args.batch * ntok
    d, hid, heads = model.embed_dim, model.blocks[0].mlp.fc1.out_features, model.num_heads
    _, dump = eng.forward_dump(x_global[:32].contiguous(), bits)      # block 0's real q/k/v codes, tiled to the batch
    qkv32 = torch.from_numpy(dump['act/blocks.0.attn.qact1']).reshape(32 * ntok, 3 * d)
    del dump
    qkv = qkv32.repeat((args.batch + 31) // 32, 1)[:m].contiguous().to(device)
    att_ms = time_attention(_cabi, bound, qkv, args.batch, ntok, heads, 20, stream)
    att_ops = 4.0 * args.batch * heads * ntok * ntok * 64
    att_tops = att_ops / (att_ms * 1e-3) / 1e12
    step_ms = total_ms / args.steps
    share = lambda ms_, count: round(ms_ * count / step_ms, 3) if world == 1 else None   # shares hold for the N = 1 step

    def ncu(name, key):
        for k, v in km.items():
            if k.startswith(name):
                return v.get(key)
        return None

    roofline = {'bound': 'tensor', 'kernel': 'attention_tc_kernel (tcgen05 kind::i8 QK^T + PV, TMEM accumulators, TMA; thread-per-row '
                'log-int-softmax), %d x %d heads x %d tokens' % (args.batch, heads, ntok),
                'achieved': round(att_tops, 2), 'peak': round(peak, 1), 'unit': 'TFLOP/s', 'frac': round(att_tops / peak, 4),
                'traffic': ncu('attention_tc_kernel', 'dram_bytes'), 'algorithmic_bytes': 4.0 * m * d,
                'tensor_pipe_pct': ncu('attention_tc_kernel', 'tensor_pipe_pct'), 'traffic_source': km_file,
                'us_per_launch': round(att_ms * 1e3, 1), 'share_of_step': share(att_ms, model.depth),
                'peak_source': peak_note,
                'note': 'not tensor bound: the softmax between the two products issues ~26 instructions per score element on '
                        'four warps per scheduler (ALU pipe, table lookups, TMEM reads); see DESIGN.md 4.3'}
    ln_ms = time_layernorm(_cabi, bound, m, d, 20, stream)
    hbm_peak = float(hbm)
    ln_gbs = 2.0 * m * d / (ln_ms * 1e-3) / 1e9
    roofline_ln = {'bound': 'hbm', 'kernel': 'layernorm_int_pot_kernel %d rows x %d (int8 in, int8 out)' % (m, d),
                   'achieved': round(ln_gbs, 1), 'peak': round(hbm_peak, 1), 'unit': 'GB/s', 'peak_source': '%s HBM copy bandwidth' % src,
                   'frac': round(ln_gbs / hbm_peak, 4), 'traffic': ncu('layernorm_int_pot_kernel', 'dram_bytes'),
                   'algorithmic_bytes': 2.0 * m * d, 'traffic_source': km_file,
                   'us_per_launch': round(ln_ms * 1e3, 1), 'share_of_step': share(ln_ms, 2 * model.depth + 1)}
    gemm_ms = time_gemm(_cabi, m, hid, d, _cabi.EPI_GELU | _cabi.EPI_OUT_POT, 20, stream)
    ops = 2.0 * m * hid * d
    achieved = ops / (gemm_ms * 1e-3) / 1e12
    roofline_gemm = {'bound': 'tensor', 'kernel': 'gemm_i8_bs_kernel<GELU|OUT_POT> fc1 %dx%dx%d (tcgen05 kind::i8)' % (m, hid, d),
                     'achieved': round(achieved, 2), 'peak': round(peak, 1), 'unit': 'TFLOP/s', 'frac': round(achieved / peak, 4),
                     'traffic': ncu('gemm_i8_bs_kernel<21>', 'dram_bytes'), 'algorithmic_bytes': float(m * d + hid * d + m * hid),
                     'tensor_pipe_pct': ncu('gemm_i8_bs_kernel<21>', 'tensor_pipe_pct'), 'traffic_source': km_file,
                     'us_per_launch': round(gemm_ms * 1e3, 1),
                     'share_of_step': share(gemm_ms, model.depth), 'peak_source': peak_note}
    tensor_pipe = {k: v.get('tensor_pipe_pct') for k, v in km.items() if v.get('tensor_pipe_pct') is not None}
    model_tops = value / world * GOP_PER_IMAGE / 1e3
    whole = {'achieved_tops': round(model_tops, 2), 'frac_of_peak': round(model_tops / peak, 4),
             'frac_of_nominal_int8': round(model_tops / INT8_NOMINAL_TOPS, 4), 'gop_per_image': GOP_PER_IMAGE,
             'note': 'per GPU'}
    if world == 1:
        lib_peak = int8_library_peak(device)
        whole['int8_library_peak'] = lib_peak
        if lib_peak['tops']:
            whole['frac_of_int8_library_peak'] = round(model_tops / lib_peak['tops'], 4)

    swin = None
    if world == 1 and not args.no_swin:
        try:
            swin = swin_side_measurement(device)
        except Exception as e:      # the headline line must not depend on the side measurement
            swin = {'unavailable': '%s: %s' % (type(e).__name__, e)}

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, per = cpu_port_rate(extract_state(model), bits, args.cpu_sample, 2, 1, threads)
        cpu = {'value': round(rate, 3), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
               'sample': '%d-image forward x2 after 1 warm-up, oracle/fakequant_forward.py (CPU restatement of '
                         'models/vit_fquant.py:700-799; /root/reference is not on the GPU box)' % args.cpu_sample}

    in_bytes = x_host.numel()           # uint8 pixels
    launches = bound.launches
    line = {
        'metric': 'images/sec W8A8 PoT DeiT-S b256', 'value': round(value, 1), 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warm, 'ms_per_step': round(total_ms / args.steps, 4),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'int8 (s32 accumulate)', 'data': 'synthetic',
        'config': {'workload': '%s W8A8 PoT minmax quantized forward, bit_config [8]*50, random-init weights' % args.model,
                   'global_batch': args.batch, 'per_gpu_batch': per_rank, 'parallelism': 'dp%d' % world,
                   'sharding': ('one %d-image batch sharded over %d GPUs, NCCL all-gather of the logits inside the step'
                                % (args.batch, world)) if world > 1 else 'single GPU, no collective',
                   'l2': ('per-GPU input %.0f MB fp32 < 126 MB L2: every step timed on its own, 256 MB L2 flush between steps'
                          % (x.numel() * 4 / 1e6)) if flush_needed else 'input batch %.0f MB fp32 > 126 MB L2' % (x.numel() * 4 / 1e6),
                   'calibration': 'randn(%d,3,224,224), %.1f s, excluded' % (args.calib_batch, calib_s)},
        'e2e': {'value': round(e2e_value, 1), 'unit': 'images/s', 'h2d_bytes_per_step': in_bytes * world,
                'd2h_bytes_per_step': logits_host.numel() * 4 * world, 'steps': n_e2e,
                'api': 'IntegerEngine.forward_host_pipelined on uint8 pixels + mean / std (p2v_vit_forward_u8): pinned H2D of '
                       'batch i+1 and D2H of the (gathered) logits i-1 overlap forward i on their own streams',
                'input': 'uint8 pixels [%d, 3, 224, 224] per GPU per step; the device applies (p / 255 - mean) / std' % per_rank,
                'host_affinity': numa_note},
        'gpu_launches': (warm + args.steps + n_e2e + 2 + (1 if world > 1 else 0)) * launches + 70,
        'launches_per_step': launches,
        'roofline': roofline, 'roofline_fc1_gemm': roofline_gemm, 'roofline_layernorm': roofline_ln,
        'tensor_pipe_util_pct': tensor_pipe, 'whole_model': whole, 'weak': weak, 'cpu_baseline': cpu,
        'clocks': clocks, 'config5_swin_tiny': swin,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
