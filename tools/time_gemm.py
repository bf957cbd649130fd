"""CUDA-event timing of p2v_gemm_i8 on one shape (L2-warm, back to back): python tools/time_gemm.py M N K [flags] [f32]

flags: P2V_EPI_* bits (1 GELU, 2 RESIDUAL, 4 OUT_POT); a trailing `f32` also asks for the dequantized fp32 output (the
head GEMM's configuration).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_vit_b200 import _cabi as cabi  # noqa: E402


def main():
    m, n, k = (int(v) for v in sys.argv[1:4])
    flags = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    f32 = len(sys.argv) > 5
    rng = np.random.default_rng(0)
    a = torch.from_numpy(rng.integers(-128, 128, size=(m, k), dtype=np.int64).astype(np.int8)).cuda()
    w = torch.from_numpy(rng.integers(-100, 101, size=(n, k), dtype=np.int64).astype(np.int8)).cuda()
    vec = lambda v: torch.full((n,), v, dtype=torch.float32, device='cuda')
    keep = [vec(2.0 ** -12), vec(0.01), vec(2.0 ** -5), vec(2.0 ** 5), vec(2.0 ** -6), vec(2.0 ** -5)]
    res = torch.zeros(m, n, dtype=torch.int8, device='cuda')
    out = torch.zeros(m, n, dtype=torch.int8, device='cuda')
    of = torch.zeros(m, n, dtype=torch.float32, device='cuda')
    e = cabi.Epilogue()
    e.acc_scale, e.bias, e.out_scale, e.out_rscale = (t.data_ptr() for t in keep[:4])
    e.flags = flags | (cabi.EPI_OUT_F32 if f32 else 0)
    if f32:
        e.out_f32 = of.data_ptr()
    if flags & 2:
        e.res_scale, e.out2_scale, e.residual = keep[4].data_ptr(), keep[5].data_ptr(), res.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: cabi.check(cabi.lib().p2v_gemm_i8(a.data_ptr(), k, w.data_ptr(), out.data_ptr(), n, m, n, k, C.byref(e), st))
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    reps = 50
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        call()
    ev[1].record()
    torch.cuda.synchronize()
    print('gemm %d x %d x %d flags %d%s: %.2f us / launch' % (m, n, k, flags, ' +f32' if f32 else '', ev[0].elapsed_time(ev[1]) * 1e3 / reps))


if __name__ == '__main__':
    main()
