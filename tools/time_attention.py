"""Device time of the two attention kernels on a DeiT-S sized layer (b x 6 heads x 197 tokens), CUDA events.

    python tools/time_attention.py [batch] [spread]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_vit_b200 import _cabi  # noqa: E402
from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    spread = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    n, heads = 197, 6
    torch.manual_seed(0)
    qkv = torch.randint(-spread, spread + 1, (b * n, 3 * heads * 64), dtype=torch.int8, device='cuda')
    out = torch.empty(b * n, heads * 64, dtype=torch.int8, device='cuda')
    p = AttentionPlan(score_mul=float(2.0 ** -9), score_zp=0.0, out_mul=2.0 ** -15 * 2.0 ** -1, out_zp=0.0, levels=16,
                      exp_lut=softmax_exp_lut(torch.tensor([2.0 ** -3])), in_zp=0.0)
    lut = p.exp_lut.cuda()
    lib = _cabi.lib()
    st = torch.cuda.Stream()
    res = {}
    for name, legacy in (('tcgen05', 0), ('mma.sync', 1)):
        c = _cabi.Attention()
        c.score_mul, c.score_zp, c.out_mul, c.out_zp, c.softmax_levels = p.score_mul, p.score_zp, p.out_mul, p.out_zp, 16
        c.in_zp, c.exp_lut, c.lut_sig_bits, c.force_legacy = 0.0, lut.data_ptr(), p.lut_sig_bits, legacy
        with torch.cuda.stream(st):
            for _ in range(3):
                _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), st.cuda_stream))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(st)
            for _ in range(20):
                _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), st.cuda_stream))
            t1.record(st)
        st.synchronize()
        res[name] = (t0.elapsed_time(t1) / 20 * 1e3, out.clone())
        print('%-9s %8.1f us per layer  (b=%d, %d items, spread %d)' % (name, res[name][0], b, b * heads, spread))
    print('outputs equal:', bool(torch.equal(res['tcgen05'][1], res['mma.sync'][1])))


if __name__ == '__main__':
    main()
