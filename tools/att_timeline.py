"""Phase timeline (SM clock cycles) of CTA 0's softmax warps in the tcgen05 attention kernel.

    python tools/att_timeline.py [batch]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_vit_b200 import _cabi  # noqa: E402
from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, heads = 197, 6
torch.manual_seed(0)
qkv = torch.randint(-8, 9, (b * n, 3 * heads * 64), dtype=torch.int8, device='cuda')
out = torch.empty(b * n, heads * 64, dtype=torch.int8, device='cuda')
p = AttentionPlan(score_mul=float(2.0 ** -8), score_zp=0.0, out_mul=2.0 ** -15 * 2.0 ** -1, out_zp=0.0, levels=16,
                  exp_lut=softmax_exp_lut(torch.tensor([2.0 ** -3])), in_zp=0.0)
lut = p.exp_lut.cuda()
c = _cabi.Attention()
c.score_mul, c.score_zp, c.out_mul, c.out_zp, c.softmax_levels = p.score_mul, p.score_zp, p.out_mul, p.out_zp, 16
c.in_zp, c.exp_lut, c.lut_sig_bits = 0.0, lut.data_ptr(), p.lut_sig_bits
lib = _cabi.lib()
tl = torch.zeros(12 * 16 * 8, dtype=torch.int64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), st))
_cabi.check(lib.p2v_attention_tc_set_timeline(tl.data_ptr()))
_cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(c), st))
torch.cuda.synchronize()
_cabi.check(lib.p2v_attention_tc_set_timeline(None))
t = tl.cpu().reshape(12, 16, 8)
t0 = int(t[t > 0].min())
names = ['wait S', 'S ready', 'pass1', 'pass2', 'pass3+P', 'O ready', 'epilogue', 'rebias']
print('cycles since the first stamp; columns: ' + ', '.join(names))
for i in range(min(5, (b * heads + 147) // 148)):
    for w in range(16):
        row = t[i, w]
        if int(row.max()) == 0:
            continue
        vals = [int(v) - t0 if int(v) else -1 for v in row]
        d = [vals[k] - vals[k - 1] if vals[k] >= 0 and vals[k - 1] >= 0 else 0 for k in range(1, 8)]
        print('item %2d warp %2d: start %7d | S wait %5d  p1 %5d  p2 %5d  p3 %5d  O wait %5d  epi %5d  rebias %5d'
              % ((i, w, vals[0]) + tuple(d)))
