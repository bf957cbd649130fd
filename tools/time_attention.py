"""Device time of the two attention kernels on a DeiT-S sized layer (b x 6 heads x 197 tokens), CUDA events.

    python tools/time_attention.py [batch] [spread]        # random codes in [-spread, spread]
    python tools/time_attention.py [batch] real            # the calibrated deit_small of bench.py: real qkv codes and
                                                           # attention parameters of blocks 0, 5 and 11
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_vit_b200 import _cabi  # noqa: E402
from diff_vit_b200.plan import AttentionPlan, softmax_exp_lut  # noqa: E402


def time_both(qkv, b, n, heads, att, label):
    out = torch.empty(b * n, heads * 64, dtype=torch.int8, device='cuda')
    lib = _cabi.lib()
    st = torch.cuda.Stream()
    res = {}
    for name, legacy in (('tcgen05', 0), ('mma.sync', 1)):
        att.force_legacy = legacy
        with torch.cuda.stream(st):
            for _ in range(3):
                _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(att), st.cuda_stream))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(st)
            for _ in range(20):
                _cabi.check(lib.p2v_attention_int(qkv.data_ptr(), out.data_ptr(), b, n, heads, C.byref(att), st.cuda_stream))
            t1.record(st)
        st.synchronize()
        res[name] = (t0.elapsed_time(t1) / 20 * 1e3, out.clone())
    att.force_legacy = 0
    print('%-28s tcgen05 %7.1f us   mma.sync %7.1f us   outputs equal: %s'
          % (label, res['tcgen05'][0], res['mma.sync'][0], bool(torch.equal(res['tcgen05'][1], res['mma.sync'][1]))))


def main():
    if os.environ.get('P2V_ATT_SKEW'):
        _cabi.check(_cabi.lib().p2v_attention_tc_set_skew(int(os.environ['P2V_ATT_SKEW'])))
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    mode = sys.argv[2] if len(sys.argv) > 2 else '60'
    n, heads = int(os.environ.get('P2V_ATT_N', '197')), 6      # P2V_ATT_N: token count (what the ragged 197 = 6 x 32 + 5 costs)
    torch.manual_seed(0)
    if mode != 'real':
        spread = int(mode)
        qkv = torch.randint(-spread, spread + 1, (b * n, 3 * heads * 64), dtype=torch.int8, device='cuda')
        p = AttentionPlan(score_mul=float(2.0 ** -9), score_zp=0.0, out_mul=2.0 ** -15 * 2.0 ** -1, out_zp=0.0, levels=16,
                          exp_lut=softmax_exp_lut(torch.tensor([2.0 ** -3])), in_zp=0.0)
        lut = p.exp_lut.cuda()
        c = _cabi.Attention()
        c.score_mul, c.score_zp, c.out_mul, c.out_zp, c.softmax_levels = p.score_mul, p.score_zp, p.out_mul, p.out_zp, 16
        c.in_zp, c.exp_lut, c.lut_sig_bits = 0.0, lut.data_ptr(), p.lut_sig_bits
        time_both(qkv, b, n, heads, c, 'random +-%d, b=%d, n=%d' % (spread, b, n))
        return
    import diff_vit_b200 as dv
    model = dv.deit_small_patch16_224(pretrained=False, cfg=dv.Config(True, True, 'minmax')).eval().cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    dv.calibrate_model(model, [torch.randn(32, 3, 224, 224, device='cuda', generator=g)])
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(32, 3, 224, 224, device='cuda', generator=g)
    eng = model.integer_engine()
    bits = [8] * 50
    _, dump = eng.forward_dump(x, bits)
    bound = eng.bound(bits)
    for layer in (0, 5, 11):
        q = torch.from_numpy(dump['act/blocks.%d.attn.qact1' % layer]).reshape(32 * n, 3 * heads * 64)
        qkv = q.repeat((b + 31) // 32, 1)[:b * n].contiguous().cuda()
        sm = dump['softmax/blocks.%d.attn.log_int_softmax' % layer]
        att = bound.blocks[layer].attn
        time_both(qkv, b, n, heads, att, 'deit_small block %d, b=%d' % (layer, b))
        print('    score_mul 2^%d, softmax codes: min %d max %d mean %.2f' % (
            round(torch.log2(torch.tensor(att.score_mul)).item()), sm.min(), sm.max(), sm.mean()))


if __name__ == '__main__':
    main()
