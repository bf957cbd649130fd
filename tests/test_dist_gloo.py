"""The N > 1 host logic on CPU: world_size-2 gloo processes calibrate on shards of the calibration batch
(statistics all-reduced) and must end with exactly the single-process quantization parameters; logits
gather and sharded top-k accounting are checked as well.  No GPU, no quantized arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, build_micro


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import diff_vit_b200 as dv
        from diff_vit_b200 import dist as dvd
        torch.set_num_threads(2)
        z = np.load(os.path.join(GOLDEN, 'micro_minmax.npz'))
        model = build_micro(z)
        x = torch.from_numpy(z['x_calib'])
        dvd.calibrate_model_distributed(model, [dvd.shard(x)])
        scales = {}
        for name, m in model.named_modules():
            if isinstance(m, dv.QAct) and m.quantizer.scale is not None:
                scales['scale/' + name] = m.quantizer.scale.numpy()
            if isinstance(m, (dv.QLinear, dv.QConv2d)):
                for bit, s in m.quantizer.dic_scale.items():
                    scales['wscale/%s/%s' % (name, bit)] = s.numpy().reshape(-1)
            if isinstance(m, (dv.Attention, dv.Mlp)):
                scales['cs/' + name] = m.channel_scale.numpy()
        bad = [k for k, v in scales.items() if not np.array_equal(v.reshape(-1), z[k].reshape(-1))]
        class Fake(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.zeros(1))

            def forward(self, data, bit_config, plot):
                return data, [], []
        # logits gather + sharded accuracy bookkeeping on fake logits
        g = torch.Generator().manual_seed(7)
        logits = torch.randn(8, 16, generator=g)
        target = torch.randint(0, 16, (8,), generator=g)
        mine = dvd.shard(logits)
        full = dvd.gather_logits(mine)
        ok_gather = torch.equal(full, logits)
        # shards of unequal length (5 rows on 2 ranks: 3 + 2) and an empty shard (1 row on 2 ranks: 1 + 0)
        for rows in (5, 1):
            part = dvd.shard(logits[:rows])
            ok_gather = ok_gather and torch.equal(dvd.gather_logits(part), logits[:rows])
        t1, _, n1 = dvd.validate(Fake(), [(dvd.shard(logits[:1]), dvd.shard(target[:1]))], [8])
        ok_gather = ok_gather and n1 == 1

        top1, top5, n = dvd.validate(Fake(), [(mine, dvd.shard(target))], [8])
        ref1, ref5 = dvd.accuracy(logits, target, topk=(1, 5))
        # percentile observer: the exact distributed order statistic equals torch.quantile on the whole batch
        from diff_vit_b200.ptq.observer import build_observer
        from diff_vit_b200.ptq.bit_type import BIT_TYPE_DICT
        act = torch.randn(6, 50, 128, generator=g) * 3
        whole = build_observer('percentile', 'activation', BIT_TYPE_DICT['int8'], 'layer_wise')
        whole.update(act)
        with dvd.calibration_group():
            part = build_observer('percentile', 'activation', BIT_TYPE_DICT['int8'], 'layer_wise')
            part.update(dvd.shard(act))
        ok_pct = (float(part.max_val) == float(whole.max_val) and float(part.min_val) == float(whole.min_val)
                  and float(whole.max_val) == float(torch.quantile(act.reshape(-1), 0.99999)))
        # the Swin family (BASELINE config 5) through the same distributed calibration: sharded batch, all-reduced
        # statistics, single-process (= the reference's) scales
        from test_swin_golden import build_swin_micro
        zs = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
        swin = build_swin_micro(zs)
        dvd.calibrate_model_distributed(swin, [dvd.shard(torch.from_numpy(zs['x_calib']))])
        bad_swin, n_swin = [], 0
        for name, m in swin.named_modules():
            if isinstance(m, dv.QAct) and m.quantizer.scale is not None and 'mlp.qact0' not in name:
                n_swin += 1
                if not np.array_equal(zs['scale/' + name].reshape(-1), m.quantizer.scale.numpy().reshape(-1)):
                    bad_swin.append(name)
            if isinstance(m, dv.Mlp) and not np.array_equal(zs['cs/' + name], m.best_scale[-1].numpy()):
                bad_swin.append(name + '/cs')
        ret[rank] = (len(scales), bad, ok_gather, abs(top1 - float(ref1)) < 1e-9 and abs(top5 - float(ref5)) < 1e-9, n,
                     ok_pct, n_swin, bad_swin)
    finally:
        dist.destroy_process_group()


def test_two_rank_calibration_matches_single_process_and_reference():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        count, bad, ok_gather, ok_acc, n, ok_pct, n_swin, bad_swin = ret[rank]
        assert n_swin >= 50 and bad_swin == [], 'rank %d: Swin scales differ from the reference calibration: %s' % (rank, bad_swin[:5])
        assert ok_pct, 'sharded percentile calibration differs from the single-process quantile'
        assert count == 71
        assert bad == [], 'rank %d: parameters differ from the single-process reference calibration: %s' % (rank, bad[:5])
        assert ok_gather and ok_acc and n == 8


def test_order_statistics_exact_on_cpu():
    """The radix select behind the percentile observer (torch path of the digit histogram)."""
    from diff_vit_b200.ptq.observer import gpu_stats
    g = torch.Generator().manual_seed(3)
    x = torch.randn(100003, generator=g) * 5
    x[:4] = torch.tensor([0.0, -0.0, 1e-30, -1e-30])
    ref = torch.sort(x).values
    ranks = [0, 1, 2, 50001, 100002, 100001]
    assert torch.equal(gpu_stats.order_statistics(x, ranks), ref[ranks])
    hi, lo = gpu_stats.quantile_pair(x, 0.99999, x.numel())
    assert float(hi) == float(torch.quantile(x, 0.99999)) and float(lo) == float(torch.quantile(x, 1 - 0.99999))
    with pytest.raises(IndexError):
        gpu_stats.order_statistics(x, [x.numel()])


def test_shard_covers_batch():
    from diff_vit_b200 import dist as dvd
    x = torch.arange(10).reshape(10, 1)
    parts = [dvd.shard(x, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), x) and [len(p) for p in parts] == [3, 3, 3, 1]
