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
        # logits gather + sharded accuracy bookkeeping on fake logits
        g = torch.Generator().manual_seed(7)
        logits = torch.randn(8, 16, generator=g)
        target = torch.randint(0, 16, (8,), generator=g)
        mine = dvd.shard(logits)
        full = dvd.gather_logits(mine)
        ok_gather = torch.equal(full, logits)

        class Fake(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.zeros(1))

            def forward(self, data, bit_config, plot):
                return data, [], []
        top1, top5, n = dvd.validate(Fake(), [(mine, dvd.shard(target))], [8])
        ref1, ref5 = dvd.accuracy(logits, target, topk=(1, 5))
        ret[rank] = (len(scales), bad, ok_gather, abs(top1 - float(ref1)) < 1e-9 and abs(top5 - float(ref5)) < 1e-9, n)
    finally:
        dist.destroy_process_group()


def test_two_rank_calibration_matches_single_process_and_reference():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        count, bad, ok_gather, ok_acc, n = ret[rank]
        assert count == 71
        assert bad == [], 'rank %d: parameters differ from the single-process reference calibration: %s' % (rank, bad[:5])
        assert ok_gather and ok_acc and n == 8


def test_shard_covers_batch():
    from diff_vit_b200 import dist as dvd
    x = torch.arange(10).reshape(10, 1)
    parts = [dvd.shard(x, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), x) and [len(p) for p in parts] == [3, 3, 3, 1]
