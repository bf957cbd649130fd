"""The N > 1 path on real GPUs: two ranks, one process per GPU, NCCL (BASELINE config 2's partition).
Skipped on a single-GPU box; tests/test_dist_gloo.py covers the same host logic on CPU with gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, build_micro

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import diff_vit_b200 as dv
        from diff_vit_b200 import dist as dvd
        z = np.load(os.path.join(GOLDEN, 'micro_minmax.npz'))
        # calibration on shards, statistics all-reduced over NCCL: the reference's single-process scales
        model = build_micro(z).to(dev)
        x = torch.from_numpy(z['x_calib']).to(dev)
        dvd.calibrate_model_distributed(model, [dvd.shard(x)])
        bad = []
        for name, m in model.named_modules():
            if isinstance(m, dv.QAct) and m.quantizer.scale is not None:
                want, got = z['scale/' + name].reshape(-1), m.quantizer.scale.cpu().numpy().reshape(-1)
                ok = np.array_equal(want, got) if want.size == 1 else (
                    np.array_equal(want / want.min(), got / got.min()) and abs(want.min() / got.min() - 1) < 1e-4)
                if not ok:
                    bad.append(name)
            if isinstance(m, (dv.Attention, dv.Mlp)) and not np.array_equal(z['cs/' + name], m.channel_scale.cpu().numpy()):
                bad.append(name + '/cs')
        # sharded quantized forward + NCCL gather of the logits == the unsharded forward (and the reference's golden)
        xe = torch.from_numpy(z['x_eval']).to(dev)
        bc = [8] * 10
        full, _, _ = model(xe, bc, False)
        part, _, _ = model(dvd.shard(xe).contiguous(), bc, False)
        gathered = dvd.gather_logits(part)
        same = bool(torch.equal(gathered, full))
        lsb = float(model.act_out.quantizer.scale)
        golden = float(np.abs(gathered.cpu().numpy() - z['w8/logits']).max()) <= lsb
        # unequal shards (5 images on 2 ranks) through the padded gather
        odd = xe[:5]
        p5, _, _ = model(dvd.shard(odd).contiguous(), bc, False)
        same = same and bool(torch.equal(dvd.gather_logits(p5), full[:5]))
        top1, top5, n = dvd.validate(model, [(dvd.shard(xe).contiguous(), dvd.shard(torch.from_numpy(z['w8/logits']).argmax(1)))], bc)
        # the Swin family on the same plumbing: NCCL-reduced calibration == the reference's scales, sharded forward on
        # the Swin integer engine + gather == the unsharded forward == the reference's golden logits
        from test_swin_golden import build_swin_micro
        zs = np.load(os.path.join(GOLDEN, 'swin_micro.npz'))
        swin = build_swin_micro(zs).to(dev)
        dvd.calibrate_model_distributed(swin, [dvd.shard(torch.from_numpy(zs['x_calib']).to(dev))])
        bad_swin = [name for name, m in swin.named_modules()
                    if isinstance(m, dv.QAct) and m.quantizer.scale is not None and 'mlp.qact0' not in name
                    and m.quantizer.scale.numel() == 1
                    and not np.array_equal(zs['scale/' + name].reshape(-1), m.quantizer.scale.cpu().numpy().reshape(-1))]
        xs = torch.from_numpy(zs['x_eval']).to(dev)
        with torch.no_grad():
            sfull = swin(xs)
            spart = swin(dvd.shard(xs).contiguous())
        swin_ok = (swin._engine_off is None and bool(torch.equal(dvd.gather_logits(spart), sfull))
                   and float(np.abs(sfull.cpu().numpy() - zs['w8/logits']).max()) <= 2 * float(swin.act_out.quantizer.scale))
        ret[rank] = (bad, same, golden, top1, n, bad_swin, swin_ok)
    finally:
        dist.destroy_process_group()


def test_two_gpus_nccl_calibration_and_logits_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        bad, same, golden, top1, n, bad_swin, swin_ok = ret[rank]
        assert bad == [], bad
        assert bad_swin == [] and swin_ok, bad_swin
        assert same and golden
        assert n == 6 and top1 >= 99.0       # all six eval images, labels = the reference's own argmax
