#!/usr/bin/env python
"""Post-training quantization entry point with the command line of the reference's test_quant.py
(flags at test_quant.py:20-53): build -> calibrate -> (mixed-precision search | single bit_config) -> validate,
with the quantized forward running in the sm_100a integer engine.

Differences forced by this environment, all explicit:
  * no network: `--checkpoint FILE` (a state_dict with the reference's key names) replaces `pretrained=True`;
    without it the weights are the factories' random init;
  * `--data synthetic[:N]` (default) evaluates on N seeded Gaussian batches whose labels are the model's own
    all-8-bit predictions, so accuracy reads as agreement with W8A8; an ImageNet folder is used when given
    (needs torchvision);
  * `--mode 2` (PSAQ generated calibration data) is out of scope; modes 0 / 1 calibrate on real / Gaussian data.
One process per GPU under torchrun: the calibration batch and the validation set are sharded, observer statistics
and hit counts are reduced over NCCL.
"""
import argparse
import os
import random

import numpy as np
import torch

import diff_vit_b200 as dv
from diff_vit_b200 import dist as dvd
from diff_vit_b200 import search


def arguments():
    ap = argparse.ArgumentParser(description='P2-ViT post-training quantization on the B200 integer engine')
    ap.add_argument('--model', default='deit_tiny', choices=sorted(dv._MODELS))
    ap.add_argument('--data', default='synthetic', help="ImageNet root, or 'synthetic[:batches]'")
    ap.add_argument('--checkpoint', default=None, help='state_dict file with the reference key names')
    ap.add_argument('--quant', default=False, action='store_true')
    ap.add_argument('--ptf', default=True, type=lambda v: str(v).lower() not in ('0', 'false'))
    ap.add_argument('--lis', default=True, type=lambda v: str(v).lower() not in ('0', 'false'))
    ap.add_argument('--quant-method', default='minmax', choices=['minmax', 'ema', 'omse', 'percentile'])
    ap.add_argument('--mixed', default=False, action='store_true', help='run the bit-width search (test_quant.py:253-408)')
    ap.add_argument('--bits', default='8', help="bit_config without --mixed: '4', '8' or a comma-separated list")
    ap.add_argument('--calib-batchsize', default=32, type=int)
    ap.add_argument('--mode', default=1, type=int, help='0 real-data calibration, 1 Gaussian noise')
    ap.add_argument('--calib-iter', default=1, type=int)
    ap.add_argument('--val-batchsize', default=256, type=int)
    ap.add_argument('--num-workers', default=8, type=int)
    ap.add_argument('--device', default='cuda', type=str)
    ap.add_argument('--print-freq', default=10, type=int)
    ap.add_argument('--seed', default=0, type=int)
    return ap.parse_args()


def set_seed(value):
    random.seed(value)
    np.random.seed(value)
    torch.manual_seed(value)
    torch.cuda.manual_seed_all(value)


# per-family preprocessing of the reference (test_quant.py:99-113): mean, std, crop fraction
_PREPROCESS = {'deit': ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225), 0.875),
               'vit': ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5), 0.9)}


def imagenet_loader(root, split, family, batch, workers):
    from torchvision import datasets, transforms
    mean, std, crop = _PREPROCESS[family]
    tf = transforms.Compose([transforms.Resize(int(224 / crop), interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(224), transforms.ToTensor(), transforms.Normalize(mean, std)])
    return torch.utils.data.DataLoader(datasets.ImageFolder(os.path.join(root, split), tf), batch_size=batch,
                                       shuffle=False, num_workers=workers, pin_memory=True)


def main():
    args = arguments()
    set_seed(args.seed)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    rank = int(os.environ.get('RANK', '0'))
    device = torch.device(args.device, local) if args.device == 'cuda' else torch.device(args.device)
    if device.type == 'cuda':
        torch.cuda.set_device(device)
    if world > 1:
        torch.distributed.init_process_group('nccl' if device.type == 'cuda' else 'gloo')
    say = print if rank == 0 else (lambda *a, **k: None)

    model = dv.str2model(args.model)(pretrained=False, cfg=dv.Config(args.ptf, args.lis, args.quant_method))
    if args.checkpoint:
        model.load_state_dict(torch.load(args.checkpoint, map_location='cpu'))
    model = model.to(device).eval()
    family = args.model.split('_')[0]
    synthetic = args.data.startswith('synthetic')

    # ---- calibration -------------------------------------------------------------------------------------------
    if args.quant:
        gen = torch.Generator(device=device).manual_seed(args.seed)
        if args.mode == 1 or synthetic:
            say('Calibrating with Gaussian noise...')
            batches = [torch.randn(args.calib_batchsize, 3, 224, 224, device=device, generator=gen)
                       for _ in range(max(1, args.calib_iter))]
        elif args.mode == 0:
            say('Calibrating with real data...')
            loader = imagenet_loader(args.data, 'train', family, args.calib_batchsize, args.num_workers)
            batches = [x.to(device) for i, (x, _) in zip(range(args.calib_iter), loader)]
        else:
            raise SystemExit('--mode %d (generated calibration data) is not part of this package' % args.mode)
        if world > 1:
            dvd.calibrate_model_distributed(model, [dvd.shard(b) for b in batches])
        else:
            dv.calibrate_model(model, batches)

    # ---- validation set ----------------------------------------------------------------------------------------
    n_layers = len(model.flops())
    if synthetic:
        count = int(args.data.split(':')[1]) if ':' in args.data else 2
        gen = torch.Generator(device=device).manual_seed(args.seed + 1 + rank)
        val = []
        for _ in range(count):
            x = torch.randn(args.val_batchsize // world, 3, 224, 224, device=device, generator=gen)
            with torch.no_grad():
                ref = model(x, [8] * n_layers, False)[0] if args.quant else model(x)[0]
            # quantized logits tie often; accuracy() ranks with topk(5), whose first column can differ from topk(1)'s
            # pick among tied maxima, so the labels are taken from the very same call
            val.append((x, ref.float().cpu().topk(min(5, ref.shape[1]), 1, True, True)[1][:, 0].contiguous()))
    else:
        loader = imagenet_loader(args.data, 'val', family, args.val_batchsize, args.num_workers)
        val = ((dvd.shard(x).to(device), dvd.shard(y)) for x, y in loader) if world > 1 else \
              ((x.to(device), y) for x, y in loader)
        val = list(val) if args.mixed else val

    def evaluate(bit_config):
        top1, top5, images = dvd.validate(model, val, bit_config)
        say(' * Prec@1 %.3f Prec@5 %.3f  (%d images)  %s' % (top1, top5, images, bit_config))
        return top1

    if not args.quant:
        with torch.no_grad():
            hits = sum(int((model(x)[0].argmax(1).cpu() == y.cpu()).sum()) for x, y in val)
        say(' * float model: %d hits' % hits)
        return
    if args.mixed:
        say('Pareto frontier + evolutionary search...')
        population, seen = search.search(model, val, model.global_distance, rng=random.Random(args.seed), log=say)
        say('best configurations (%d evaluated):' % len(seen))
        for cfg, acc in population[:5]:
            say('  %.3f  %s' % (acc, cfg))
    else:
        bits = [int(b) for b in args.bits.split(',')]
        evaluate(bits * n_layers if len(bits) == 1 else bits)


if __name__ == '__main__':
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
