"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL on B200, gloo on CPU tests).

The quantized forward is embarrassingly parallel over images, so inference shards the batch with no
data-path collective.  Collectives appear only where SURVEY.md section 8e puts them:
  * calibration: MAX / MIN all-reduce of the observers' range statistics and SUM all-reduce of the
    candidate-scale squared-error sums, so that every rank derives the same scales the single-process
    calibration of the concatenated batch would;
  * evaluation: all-gather of the logits (or of top-k hit counts).
"""
import contextlib

import torch
import torch.distributed as dist

_GROUP = None        # process group the observers reduce over while calibrating (None = single process)
_ACTIVE = False


def is_active():
    return _ACTIVE and dist.is_available() and dist.is_initialized() and dist.get_world_size(_GROUP) > 1


@contextlib.contextmanager
def calibration_group(group=None):
    """Within this context every observer all-reduces its statistics over `group`."""
    global _GROUP, _ACTIVE
    prev = (_GROUP, _ACTIVE)
    _GROUP, _ACTIVE = group, True
    try:
        yield
    finally:
        _GROUP, _ACTIVE = prev


def reduce_max_(t):
    if is_active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_GROUP)
    return t


def reduce_min_(t):
    if is_active():
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=_GROUP)
    return t


def reduce_sum_(t):
    if is_active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_GROUP)
    return t


def global_mean(sq_err, dims=None):
    """Mean of `sq_err` over `dims` (all dims when None) across every rank: SUM all-reduce of the local
    sums (accumulated in fp64 so that the arg-min over candidate scales agrees with the single-process
    run) divided by the global element count.  Single process: plain `.mean`, the reference's reduction."""
    if not is_active():
        return sq_err.mean() if dims is None else sq_err.mean(dim=dims)
    local = sq_err.double().sum() if dims is None else sq_err.double().sum(dim=dims)
    count = torch.tensor([sq_err.numel() if dims is None else sq_err.numel() // max(local.numel(), 1)],
                         dtype=torch.float64, device=sq_err.device)
    dist.all_reduce(local, op=dist.ReduceOp.SUM, group=_GROUP)
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=_GROUP)
    return (local / count).float()


def shard(x, rank=None, world=None):
    """Contiguous shard of a batch for this rank (256 / G images per GPU in BASELINE config 2).  With a batch that
    does not divide by the world size the trailing ranks get a shorter, possibly empty, shard."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (x.shape[0] + world - 1) // world
    return x[min(rank * per, x.shape[0]):min((rank + 1) * per, x.shape[0])]


def gather_logits(logits, group=None):
    """All-gather of per-rank logits [b_r, classes] -> [sum b_r, classes], rank order.  Shards may differ in length
    (ImageNet's last validation batch of 80 images on 8 ranks) or be empty: the lengths are gathered first, every rank
    contributes a buffer padded to the longest shard, and the padding is cut out again."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return logits
    world = dist.get_world_size(group)
    local = logits.contiguous()
    lengths = torch.zeros(world, dtype=torch.int64, device=local.device)
    lengths[dist.get_rank(group)] = local.shape[0]
    dist.all_reduce(lengths, op=dist.ReduceOp.SUM, group=group)
    lengths = [int(v) for v in lengths.tolist()]
    longest = max(lengths)
    if min(lengths) == longest:                      # the common case: one collective, no padding
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        return torch.cat(parts, dim=0)
    padded = local.new_zeros((longest,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, lengths)], dim=0)


def calibrate_model_distributed(model, local_batches, group=None):
    """`diff_vit_b200.calibrate_model` on this rank's shard of every calibration batch, with all range and
    error statistics reduced over `group`: all ranks end with identical quantization parameters."""
    from . import calibrate_model
    with calibration_group(group):
        return calibrate_model(model, local_batches)


def accuracy(output, target, topk=(1,)):
    """Precision@k in percent (reference: test_quant.py:488-501)."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand(maxk, -1))
    return [correct[:k].reshape(-1).float().sum(0).mul_(100.0 / target.size(0)) for k in topk]


def validate(model, batches, bit_config, group=None):
    """Sharded evaluation loop (reference: test_quant.py:411-466 without the ImageNet loader): each rank
    runs its shard, hit counts are summed over the group.  Returns (top1 %, top5 %, images)."""
    hits = torch.zeros(3, dtype=torch.float64)
    for data, target in batches:
        if data.shape[0] == 0:        # an empty shard of a short last batch: nothing to run, but stay in the collective
            continue
        with torch.no_grad():
            out, _, _ = model(data, bit_config, False)
        k = min(5, out.shape[1])
        p1, pk = accuracy(out.float().cpu(), target.cpu(), topk=(1, k))
        hits += torch.tensor([float(p1) * len(target) / 100.0, float(pk) * len(target) / 100.0, len(target)],
                             dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = next(model.parameters()).device if dist.get_backend(group) == 'nccl' else torch.device('cpu')
        h = hits.to(dev)
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        hits = h.cpu()
    n = max(float(hits[2]), 1.0)
    return 100.0 * float(hits[0]) / n, 100.0 * float(hits[1]) / n, int(hits[2])
