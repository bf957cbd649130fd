"""Mixed-precision bit-width search over a calibrated model (the caller of the quantized forward in
test_quant.py:253-408): candidate generation under the model-size constraint, sensitivity ranking, and the
evolutionary refinement that evaluates hundreds of `bit_config`s.

One calibration serves every candidate: the integer engine keeps a plan per `bit_config` (re-selected from the
calibrated state, never re-calibrated), so a fitness evaluation is one quantized forward per validation batch.

Departures from the reference script, which cannot run as shipped (`mean_hessian` is undefined at :257/:313):
  * the per-layer sensitivities are an argument (uniform by default; the reference means Hessian traces);
  * `global_distance[i]` rows are indexed by bit-type NAME (int4 / int8), where the script's `k in {0, 1}`
    reads the uint3 / uint4 entries;
  * a child that violates the size constraint or repeats an earlier one is skipped, where the script appends it
    with the accuracy of whatever was evaluated before it.
"""
import random

from .ptq.bit_type import BIT_TYPE_LIST

BIT_CHOICE = (4, 8)
# position of a weight bit width inside a `global_distance` row (BIT_TYPE_LIST order without uint8)
_DISTANCE_INDEX = {b.bits: i for i, b in enumerate(b for b in BIT_TYPE_LIST if b.name != 'uint8') if b.signed}


def model_size(flops, bit_config):
    """sum_i FLOPs_i * bits_i, the size proxy of test_quant.py:262,281."""
    return sum(f * b for f, b in zip(flops, bit_config))


def size_constraint(flops, ratio=1.1, base_bits=4):
    return ratio * sum(f * base_bits for f in flops)


def candidate_configs(flops, rng=None, limit=50, ratio=1.1, max_draws=1 << 18):
    """Random candidates as test_quant.py:264-287 draws them: 8 bits for the patch embedding, one width per
    (qkv, proj) and per (fc1, fc2) pair, a free choice for the head; kept when within the size constraint.
    Plain rejection sampling like the script's: at the published ratio 1.1 only ~1e-4 of the draws qualify (at most
    three 8-bit pairs), so the draw count is bounded and fewer than `limit` candidates may come back."""
    rng = rng or random
    n = len(flops)
    bound = size_constraint(flops, ratio)
    out = []
    for _ in range(max_draws):
        pairs = [rng.choice(BIT_CHOICE) for _ in range(n // 2 - 1)]
        cfg = [max(BIT_CHOICE)] + [b for b in pairs for _ in range(2)] + [rng.choice(BIT_CHOICE)]
        if model_size(flops, cfg) <= bound and cfg not in out:
            out.append(cfg)
        if len(out) > limit:
            break
    return out


def omega(bit_config, global_distance, sensitivity=None):
    """Predicted loss increase of a configuration: sum_i sensitivity_i * distance_i[bits_i] over the linear layers
    (layer 0, the patch embedding, has no distance row); test_quant.py:291-316."""
    total = 0.0
    for i, bits in enumerate(bit_config[1:]):
        s = 1.0 if sensitivity is None else float(sensitivity[i])
        total += s * float(global_distance[i][_DISTANCE_INDEX[bits]])
    return total


def rank_by_omega(configs, global_distance, sensitivity=None):
    return sorted(([cfg, omega(cfg, global_distance, sensitivity)] for cfg in configs), key=lambda e: e[1])


def evolutionary_search(initial, fitness, flops, rng=None, pop_size=25, iterations=8, mutate_size=10,
                        mutate_prob=0.5, crossover_size=10, crossover_prob=0.5, ratio=1.1, log=None):
    """test_quant.py:340-402: keep the `pop_size` fittest configurations; each iteration adds `mutate_size` mutated
    and `crossover_size` crossed-over children that satisfy the size constraint.  `fitness(bit_config) -> float`
    (higher is better, e.g. top-1 of `dist.validate`); results are memoised, so a configuration is evaluated once."""
    rng = rng or random
    bound = size_constraint(flops, ratio)
    seen = {}

    def score(cfg):
        key = tuple(cfg)
        if key not in seen:
            seen[key] = float(fitness(list(cfg)))
        return seen[key]

    parents = sorted(([list(c), score(c)] for c in initial[:pop_size]), key=lambda e: e[1], reverse=True)
    for it in range(iterations):
        children = []
        tried = 0
        while len(children) < mutate_size and tried < 100 * mutate_size:
            tried += 1
            old = rng.choice(parents)[0]
            new = [b if rng.random() < mutate_prob else rng.choice(BIT_CHOICE) for b in old]
            if model_size(flops, new) <= bound and tuple(new) not in seen:
                children.append([new, score(new)])
        made = tried = 0
        while made < crossover_size and tried < 100 * crossover_size and len(parents) > 1:
            tried += 1
            a, b = rng.choice(parents)[0], rng.choice(parents)[0]
            if a == b:
                continue
            new = [x if rng.random() < crossover_prob else y for x, y in zip(a, b)]
            if model_size(flops, new) <= bound and tuple(new) not in seen:
                children.append([new, score(new)])
                made += 1
        for child in children:
            if child[1] > parents[-1][1] or len(parents) < pop_size:
                parents.append(child)
        parents = sorted(parents, key=lambda e: e[1], reverse=True)[:pop_size]
        if log is not None:
            log('evolution %d: best %.3f, %d configurations evaluated' % (it, parents[0][1], len(seen)))
    return parents, seen


def search(model, val_batches, global_distance, sensitivity=None, rng=None, top=25, group=None, **evo):
    """The whole driver on a calibrated, quantized model: candidates -> omega ranking -> evolutionary search with
    top-1 accuracy on `val_batches` ([(images, labels), ...]) as fitness.  Returns (population, evaluated)."""
    from . import dist as _dist
    flops = model.flops()
    ranked = rank_by_omega(candidate_configs(flops, rng), global_distance, sensitivity)

    def fitness(cfg):
        return _dist.validate(model, val_batches, cfg, group)[0]
    return evolutionary_search([c for c, _ in ranked[:top]], fitness, flops, rng, **evo)
