"""Mixed-precision search driver (host logic; the fitness here is synthetic, the GPU run is in test_gpu_model)."""
import random

import numpy as np
import torch

from diff_vit_b200 import search


def _flops(depth=12, d=384, n=197, hid=1536, classes=1000):
    out = [3 * 16 * 16 * d * 196]
    for _ in range(depth):
        out += [n * d * 3 * d, n * d * d, n * d * hid, n * hid * d]
    return out + [d * classes]


def test_candidates_respect_constraint_and_layout():
    flops = _flops()
    rng = random.Random(0)
    cands = search.candidate_configs(flops, rng, limit=50, ratio=1.5)
    assert len(cands) == 51 and len({tuple(c) for c in cands}) == 51
    bound = search.size_constraint(flops, 1.5)
    for c in cands:
        assert len(c) == 50 and c[0] == 8 and set(c) <= {4, 8}
        assert search.model_size(flops, c) <= bound
        assert all(c[1 + 2 * i] == c[2 + 2 * i] for i in range(24))      # (qkv, proj) and (fc1, fc2) share a width


def test_omega_reads_the_named_distance_columns():
    rows = [[100.0, 50.0, 3.0, 1.0]] * 49          # uint3, uint4, int4, int8 distances per layer
    assert search.omega([8] + [4] * 49, rows) == 3.0 * 49
    assert search.omega([8] + [8] * 49, rows) == 1.0 * 49
    sens = np.arange(49, dtype=np.float64)
    assert search.omega([8] + [4] * 49, rows, sens) == 3.0 * sens.sum()
    ranked = search.rank_by_omega([[8] + [4] * 49, [8] * 50], rows)
    assert ranked[0][0] == [8] * 50 and ranked[0][1] < ranked[1][1]


def test_evolution_improves_a_synthetic_fitness_and_memoises():
    flops = _flops()
    rng = random.Random(1)
    target = [8, 8, 8] + [4] * 46 + [8]             # the "sensitive" layers: fitness counts matches

    calls = []

    def fitness(cfg):
        calls.append(tuple(cfg))
        return sum(1.0 for a, b in zip(cfg, target) if a == b)
    init = search.candidate_configs(flops, rng, limit=30, ratio=1.5)
    start = max(fitness(c) for c in init)
    calls.clear()
    pop, seen = search.evolutionary_search(init, fitness, flops, rng, pop_size=10, iterations=6, ratio=1.5)
    assert len(calls) == len(set(calls)) == len(seen)                    # every configuration evaluated once
    assert len(pop) == 10 and pop[0][1] >= start and pop[0][1] >= pop[-1][1]
    bound = search.size_constraint(flops, 1.5)
    assert all(search.model_size(flops, c) <= bound for c, _ in pop)
    # the published constraint (1.1 x the all-4-bit size) leaves room for at most three 8-bit pairs
    tight = search.candidate_configs(flops, random.Random(2), limit=3, max_draws=1 << 16)
    assert all(search.model_size(flops, c) <= search.size_constraint(flops) and c.count(8) <= 8 for c in tight)


def test_calibration_keeps_the_distance_rows(micro_model):
    gd = micro_model.global_distance
    assert len(gd) == len(micro_model.flops()) - 1 and all(len(row) == 4 for row in gd)
    cfg = [8] * len(micro_model.flops())
    assert search.omega(cfg, gd) >= 0.0 and torch.isfinite(torch.tensor(search.omega(cfg, gd)))


def test_test_quant_entry_point_float_path_on_cpu():
    """The command-line entry point parses the reference's flags and runs the float model on CPU (the quantized
    branch needs the GPU engine and is exercised by the GPU suite / the bench box)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'test_quant.py'), '--model', 'deit_tiny', '--device', 'cpu',
                          '--val-batchsize', '2', '--data', 'synthetic:1', '--ptf', 'True', '--lis', 'True',
                          '--quant-method', 'minmax', '--calib-batchsize', '4', '--seed', '1'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'float model: 2 hits' in out.stdout
    bad = subprocess.run([sys.executable, os.path.join(ROOT, 'test_quant.py'), '--quant', '--device', 'cpu',
                          '--val-batchsize', '2', '--calib-batchsize', '2', '--data', 'synthetic:1'],
                         capture_output=True, text=True, timeout=900)
    assert bad.returncode != 0 and 'no CPU fallback' in (bad.stderr + bad.stdout)
