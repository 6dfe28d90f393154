"""GPU tests of the gating battles (reference ai.py:1975-2114): per-side networks AND per-side search settings,
colour bookkeeping, sharding over ranks."""
import copy

import numpy as np
import pytest

from test_gpu_search import fake_evaluator_torch

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from tetris_reinforcement_learning_b200.config import Config
    base = dict(visual=False, ruleset="s2", model="pytorch", MAX_ITER=6, training=False)
    base.update(kw)
    return Config(**base)


def _shifted(ev, shift):
    """A second deterministic evaluator: the first one with rolled logits and mirrored values."""
    def ev2(grids, extras):
        v, l = ev(grids, extras)
        return (1.0 - v).contiguous(), l.roll(shift, dims=1).contiguous()
    return ev2


def test_each_side_searches_with_its_own_settings():
    """Network 1 plays player (game_id & 1) (ai.py:2087-2091); a search runs with the config of the side to move at
    the root (ai.py:2012-2016): with MAX_ITER 7 vs 3 the iteration count of every search tells who ran it."""
    import ctypes
    import torch
    from tetris_reinforcement_learning_b200 import _native
    from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, search_params_from_config
    dev = torch.device("cuda:0")
    c1, c2 = _cfg(MAX_ITER=7), _cfg(MAX_ITER=3, CPUCT=1.5)
    eng = SelfPlayEngine(c1, fake_evaluator_torch(dev), 32, seed=9, save_all=True, restart_finished=False, max_rounds=3,
                         use_cuda_graph=False, sample_cap=8192)
    pair = (_native.SearchParams * 2)(search_params_from_config(c1, 9, False), search_params_from_config(c2, 9, False))
    for p in pair:
        p.save_all = 1
        p.max_rounds = 3
    dev_pair = torch.frombuffer(bytearray(bytes(pair)), dtype=torch.uint8).to(dev)
    eng.buf.params2 = dev_pair.data_ptr()
    eng.step(60)
    samples, ends = eng.drain()
    assert len(samples) > 100 and len(ends) == 32
    owner = (samples["game_id"] ^ samples["turn"]) & 1
    assert (samples["iterations"][owner == 0] == 7).all() and (samples["iterations"][owner == 1] == 3).all()
    assert (owner == 0).any() and (owner == 1).any()
    assert eng.status_bits() == 0


def test_battle_bookkeeping_and_sharding_invariance():
    """battle_networks with two deterministic evaluators: wins sum to the number of games; the games of a 2-way shard
    (ids 0,2,4,.. and 1,3,5,..) are the same games as the unsharded run, so the win counts add up exactly."""
    import torch
    from tetris_reinforcement_learning_b200 import training
    dev = torch.device("cuda:0")
    ev1 = fake_evaluator_torch(dev)
    ev2 = _shifted(ev1, 17)
    c1, c2 = _cfg(MAX_ITER=6), _cfg(MAX_ITER=4)
    full, acc = training.battle_networks(ev1, c1, ev2, c2, 0.5, "more", 24, seed=5)
    assert full.sum() == 24 and acc in (True, False, None)
    parts = [training.battle_networks(ev1, c1, ev2, c2, None, "more", 12, seed=5, first_game_id=r, game_id_stride=2)[0]
             for r in range(2)]
    assert np.array_equal(parts[0] + parts[1], full)
    # swapping the networks swaps the colours: same games seen from the other side
    swapped, _ = training.battle_networks(ev2, c2, ev1, c1, None, "more", 24, seed=5)
    assert swapped.sum() == 24


@pytest.mark.parametrize("family", ["alphasame16", "aux32", "mixed"])
def test_battle_with_cached_trunk_evaluators(family):
    """Two real networks: every leaf goes through one network's cached-trunk path (DualCachedEvaluator).  A network
    against an exact copy of itself must score like a single-network engine would: all games finish, wins sum up."""
    import torch
    from tetris_reinforcement_learning_b200 import architectures as arch, training
    torch.manual_seed(0)
    a16 = arch.AlphaSame(arch.AlphaSameConfig(blocks=2, filters=16)).cuda().eval()
    x32 = arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=1, filters=32)).cuda().eval()
    n1, n2 = {"alphasame16": (a16, copy.deepcopy(a16)), "aux32": (x32, copy.deepcopy(x32)), "mixed": (a16, x32)}[family]
    cfg = _cfg(MAX_ITER=6)
    wins, _ = training.battle_networks(n1, cfg, n2, cfg, None, "more", 16, seed=3)
    assert wins.sum() == 16
