#!/usr/bin/env python
"""Pin the search oracle and the feature oracle against the UNMODIFIED reference
(build container only):

    python oracle/pin_mcts_against_reference.py [--searches N] [--iters K]

  1. features: oracle.features_oracle.encode == ai.game_to_X on random mid-game states;
  2. search: oracle.mcts_oracle.search == ai.MCTS driven by the same fake evaluator and the same
     Philox tape, over four config families (eval; training defaults with playout-cap + Gamma
     noise + temperature; forced playouts + pruning; absolute FPU / tanh / no root softmax):
     chosen move, post-prune visit counts, priors, save flag, node count, root value.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import features_oracle, mcts_oracle, oracle, refharness as rh  # noqa: E402
from oracle.pin_against_reference import pick_move, random_midgame  # noqa: E402


def config_families(ai, iters):
    base = dict(visual=False, ruleset="s2", model="pytorch", MAX_ITER=iters, CPUCT=0.75)
    return [
        ("eval", ai.Config(training=False, **base)),
        ("training-defaults", ai.Config(training=True, **base)),
        ("forced+pruning", ai.Config(training=True, use_forced_playouts_and_policy_target_pruning=True,
                                     use_playout_cap_randomization=False, **base)),
        ("absolute-tanh-noroot", ai.Config(training=True, FpuStrategy="absolute", use_tanh=True,
                                           use_root_softmax=False, **base)),
    ]


def tanh_wrap(evaluate):
    """tanh nets output values in (-1, 1): stretch the fake value accordingly."""
    def f(rec):
        v, p = evaluate(rec)
        return 2 * v - 1, p
    return f


def positions(n, seed):
    """Mid-game positions reached by biased-random legal play from mixed starts."""
    rng = np.random.default_rng(seed)
    games = random_midgame(rng, n, seed)
    out = []
    for i in range(n):
        rec = games[i:i + 1].copy()
        for _ in range(int(rng.integers(0, 6))):
            if mcts_oracle.is_terminal(rec[0]) or mcts_oracle.no_move(rec[0]):
                break
            legal = mcts_oracle.legal_moves(rec[0])
            if legal.size == 0:
                break
            oracle.env_step(rec, np.array([pick_move(rng, rec, legal, seed)], np.uint16), True, seed)
        if mcts_oracle.is_terminal(rec[0]) or mcts_oracle.no_move(rec[0]) or mcts_oracle.legal_moves(rec[0]).size == 0:
            continue
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--searches", type=int, default=24)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--seed", type=int, default=20261018)
    args = ap.parse_args()
    if not rh.available():
        print("reference checkout not available")
        return 2
    m = rh.full_modules()
    t0 = time.time()
    bad = 0
    pos = positions(max(args.searches, 40), args.seed)

    nfeat = 0
    for rec in pos:
        g1, e1 = features_oracle.encode(rec[0])
        g2, e2 = rh.reference_game_to_X(rec[0])
        nfeat += 1
        if not (np.array_equal(g1, g2) and np.array_equal(e1, e2)):
            bad += 1
            print("  FEATURE MISMATCH")
    print(f"features: {nfeat} states, mismatches {bad}")

    fams = config_families(m.ai, args.iters)
    n_search = 0
    for k in range(args.searches):
        rec = pos[k % len(pos)]
        name, cfg = fams[k % len(fams)]
        ev = tanh_wrap(features_oracle.fake_evaluate) if cfg.use_tanh else features_oracle.fake_evaluate
        gid = int(rec[0]["game_id"])  # the garbage draws are keyed by the game's own id
        want = rh.reference_mcts(cfg, rec[0], ev, mcts_oracle.SearchTape(args.seed, gid, k))
        got = mcts_oracle.search(cfg, rec, ev, mcts_oracle.SearchTape(args.seed, gid, k))
        n_search += 1
        same = (want["move"] == got["move"] and want["moves"] == got["moves"] and
                want["visits_post"] == got["visits_post"] and want["save"] == got["save"] and
                want["n_nodes"] == got["n_nodes"] and want["root_visits"] == got["root_visits"] and
                want["priors"] == got["priors"] and want["root_value_avg"] == got["root_value_avg"] and
                want["garbage_draws"] == got["garbage_draws"])
        if not same:
            bad += 1
            print(f"  SEARCH MISMATCH #{k} [{name}] move {want['move']} vs {got['move']}, nodes {want['n_nodes']} vs "
                  f"{got['n_nodes']}, visits equal {want['visits_post'] == got['visits_post']}, priors equal "
                  f"{want['priors'] == got['priors']}")
    print(f"search: {n_search} searches x {args.iters} iterations (4 config families), mismatches {bad}")
    print(f"total mismatches {bad} ({time.time() - t0:.1f} s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
