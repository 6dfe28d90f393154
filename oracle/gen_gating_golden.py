"""Golden vectors of the gating bookkeeping, produced by the UNMODIFIED reference (build container only):
`_check_threshold` (ai.py:2055-2069) on a grid, and the win tally of `_battle_networks_async` (ai.py:2087-2114) with
the games replaced by predetermined (winner, side) outcomes.  -> tests/golden/gating_golden.json

    python oracle/gen_gating_golden.py
"""
import asyncio
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness  # noqa: E402


def main():
    out_path = os.path.join(ROOT, "tests", "golden", "gating_golden.json")
    m = refharness.full_modules()
    ai = m.ai
    thresholds = []
    for games in (1, 2, 7, 10, 40, 200):
        for thr in (None, 0.5, 0.52, 0.55, 0.6):
            for ttype in ("more", "moreorequal"):
                for w0 in np.linspace(0, games, 9):
                    w0 = float(np.round(w0 * 2) / 2)
                    for w1 in (games - w0, max(0.0, games - w0 - 1.0)):
                        thresholds.append([w0, float(w1), games, thr, ttype, ai._check_threshold([w0, w1], games, thr, ttype)])

    class FakeEvaluator:
        def __init__(self, *a, **k):
            pass

        async def start(self):
            pass

        async def stop(self):
            pass

    rng = np.random.default_rng(7)
    tallies = []
    real_game, real_ev = ai.aplay_battle_game, ai.BatchedEvaluator
    try:
        ai.BatchedEvaluator = FakeEvaluator
        for games in (1, 2, 5, 16, 33):
            winners = [int(x) for x in rng.integers(-1, 2, size=games)]

            async def fake_game(c1, c2, ev1, ev2, side, _w=winners, _k=[0]):
                w = _w[_k[0]]
                _k[0] += 1
                return w, side

            ai.aplay_battle_game = fake_game
            cfg = type("C", (), {"ruleset": "s2"})()
            wins = asyncio.run(ai._battle_networks_async(None, cfg, None, cfg, games))
            # the reference gives game i the side i % 2 (ai.py:2091); games are created in order
            tallies.append({"results": [[w, i % 2] for i, w in enumerate(winners)], "wins": [float(wins[0]), float(wins[1])]})
    finally:
        ai.aplay_battle_game, ai.BatchedEvaluator = real_game, real_ev
    with open(out_path, "w") as f:
        json.dump({"check_threshold": thresholds, "tally": tallies,
                   "source": "reference ai._check_threshold / ai._battle_networks_async (oracle/gen_gating_golden.py)"}, f)
    print(len(thresholds), "threshold cases,", len(tallies), "tallies ->", out_path)


if __name__ == "__main__":
    main()
