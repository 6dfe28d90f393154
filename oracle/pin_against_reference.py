#!/usr/bin/env python
"""Pin the CPU oracle against the UNMODIFIED reference (build container only).

    python oracle/pin_against_reference.py [--movegen N] [--games N] [--seed S]

Runs three differential sweeps and exits non-zero on the first class of mismatch:
  1. movegen: oracle masks == get_move_matrix(player, 'convolutional') on the reference's own
     fixture boards (util.py:34-37, known answers of SURVEY A.8) and on N random calls
     (4 board families incl. adversarial caves, alt piece presented as hold or as queue head);
  2. attack: Stats.get_attack (s2) on the full grid of (rows, tspin, mini, all_clear, combo,
     b2b, level) states;
  3. env step: random legal moves played through Game.make_move with tape RNG, full state
     compared after every move (add_bag True and False, garbage exchange, top-outs).
The results of the last run are recorded in DESIGN.md.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oracle, refharness as rh  # noqa: E402
from tetris_reinforcement_learning_b200 import synth  # noqa: E402
from tetris_reinforcement_learning_b200.state import (GAME_DTYPE, game_state_dict,  # noqa: E402
                                                      reference_game_state_dict)


def sweep_movegen(n_calls, seed):
    bad = 0
    fb = rh.fixture_boards()
    for name, rows in fb.items():
        for i in range(7):
            ref = rh.movegen_packed(rows, i, i)
            mine = oracle.movegen_one(rows, i, i)[0]
            bad += int(not (ref == mine).all())
    n_boards = max(1, n_calls // 7)
    boards, cur, alt = synth.movegen_workload(n_boards, seed=seed, caves=True)
    placements = 0
    max_push = max_emit = 0
    for j in range(boards.shape[0]):
        c, a = int(cur[j]), int(alt[j])
        mode = j % 4
        if mode == 3:
            a = c  # held == current
        if mode == 2 and j % 8 == 2:
            c = 255  # no active piece, hold only
        ref = rh.movegen_packed(boards[j], c, a, alt_is_held=(mode != 1) or c == 255)
        mine, st, npush, nemit = oracle.movegen_one(boards[j], c, a)
        placements += int(ref.sum())
        max_push, max_emit = max(max_push, npush), max(max_emit, nemit)
        if not (ref == mine).all() or st != 0:
            bad += 1
            if bad < 5:
                print("  MOVEGEN MISMATCH call", j, "cur", c, "alt", a, int(ref.sum()), int(mine.sum()), st)
    print(f"movegen: {boards.shape[0] + 35} calls, {placements} placements, max pushes {max_push}, "
          f"max emissions {max_emit}, mismatches {bad}")
    return bad


def sweep_attack(ruleset="s2"):
    m = rh.modules()
    bad = n = 0
    mine = oracle.get_attack_s1 if ruleset == "s1" else oracle.get_attack_s2
    for rows in range(0, 5):
        for tspin in (False, True):
            for mini in (False, True):
                for pc in (False, True):
                    for combo in range(0, 21):
                        for b2b in list(range(-1, 12)) + [23, 24, 66, 67, 184, 185, 503, 504, 1369, 1370]:
                            for lvl in range(0, 9):
                                s = m.stats.Stats(ruleset)
                                s.combo, s.b2b, s.b2b_level = combo, b2b, lvl
                                a = s.get_attack(rows, tspin, mini, pc, "T")
                                got = mine(rows, tspin, mini, pc, combo, b2b, lvl)
                                n += 1
                                if got != (a, s.combo, s.b2b, s.b2b_level):
                                    bad += 1
                                    if bad < 5:
                                        print("  ATTACK MISMATCH", rows, tspin, mini, pc, combo, b2b, lvl,
                                              (a, s.combo, s.b2b, s.b2b_level), got)
    print(f"attack ({ruleset}): {n} cases, mismatches {bad}")
    return bad


def random_midgame(rng, n, seed, ruleset="s2"):
    """Packed games with synthetic boards, pending garbage and non-trivial stats."""
    games = oracle.game_setup(n, first_game_id=1000, seed=seed, ruleset=ruleset)
    boards = synth.random_boards(2 * n, seed=int(rng.integers(1 << 30)), caves=True)
    for i in range(n):
        for pl in range(2):
            p = games[i]["players"][pl]
            kind = rng.random()
            if kind < 0.35:      # messy synthetic stack
                rows = boards[2 * i + pl].copy()
                rows[:22] = 0    # keep the spawn area free so the game is not over at once
                p["rows"] = rows
            elif kind < 0.7:     # nearly-full rows: line clears, quads, spins
                h = int(rng.integers(1, 12))
                rows = np.zeros(40, dtype=np.uint16)
                for r in range(40 - h, 40):
                    holes = rng.choice(10, size=int(rng.integers(1, 4)), replace=False)
                    rows[r] = 0x3FF & ~int(sum(1 << int(c) for c in holes))
                p["rows"] = rows
            elif kind < 0.8:     # all-clear bait: one or two rows missing a few cells
                rows = np.zeros(40, dtype=np.uint16)
                start = int(rng.integers(0, 7))
                rows[39] = 0x3FF & ~(0xF << start)
                if rng.random() < 0.5:
                    rows[38] = rows[39]
                p["rows"] = rows
                p["piece"] = 4  # I
            if rng.random() < 0.4:
                k = int(rng.integers(1, 9))
                p["n_recv"] = k
                p["recv"][:k] = rng.integers(0, 10, size=k)
            if rng.random() < 0.5:
                p["held"] = int(rng.integers(0, 7))
            p["combo"] = int(rng.integers(0, 6)) if rng.random() < 0.5 else 0
            p["b2b"] = int(rng.integers(-1, 8))
            p["b2b_level"] = int(rng.integers(0, 3)) if ruleset == "s2" else int(rng.integers(0, 5))
        games[i]["turn"] = int(rng.integers(0, 2))
    return games


def pick_move(rng, rec, legal, seed):
    """Random legal move, biased (by trial-stepping copies through the oracle) towards moves
    that clear lines / register spins so that the attack and garbage paths are exercised."""
    if rng.random() < 0.35:
        return int(rng.choice(legal))
    trial = np.repeat(rec, legal.size)
    out = oracle.env_step(trial, legal.astype(np.uint16), False, seed)
    score = out["rows_cleared"].astype(float) * 2 + (out["flags"] & 3 != 0) * 1.5 + \
        (out["flags"] & 4 != 0) * 5 + out["attack"] + rng.random(legal.size) * 0.5
    top = np.argsort(-score)[:3]
    return int(legal[int(rng.choice(top))])


def sweep_env(n_games, plies, seed, ruleset="s2"):
    rng = np.random.default_rng(seed)
    tape = rh.install_tape(seed)
    games = random_midgame(rng, n_games, seed, ruleset)
    bad = moves_played = clears = sends = recvs = tops = holds = spins = pcs = 0
    for i in range(n_games):
        rec = games[i:i + 1].copy()
        ref = rh.make_game(rec[0])
        add_bag = bool(i % 2)
        for ply in range(plies):
            r0 = rec[0]
            mover = r0["players"][int(r0["turn"])]
            if mover["game_over"] or r0["players"][1 - int(r0["turn"])]["game_over"]:
                break
            cur = int(mover["piece"])
            alt = int(mover["held"]) if int(mover["held"]) != 255 else (
                int(mover["queue"][0]) if int(mover["qlen"]) > 0 else 255)
            if cur == 255 and int(mover["held"]) == 255:
                break  # Game.no_move
            mask = oracle.movegen_one(mover["rows"], cur, alt)[0]
            legal = np.flatnonzero(mask.reshape(-1))
            if legal.size == 0:
                break
            mv = pick_move(rng, rec, legal, seed)
            gid, pre_rng, pre_bag = int(r0["game_id"]), int(r0["rng_ctr"]), int(r0["bag_ctr"])
            out = oracle.env_step(rec, np.array([mv], dtype=np.uint16), add_bag, seed)[0]
            new_rng, new_bag = rh.step(ref, mv, add_bag, tape, gid, pre_rng, pre_bag)
            moves_played += 1
            clears += int(out["rows_cleared"] > 0)
            sends += int(out["attack"] > 0)
            recvs += int(bool(out["flags"] & 0x20))
            tops += int(bool(out["flags"] & 0x10))
            holds += int(bool(out["flags"] & 0x8))
            spins += int(bool(out["flags"] & 0x3))
            pcs += int(bool(out["flags"] & 0x4))
            a, b = game_state_dict(rec[0]), reference_game_state_dict(ref)
            if (a != b or new_rng != int(rec[0]["rng_ctr"]) or new_bag != int(rec[0]["bag_ctr"])
                    or out["status"] != 0):
                bad += 1
                if bad < 4:
                    print("  ENV MISMATCH game", i, "ply", ply, "move", mv, "status", out["status"])
                    for pl in range(2):
                        for k in a["players"][pl]:
                            if a["players"][pl][k] != b["players"][pl][k]:
                                print("    p", pl, k, a["players"][pl][k], "!=", b["players"][pl][k])
                    if a["turn"] != b["turn"]:
                        print("    turn", a["turn"], b["turn"])
                break
    print(f"env ({ruleset}): {moves_played} moves over {n_games} games; clears {clears}, all-clears {pcs}, attacks {sends}, "
          f"garbage receipts {recvs}, top-outs {tops}, holds {holds}, spin/mini flags {spins}; mismatches {bad}")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--movegen", type=int, default=7000)
    ap.add_argument("--games", type=int, default=300)
    ap.add_argument("--plies", type=int, default=60)
    ap.add_argument("--seed", type=int, default=20261018)
    args = ap.parse_args()
    if not rh.available():
        print("reference checkout not available; nothing to pin against")
        return 2
    t = time.time()
    bad = sweep_movegen(args.movegen, args.seed)
    for ruleset in ("s2", "s1"):
        bad += sweep_attack(ruleset)
        bad += sweep_env(args.games, args.plies, args.seed, ruleset)
    print(f"total mismatches {bad}  ({time.time() - t:.1f} s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
