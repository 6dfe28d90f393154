#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py

Every OUTPUT array in these files is computed by the reference's own code
(move_generation.get_move_matrix, Stats.get_attack, Game.make_move) imported from
/root/reference; this repo's oracle is only used to choose inputs (which moves to play)
and to supply the Philox tape that replaces the reference's `random` draws.

Files
  movegen_golden.npz  boards[n,40] u16, cur[n] u8, alt[n] u8, alt_is_held[n] u8,
                      mask_bits[n,362] u32 (bit-packed get_move_matrix(..., 'convolutional')),
                      fixture_names / fixture_counts (SURVEY A.8 known answers)
  attack_golden.npz   inputs[n,7] i16 (rows,tspin,mini,all_clear,combo,b2b,level),
                      outputs[n,4] i16 (attack, combo', b2b', level')
  env_golden.npz      before[n] / after[n] packed TrlGame bytes, moves[n] u16, add_bag[n] u8,
                      seed; `after` is the reference's state after Game.make_move
  attack_golden_s1.npz, env_golden_s1.npz   the same for ruleset 's1' 
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle, refharness as rh  # noqa: E402
from oracle.pin_against_reference import pick_move, random_midgame  # noqa: E402
from tetris_reinforcement_learning_b200 import synth  # noqa: E402
from tetris_reinforcement_learning_b200.state import GAME_DTYPE, pack_game, pack_mask  # noqa: E402

SEED = 20261018
OUT = os.path.join(ROOT, "tests", "golden")


def gen_movegen(n_boards=220):
    fb = rh.fixture_boards()
    boards, cur, alt, held, masks = [], [], [], [], []
    names, counts = [], []
    for name, rows in fb.items():
        row_counts = []
        for i in range(7):
            m = rh.movegen_packed(rows, i, i)
            boards.append(rows); cur.append(i); alt.append(i); held.append(1); masks.append(m)
            row_counts.append(int(m.sum()))
        names.append(name); counts.append(row_counts)
    b, c, a = synth.movegen_workload(n_boards, seed=SEED, caves=True)
    for j in range(b.shape[0]):
        cj, aj, mode = int(c[j]), int(a[j]), j % 4
        if mode == 3:
            aj = cj
        if mode == 2 and j % 8 == 2:
            cj = 255
        is_held = (mode != 1) or cj == 255
        m = rh.movegen_packed(b[j], cj, aj, alt_is_held=is_held)
        boards.append(b[j]); cur.append(cj); alt.append(aj); held.append(int(is_held)); masks.append(m)
    masks = pack_mask(np.stack(masks))
    np.savez_compressed(os.path.join(OUT, "movegen_golden.npz"),
                        boards=np.stack(boards).astype(np.uint16), cur=np.array(cur, np.uint8),
                        alt=np.array(alt, np.uint8), alt_is_held=np.array(held, np.uint8),
                        mask_bits=masks, fixture_names=np.array(names),
                        fixture_counts=np.array(counts, np.int32), seed=SEED)
    print("movegen_golden:", len(boards), "calls,", int(np.unpackbits(masks.view(np.uint8)).sum()), "placements")


def gen_attack(ruleset="s2"):
    m = rh.modules()
    ins, outs = [], []
    for rows in range(0, 5):
        for tspin in (0, 1):
            for mini in (0, 1):
                for pc in (0, 1):
                    for combo in range(0, 14):
                        for b2b in list(range(-1, 7)) + [23, 24, 66, 67, 1369, 1370]:
                            for lvl in (0, 1, 2, 3, 8):
                                s = m.stats.Stats(ruleset)
                                s.combo, s.b2b, s.b2b_level = combo, b2b, lvl
                                a = s.get_attack(rows, bool(tspin), bool(mini), bool(pc), "T")
                                ins.append((rows, tspin, mini, pc, combo, b2b, lvl))
                                outs.append((a, s.combo, s.b2b, s.b2b_level))
    name = "attack_golden.npz" if ruleset == "s2" else f"attack_golden_{ruleset}.npz"
    np.savez_compressed(os.path.join(OUT, name), inputs=np.array(ins, np.int16), outputs=np.array(outs, np.int16))
    print(name, len(ins), "cases")


def gen_env(n_games=260, plies=40, ruleset="s2"):
    rng = np.random.default_rng(SEED)
    tape = rh.install_tape(SEED)
    games = random_midgame(rng, n_games, SEED, ruleset)
    before, after, moves, bags = [], [], [], []
    stats = dict(clears=0, pcs=0, attacks=0, recv=0, tops=0, holds=0, spins=0)
    for i in range(n_games):
        rec = games[i:i + 1].copy()
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
                break
            # legal moves from the REFERENCE
            ref = rh.make_game(r0)
            legal = np.flatnonzero(rh.movegen(ref.players[ref.turn]).reshape(-1))
            if legal.size == 0:
                break
            mv = pick_move(rng, rec, legal, SEED)
            gid, pre_rng, pre_bag, rounds = int(r0["game_id"]), int(r0["rng_ctr"]), int(r0["bag_ctr"]), int(r0["rounds"])
            turn_before = int(r0["turn"])
            before.append(rec.copy())
            new_rng, new_bag = rh.step(ref, mv, add_bag, tape, gid, pre_rng, pre_bag)
            nxt = np.zeros(1, dtype=GAME_DTYPE)
            pack_game(ref, game_id=gid, rng_ctr=new_rng, bag_ctr=new_bag, out=nxt[0])   # ruleset from ref.ruleset
            # rounds = len(history.states): grows when player 1 places with add_history (game.py:86-87)
            nxt[0]["rounds"] = rounds + (1 if (add_bag and turn_before == 1) else 0)
            after.append(nxt.copy()); moves.append(mv); bags.append(int(add_bag))
            out = oracle.env_step(rec, np.array([mv], np.uint16), add_bag, SEED)[0]  # stats only
            stats["clears"] += int(out["rows_cleared"] > 0); stats["pcs"] += int(bool(out["flags"] & 4))
            stats["attacks"] += int(out["attack"] > 0); stats["recv"] += int(bool(out["flags"] & 0x20))
            stats["tops"] += int(bool(out["flags"] & 0x10)); stats["holds"] += int(bool(out["flags"] & 8))
            stats["spins"] += int(bool(out["flags"] & 3))
            rec = nxt  # continue from the REFERENCE's state
    before = np.concatenate(before); after = np.concatenate(after)
    name = "env_golden.npz" if ruleset == "s2" else f"env_golden_{ruleset}.npz"
    np.savez_compressed(os.path.join(OUT, name),
                        before=before.view(np.uint8).reshape(len(before), -1),
                        after=after.view(np.uint8).reshape(len(after), -1),
                        moves=np.array(moves, np.uint16), add_bag=np.array(bags, np.uint8), seed=SEED)
    print(name, len(moves), "transitions", stats)


def gen_mcts(n_searches=24, iters=48, out_name="mcts_golden.npz"):
    """mcts_golden.npz: reference ai.MCTS outputs under the fake evaluator and the Philox tape.
    mcts_golden_160.npz: the same at BASELINE's MAX_ITER=160 (playout-cap searches run 400 / 80 iterations)."""
    from oracle import features_oracle, mcts_oracle
    from oracle.pin_mcts_against_reference import config_families, positions, tanh_wrap
    m = rh.full_modules()
    fams = config_families(m.ai, iters)
    pos = positions(max(40, n_searches), SEED)   # one position (game id) per search
    recs, fam_idx, moves, saves, n_nodes, root_q = [], [], [], [], [], []
    child_moves = np.full((n_searches, 512), 0xFFFF, np.uint16)
    child_visits = np.zeros((n_searches, 512), np.int32)
    child_priors = np.zeros((n_searches, 512), np.float64)
    n_children = np.zeros(n_searches, np.int32)
    for k in range(n_searches):
        rec = pos[k % len(pos)]
        name, cfg = fams[k % len(fams)]
        ev = tanh_wrap(features_oracle.fake_evaluate) if cfg.use_tanh else features_oracle.fake_evaluate
        r = rh.reference_mcts(cfg, rec[0], ev, mcts_oracle.SearchTape(SEED, int(rec[0]["game_id"]), k))
        c = len(r["moves"])
        recs.append(rec.copy()); fam_idx.append(k % len(fams)); moves.append(r["move"]); saves.append(int(r["save"]))
        n_nodes.append(r["n_nodes"]); root_q.append(r["root_value_avg"]); n_children[k] = c
        child_moves[k, :c] = r["moves"]; child_visits[k, :c] = r["visits_post"]; child_priors[k, :c] = r["priors"]
    recs = np.concatenate(recs)
    np.savez_compressed(os.path.join(OUT, out_name),
                        games=recs.view(np.uint8).reshape(len(recs), -1), family=np.array(fam_idx, np.int32),
                        family_names=np.array([f[0] for f in fams]), iters=iters, seed=SEED,
                        move=np.array(moves, np.int32), save=np.array(saves, np.uint8),
                        n_nodes=np.array(n_nodes, np.int32), root_value_avg=np.array(root_q, np.float64),
                        n_children=n_children, child_moves=child_moves, child_visits=child_visits,
                        child_priors=child_priors)
    print(out_name, n_searches, "searches x", iters, "iterations")


if __name__ == "__main__":
    if not rh.available():
        sys.exit("reference checkout not available")
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["movegen", "attack", "env", "mcts", "s1"]
    if "movegen" in which:
        gen_movegen()
    if "attack" in which:
        gen_attack()
    if "env" in which:
        gen_env()
    if "mcts" in which:
        gen_mcts()
    if "mcts160" in which:
        gen_mcts(n_searches=64, iters=160, out_name="mcts_golden_160.npz")
    if "s1" in which:   # ruleset s1: attack table (stats.py:49-86) and env transitions without the all-spin rule
        gen_attack("s1")
        gen_env(n_games=140, plies=40, ruleset="s1")
