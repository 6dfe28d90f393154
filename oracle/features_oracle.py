"""CPU ORACLE for the network-input encoding (TEST INFRASTRUCTURE): numpy restatement of
ai.game_to_X (reference ai.py:1364-1413) on packed TrlGame records, plus the deterministic
fake evaluator used to drive search-parity tests with bit-identical "network" outputs.

Parity status: pinned against the reference's game_to_X (oracle/pin_mcts_against_reference.py).
"""
import numpy as np

from tetris_reinforcement_learning_b200.const import COLS, POLICY_SIZE, PREVIEWS, ROWS


def encode(rec):
    """GAME_DTYPE scalar -> (grids float32 [2,40,10] side-to-move first, extras float32 [105])."""
    turn = int(rec["turn"])
    grids = np.zeros((2, ROWS, COLS), dtype=np.float32)
    extras = np.zeros(105, dtype=np.float32)
    for side, pl in enumerate((turn, 1 - turn)):
        p = rec["players"][pl]
        rows = np.asarray(p["rows"], dtype=np.uint16)
        grids[side] = ((rows[:, None] >> np.arange(COLS)[None, :]) & 1).astype(np.float32)
        table = np.zeros((2 + PREVIEWS, 7), dtype=np.float32)  # get_pieces, ai.py:1381-1392
        if int(p["piece"]) != 255:
            table[0, int(p["piece"])] = 1
        if int(p["held"]) != 255:
            table[1, int(p["held"])] = 1
        for j in range(min(int(p["qlen"]), PREVIEWS)):
            table[2 + j, int(p["queue"][j])] = 1
        base = side * 52
        extras[base:base + 49] = table.reshape(-1)
        extras[base + 49] = int(p["b2b"])
        extras[base + 50] = int(p["combo"])
        extras[base + 51] = int(p["n_recv"])
    extras[104] = turn  # players[turn].color == turn (player.py:253-261)
    return grids, extras


# ---- deterministic fake evaluator: integer hash of the features -> exact float32 outputs ----

_rng = np.random.default_rng(12345)
W_GRID = _rng.integers(1, 1 << 20, size=800, dtype=np.int64)
W_EXTRA = _rng.integers(1, 1 << 20, size=105, dtype=np.int64)
LOGIT_DIV = 512.0


def fake_hash(grids, extras):
    """int64 hash of integer-valued features; same formula in torch on the GPU side."""
    g = np.asarray(grids).reshape(-1).astype(np.int64)
    e = np.asarray(extras).reshape(-1).astype(np.int64) + 2  # b2b may be -1
    return int((g * W_GRID).sum() + (e * W_EXTRA).sum()) & 0x7FFFFFFF


def fake_outputs(h):
    """hash -> (value float32 in (0,1), logits float32[11583], multiples of 1/512 in [0,8))."""
    value = np.float32((h % 997) + 1) / np.float32(1000.0)
    j = np.arange(POLICY_SIZE, dtype=np.int64)
    logits = ((((h >> 3) * (2 * j + 1)) + 7 * j * j) % 4096).astype(np.float32) / np.float32(LOGIT_DIV)
    return value, logits


def fake_evaluate(rec):
    """The reference's `evaluate` contract on a packed game: (float value, float32 softmax (27,39,11))."""
    import torch
    grids, extras = encode(rec)
    value, logits = fake_outputs(fake_hash(grids, extras))
    policy = torch.softmax(torch.from_numpy(logits), dim=0).numpy().reshape(27, 39, 11)
    return float(value), policy
