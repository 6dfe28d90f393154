"""ctypes binding of the CPU oracle (oracle/trl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Parity status: pinned against the imported reference (oracle/pin_against_reference.py) and
against the reference-generated vectors in tests/golden/ (tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

from tetris_reinforcement_learning_b200.const import MASK_WORDS, POLICY_SIZE, ROWS
from tetris_reinforcement_learning_b200.state import GAME_DTYPE, STEPOUT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtrl_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/trl_oracle.c with gcc (make).  Building the checker is not using it."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("trl_oracle.c", "trl_oracle.h"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u16p, u8p, u32p = (ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint8),
                           ctypes.POINTER(ctypes.c_uint32))
        ip = ctypes.POINTER(ctypes.c_int)
        L.trl_oracle_movegen.restype = ctypes.c_int
        L.trl_oracle_movegen.argtypes = [u16p, ctypes.c_int, ctypes.c_int, u8p, u32p, ip, ip]
        L.trl_oracle_movegen_batch.restype = ctypes.c_longlong
        L.trl_oracle_movegen_batch.argtypes = [u16p, u8p, u8p, ctypes.c_int, u32p, u16p, u32p, ctypes.c_int]
        L.trl_oracle_get_attack_s2.restype = ctypes.c_int
        L.trl_oracle_get_attack_s2.argtypes = [ctypes.c_int] * 4 + [ip, ip, ip]
        L.trl_oracle_get_attack_s1.restype = ctypes.c_int
        L.trl_oracle_get_attack_s1.argtypes = [ctypes.c_int] * 4 + [ip, ip, ip]
        L.trl_oracle_env_step_batch.restype = None
        L.trl_oracle_env_step_batch.argtypes = [ctypes.c_void_p, u16p, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_int, ctypes.c_uint64]
        L.trl_oracle_game_setup_batch.restype = None
        L.trl_oracle_game_setup_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64]
        L.trl_oracle_garbage_column.restype = ctypes.c_int
        L.trl_oracle_garbage_column.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.trl_oracle_env_step_rng.restype = None
        L.trl_oracle_env_step_rng.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32,
                                              u32p, ctypes.c_void_p]
        L.trl_oracle_generate_bag.restype = None
        L.trl_oracle_generate_bag.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, u8p]
        L.trl_oracle_uniform.restype = ctypes.c_double
        L.trl_oracle_uniform.argtypes = [ctypes.c_uint64] + [ctypes.c_uint32] * 4
        L.trl_oracle_gamma.restype = ctypes.c_double
        L.trl_oracle_gamma.argtypes = [ctypes.c_uint64] + [ctypes.c_uint32] * 3 + [ctypes.c_double]
        L.trl_oracle_philox.restype = None
        L.trl_oracle_philox.argtypes = [ctypes.c_uint64] + [ctypes.c_uint32] * 4 + [u32p]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def movegen_one(rows, cur, alt):
    """-> (bool mask (27,39,11), status, n_push, n_emit)."""
    from tetris_reinforcement_learning_b200.const import POLICY_SHAPE
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    mask = np.zeros(POLICY_SIZE, dtype=np.uint8)
    st = ctypes.c_uint32(0)
    npush, nemit = ctypes.c_int(0), ctypes.c_int(0)
    lib().trl_oracle_movegen(_p(rows, ctypes.c_uint16), int(cur), int(alt), _p(mask, ctypes.c_uint8),
                             ctypes.byref(st), ctypes.byref(npush), ctypes.byref(nemit))
    return mask.reshape(POLICY_SHAPE).astype(bool), st.value, npush.value, nemit.value


def movegen_batch(boards, cur, alt, n_threads=1, want_masks=True):
    """boards uint16[n,40], cur/alt uint8[n] -> (mask_bits uint32[n,362] | None, n_moves, status, total)."""
    boards = np.ascontiguousarray(boards, dtype=np.uint16)
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    alt = np.ascontiguousarray(alt, dtype=np.uint8)
    n = boards.shape[0]
    assert boards.shape == (n, ROWS) and cur.shape == (n,) and alt.shape == (n,)
    masks = np.zeros((n, MASK_WORDS), dtype=np.uint32) if want_masks else None
    n_moves = np.zeros(n, dtype=np.uint16)
    status = np.zeros(n, dtype=np.uint32)
    total = lib().trl_oracle_movegen_batch(
        _p(boards, ctypes.c_uint16), _p(cur, ctypes.c_uint8), _p(alt, ctypes.c_uint8), n,
        _p(masks, ctypes.c_uint32) if want_masks else None, _p(n_moves, ctypes.c_uint16),
        _p(status, ctypes.c_uint32), int(n_threads))
    return masks, n_moves, status, int(total)


def get_attack_s2(rows_cleared, tspin, mini, all_clear, combo, b2b, b2b_level):
    """-> (attack, combo, b2b, b2b_level) after Stats.get_attack (ruleset s2)."""
    c, b, l = ctypes.c_int(combo), ctypes.c_int(b2b), ctypes.c_int(b2b_level)
    a = lib().trl_oracle_get_attack_s2(int(rows_cleared), int(bool(tspin)), int(bool(mini)),
                                       int(bool(all_clear)), ctypes.byref(c), ctypes.byref(b), ctypes.byref(l))
    return a, c.value, b.value, l.value


def get_attack_s1(rows_cleared, tspin, mini, all_clear, combo, b2b, b2b_level):
    """-> (attack, combo, b2b, b2b_level) after Stats.get_attack (ruleset s1)."""
    c, b, l = ctypes.c_int(combo), ctypes.c_int(b2b), ctypes.c_int(b2b_level)
    a = lib().trl_oracle_get_attack_s1(int(rows_cleared), int(bool(tspin)), int(bool(mini)),
                                       int(bool(all_clear)), ctypes.byref(c), ctypes.byref(b), ctypes.byref(l))
    return a, c.value, b.value, l.value


def env_step(games, moves, add_bag, seed):
    """Steps GAME_DTYPE array `games` in place; returns STEPOUT_DTYPE array."""
    assert games.dtype == GAME_DTYPE and games.flags["C_CONTIGUOUS"]
    moves = np.ascontiguousarray(moves, dtype=np.uint16)
    out = np.zeros(games.shape[0], dtype=STEPOUT_DTYPE)
    lib().trl_oracle_env_step_batch(games.ctypes.data, _p(moves, ctypes.c_uint16), games.shape[0],
                                    out.ctypes.data, int(bool(add_bag)), int(seed))
    return out


def game_setup(n, first_game_id, seed, ruleset="s2"):
    from tetris_reinforcement_learning_b200.state import ruleset_id
    games = np.zeros(n, dtype=GAME_DTYPE)
    lib().trl_oracle_game_setup_batch(games.ctypes.data, n, int(first_game_id), int(seed))
    games["ruleset"] = ruleset_id(ruleset)
    return games


def garbage_column(seed, game_id, ctr, stream=0):
    return lib().trl_oracle_garbage_column(int(seed), int(game_id), int(stream), int(ctr))


def env_step_rng(game, move, add_bag, seed, stream, ctr):
    """One game (GAME_DTYPE array of length 1) stepped with an explicit garbage-RNG stream and
    counter; returns (STEPOUT record, new counter)."""
    assert game.dtype == GAME_DTYPE and game.shape == (1,)
    out = np.zeros(1, dtype=STEPOUT_DTYPE)
    c = ctypes.c_uint32(int(ctr))
    lib().trl_oracle_env_step_rng(game.ctypes.data, int(move), int(bool(add_bag)), int(seed), int(stream),
                                  ctypes.byref(c), out.ctypes.data)
    return out[0], c.value


def generate_bag(seed, game_id, bag_ctr, player):
    bag = np.zeros(7, dtype=np.uint8)
    lib().trl_oracle_generate_bag(int(seed), int(game_id), int(bag_ctr), int(player), _p(bag, ctypes.c_uint8))
    return bag


def philox(seed, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().trl_oracle_philox(int(seed), int(c0), int(c1), int(c2), int(c3), _p(out, ctypes.c_uint32))
    return out


def uniform(seed, game_id, search_no, purpose, idx=0):
    return lib().trl_oracle_uniform(int(seed), int(game_id), int(search_no), int(purpose), int(idx))


def gamma(seed, game_id, search_no, child, alpha):
    return lib().trl_oracle_gamma(int(seed), int(game_id), int(search_no), int(child), float(alpha))
