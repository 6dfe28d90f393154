"""Synthetic s2 boards for parity sweeps and benchmarks (BASELINE config 2).

Three board families cycled by index, `numpy.random.default_rng(seed)`:
  F0  per-column height ~ U{0..15}, cells below the height filled with p = 0.85
  F1  n ~ U{0..13} garbage rows (one hole each) + 4 rows of p = 0.4 junk above them
  F2  bottom k ~ U{1..17} rows Bernoulli(0.5)
  F3  (optional, `caves=True`) adversarial sparse caves: bottom k ~ U{6..21} rows
      Bernoulli(p), p ~ U[0.15, 0.45] — the boards that maximise FIFO/fan-out sizes
Any accidentally full row gets one random hole.  Boards whose spawn cells collide are kept
(expected output: no placements for that piece).
"""
import numpy as np

from .const import COLS, MINOS, ROWS

DEFAULT_SEED = 20261018


def _pack(occ):
    w = (1 << np.arange(COLS)).astype(np.uint16)
    return (occ.astype(np.uint16) * w).sum(axis=-1).astype(np.uint16)


CHUNK = 16384  # boards are generated in fixed chunks so that any n is a prefix of the same stream


def _random_chunk(rng, n, caves):
    nfam = 4 if caves else 3
    fam = np.arange(n) % nfam
    occ = np.zeros((n, ROWS, COLS), dtype=bool)
    row_idx = np.arange(ROWS)[None, :, None]  # 0 = top

    idx = np.nonzero(fam == 0)[0]  # F0
    h = rng.integers(0, 16, size=(idx.size, 1, COLS))
    fill = rng.random((idx.size, ROWS, COLS), dtype=np.float32) < 0.85
    occ[idx] = (row_idx >= ROWS - h) & fill

    idx = np.nonzero(fam == 1)[0]  # F1
    g = rng.integers(0, 14, size=(idx.size, 1, 1))
    hole = rng.integers(0, COLS, size=(idx.size, ROWS, 1))
    garbage = (row_idx >= ROWS - g) & (np.arange(COLS)[None, None, :] != hole)
    junk_rows = (row_idx >= ROWS - g - 4) & (row_idx < ROWS - g)
    junk = junk_rows & (rng.random((idx.size, ROWS, COLS), dtype=np.float32) < 0.4)
    occ[idx] = garbage | junk

    idx = np.nonzero(fam == 2)[0]  # F2
    k = rng.integers(1, 18, size=(idx.size, 1, 1))
    occ[idx] = (row_idx >= ROWS - k) & (rng.random((idx.size, ROWS, COLS), dtype=np.float32) < 0.5)

    if caves:  # F3
        idx = np.nonzero(fam == 3)[0]
        k = rng.integers(6, 22, size=(idx.size, 1, 1))
        p = rng.uniform(0.15, 0.45, size=(idx.size, 1, 1))
        occ[idx] = (row_idx >= ROWS - k) & (rng.random((idx.size, ROWS, COLS), dtype=np.float32) < p)

    full = occ.all(axis=2)
    if full.any():
        bi, ri = np.nonzero(full)
        occ[bi, ri, rng.integers(0, COLS, size=bi.size)] = False
    return _pack(occ)


def random_boards(n, seed=DEFAULT_SEED, caves=False):
    """-> uint16[n, 40] bitrow boards; board i depends only on (seed, caves, i)."""
    out = np.empty((n, ROWS), dtype=np.uint16)
    for c, lo in enumerate(range(0, n, CHUNK)):
        rng = np.random.default_rng([int(seed), c, int(caves)])
        hi = min(n, lo + CHUNK)
        out[lo:hi] = _random_chunk(rng, CHUNK, caves)[: hi - lo]
    return out


def movegen_workload(n_boards, seed=DEFAULT_SEED, caves=False, hold_shift=None):
    """BASELINE config 2: every board x 7 current pieces with a deterministic second piece.

    Returns (boards uint16[n_boards*7, 40], cur uint8[n_boards*7], alt uint8[n_boards*7]);
    call j = 7*i + k uses board i, cur = k, alt = (k + 1 + i % 6) % 7 (always != cur, so every
    call enumerates two piece types, exercising _get_piece_types_to_check)."""
    boards = random_boards(n_boards, seed=seed, caves=caves)
    i = np.repeat(np.arange(n_boards), 7)
    k = np.tile(np.arange(7), n_boards)
    shift = (1 + i % 6) if hold_shift is None else hold_shift
    cur = k.astype(np.uint8)
    alt = ((k + shift) % 7).astype(np.uint8)
    return np.repeat(boards, 7, axis=0), cur, alt


def piece_letters(ids):
    return [MINOS[int(i)] for i in ids]
