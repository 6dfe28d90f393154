"""Constants and encodings of the hot path (host-side mirror of reference const.py).

Only what the data-generation path needs: board dimensions (const.py:6-12), the piece
alphabet (const.py:68), the policy-plane <-> (piece, rotation, t-spin index) maps
(const.py:72-136) and POLICY_SHAPE (const.py:122-123).  The pygame key / colour / screen
constants of the reference are UI and are deliberately absent.

The shape and kick tables used by the kernels live in csrc/trl_tables.cuh; the copies here
exist for host-side conversion (moves <-> tuples) and for tests.
"""
import numpy as np

ROWS = 40
COLS = 10
SPAWN_ROW = 23
SPAWN_Y = ROWS - SPAWN_ROW  # 17, piece.py:17-21
PREVIEWS = 5
MAX_MOVES = 1000

MINOS = "ZLOSIJT"
NONE = 255  # packed "no piece"

# piece matrix side (len(piece_dict[type]), const.py:138-175)
MATRIX_SIZE = {"Z": 3, "L": 3, "O": 2, "S": 3, "I": 4, "J": 3, "T": 3}

# rotations that own a policy plane (const.py:72-80)
policy_pieces = {"O": [0], "Z": [0, 1], "S": [0, 1], "I": [0, 1],
                 "L": [0, 1, 2, 3], "J": [0, 1, 2, 3], "T": [0, 1, 2, 3]}

_PLANE_ORDER = [("O", 1), ("Z", 2), ("S", 2), ("I", 2), ("L", 4), ("J", 4), ("T", 4)]


def _build_policy_maps():
    idx_to_piece, piece_to_idx, plane = {}, {}, 0
    for name, nrot in _PLANE_ORDER:
        piece_to_idx[name] = {r: {} for r in range(nrot)}
        for r in range(nrot):
            idx_to_piece[plane] = [name, r, 0]
            piece_to_idx[name][r][0] = plane
            plane += 1
    for tsi in (1, 2):  # T "rotation just occurred" / "... and used last kick" planes
        for r in range(4):
            idx_to_piece[plane] = ["T", r, tsi]
            piece_to_idx["T"][r][tsi] = plane
            plane += 1
    return idx_to_piece, piece_to_idx


policy_index_to_piece, policy_piece_to_index = _build_policy_maps()

POLICY_SHAPE = (len(policy_index_to_piece), ROWS - 1, COLS + 2 - 1)  # (27, 39, 11)
POLICY_SIZE = int(np.prod(POLICY_SHAPE))  # 11583
MASK_WORDS = (POLICY_SIZE + 31) // 32  # 362

# first plane / number of rotation planes per piece id (index into MINOS)
PLANE_BASE = np.array([policy_piece_to_index[m][0][0] for m in MINOS], dtype=np.int32)
PLANE_NROT = np.array([len(policy_pieces[m]) for m in MINOS], dtype=np.int32)

# mino_coords_dict (const.py:238-281) as [piece][rot][mino] = (col, row)
MINO_COORDS = {
    "Z": [[(0, 0), (1, 0), (1, 1), (2, 1)], [(1, 1), (1, 2), (2, 0), (2, 1)],
          [(0, 1), (1, 1), (1, 2), (2, 2)], [(0, 1), (0, 2), (1, 0), (1, 1)]],
    "L": [[(0, 1), (1, 1), (2, 0), (2, 1)], [(1, 0), (1, 1), (1, 2), (2, 2)],
          [(0, 1), (0, 2), (1, 1), (2, 1)], [(0, 0), (1, 0), (1, 1), (1, 2)]],
    "O": [[(0, 0), (0, 1), (1, 0), (1, 1)]] * 4,
    "S": [[(0, 1), (1, 0), (1, 1), (2, 0)], [(1, 0), (1, 1), (2, 1), (2, 2)],
          [(0, 2), (1, 1), (1, 2), (2, 1)], [(0, 0), (0, 1), (1, 1), (1, 2)]],
    "I": [[(0, 1), (1, 1), (2, 1), (3, 1)], [(2, 0), (2, 1), (2, 2), (2, 3)],
          [(0, 2), (1, 2), (2, 2), (3, 2)], [(1, 0), (1, 1), (1, 2), (1, 3)]],
    "J": [[(0, 0), (0, 1), (1, 1), (2, 1)], [(1, 0), (1, 1), (1, 2), (2, 0)],
          [(0, 1), (1, 1), (2, 1), (2, 2)], [(0, 2), (1, 0), (1, 1), (1, 2)]],
    "T": [[(0, 1), (1, 0), (1, 1), (2, 1)], [(1, 0), (1, 1), (1, 2), (2, 1)],
          [(0, 1), (1, 1), (1, 2), (2, 1)], [(0, 1), (1, 0), (1, 1), (1, 2)]],
}


def move_to_index(move):
    """(plane, col, row) reference move tuple (ai.py:1022) -> flat policy index."""
    plane, col, row = move
    return (int(plane) * POLICY_SHAPE[1] + int(row)) * POLICY_SHAPE[2] + int(col) + 2


def index_to_move(index):
    """flat policy index -> (plane, col, row) reference move tuple."""
    index = int(index)
    plane, rem = divmod(index, POLICY_SHAPE[1] * POLICY_SHAPE[2])
    row, col = divmod(rem, POLICY_SHAPE[2])
    return (plane, col - 2, row)
