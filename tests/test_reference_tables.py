"""Build-container only: host constants agree with the reference's const.py (skipped where
/root/reference is absent, e.g. on the GPU box)."""
import pytest

from oracle import refharness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")


def test_const_tables_match_reference():
    from tetris_reinforcement_learning_b200 import const
    ref = rh.modules().const
    assert (const.ROWS, const.COLS, const.SPAWN_ROW, const.PREVIEWS, const.MAX_MOVES) == \
        (ref.ROWS, ref.COLS, ref.SPAWN_ROW, ref.PREVIEWS, ref.MAX_MOVES)
    assert const.MINOS == ref.MINOS
    assert const.POLICY_SHAPE == tuple(ref.POLICY_SHAPE) and const.POLICY_SIZE == int(ref.POLICY_SIZE)
    assert const.policy_index_to_piece == ref.policy_index_to_piece
    assert const.policy_piece_to_index == ref.policy_piece_to_index
    assert const.policy_pieces == ref.policy_pieces
    for t in const.MINOS:
        assert const.MATRIX_SIZE[t] == len(ref.piece_dict[t])
        for r in range(4):
            assert [tuple(c) for c in ref.mino_coords_dict[t][r]] == list(const.MINO_COORDS[t][r])


def test_oracle_spot_check_against_live_reference(oracle):
    """A few live calls so the container CI notices a drifted reference checkout."""
    import numpy as np
    from tetris_reinforcement_learning_b200 import synth
    boards, cur, alt = synth.movegen_workload(6, seed=99, caves=True)
    for j in range(boards.shape[0]):
        ref = rh.movegen_packed(boards[j], int(cur[j]), int(alt[j]))
        mine = oracle.movegen_one(boards[j], int(cur[j]), int(alt[j]))[0]
        assert np.array_equal(ref, mine)
