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


@pytest.mark.parametrize("family", ["alphasame", "baseresnet", "auxbaseresnet"])
def test_networks_are_state_dict_compatible_with_reference(family):
    """Same parameter names/shapes and the same forward as the reference's PyTorch nets."""
    import sys
    import torch
    ref_arch = rh.full_modules().architectures  # the reference's own module
    from tetris_reinforcement_learning_b200 import architectures as arch
    torch.manual_seed(0)
    if family == "alphasame":
        ref = ref_arch.AlphaSame(ref_arch.AlphaSameConfig(blocks=3, filters=8))
        mine = arch.AlphaSame(arch.AlphaSameConfig(blocks=3, filters=8))
    elif family == "baseresnet":
        ref = ref_arch.BaseResNet(ref_arch.BaseResNetConfig(blocks=2, filters=8))
        mine = arch.BaseResNet(arch.BaseResNetConfig(blocks=2, filters=8))
    else:
        ref = ref_arch.AuxBaseResNet(ref_arch.AuxBaseResNetConfig(blocks=2, filters=8))
        mine = arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=2, filters=8))
    # non-trivial BN statistics
    for m in ref.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.normal_(0, 0.5); m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.3)
    sd = ref.state_dict()
    assert list(sd.keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(sd, strict=True)
    ref.eval(); mine.eval()
    B = 5
    g = torch.Generator().manual_seed(1)
    ins = [
        (torch.rand(B, 1, 40, 10, generator=g) < 0.3).float(), (torch.rand(B, 7, 7, generator=g) < 0.15).float(),
        torch.randint(-1, 5, (B,), generator=g), torch.randint(0, 4, (B,), generator=g), torch.randint(0, 9, (B,), generator=g),
        (torch.rand(B, 1, 40, 10, generator=g) < 0.3).float(), (torch.rand(B, 7, 7, generator=g) < 0.15).float(),
        torch.randint(-1, 5, (B,), generator=g), torch.randint(0, 4, (B,), generator=g), torch.randint(0, 9, (B,), generator=g),
        torch.randint(0, 2, (B,), generator=g),
    ]
    with torch.no_grad():
        want = ref(*ins)
        got = mine(*ins)
        grids, extras = arch.pack_inputs(*ins)
        packed = mine.forward_packed(grids, extras)
    assert len(want) == len(got)
    for w, g_ in zip(want, got):
        assert torch.allclose(w, g_, atol=1e-6, rtol=1e-5)
    assert torch.allclose(want[0], packed[0], atol=1e-6) and torch.allclose(want[1], packed[1], atol=1e-5)


def test_sample_layout_and_reflections_match_reference():
    """samples_from_search / reflect_* / policy targets vs the reference's own play_game block
    (ai.py:1613-1666), search_statistics (ai.py:1330-1361) and reflect_policy (ai.py:1452-1516)."""
    import types
    import numpy as np
    from oracle import mcts_oracle, oracle
    from oracle.pin_mcts_against_reference import positions
    from tetris_reinforcement_learning_b200 import ai as mine
    from tetris_reinforcement_learning_b200.const import index_to_move
    ref_ai = rh.full_modules().ai
    rng = np.random.default_rng(0)
    for rec in positions(12, 77)[:8]:
        moves = mcts_oracle.legal_moves(rec[0])
        visits = rng.integers(0, 9, size=len(moves))
        visits[rng.integers(0, len(moves))] += 5
        # a mock reference tree holding the same root children
        tree = ref_ai.MCTSTree()
        tree.create_node(identifier="root", data=ref_ai.NodeState())
        for m, n in zip(moves, visits):
            st = ref_ai.NodeState(move=index_to_move(int(m)))
            st.visit_count = int(n)
            tree.create_node(data=st, parent="root")
        want_policy = ref_ai.search_statistics(tree)
        got = mine.policy_target_from_visits(moves, visits)
        assert np.array_equal(np.asarray(want_policy, dtype=np.float64), got)
        want_ref = ref_ai.reflect_policy(want_policy)
        assert np.array_equal(np.asarray(want_ref, dtype=np.float64), mine.reflect_policy_array(got))
        assert mine.reflect_policy(want_policy) == want_ref
        # full sample block
        game = rh.make_game(rec[0])
        move_data = [*ref_ai.game_to_X(game)]
        blocks = mine.samples_from_search(rec[0], moves, visits, augment=True)
        k = 0
        for a in range(2):
            for o in range(2):
                d = [f.copy() if isinstance(f, np.ndarray) else f for f in move_data]
                if a == 1:
                    d[0] = ref_ai.reflect_grid(d[0]); d[1] = ref_ai.reflect_pieces(d[1])
                if o == 1:
                    d[5] = ref_ai.reflect_grid(d[5]); d[6] = ref_ai.reflect_pieces(d[6])
                d = [f.tolist() if isinstance(f, np.ndarray) else f for f in d]
                d.append(want_policy if a == 0 else want_ref)
                assert blocks[k] == d
                k += 1
        plain = mine.samples_from_search(rec[0], moves, visits, augment=False)
        assert plain[0][:-1] == [f.tolist() if isinstance(f, np.ndarray) else f for f in move_data]
