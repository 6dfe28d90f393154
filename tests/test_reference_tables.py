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
