"""GPU test of the drop-in API end to end (reference simulation.py:19-28 / ai.self_play_loop body):
instantiate_network -> make_training_set (data set on disk in the reference's format) ->
load_data_and_train_model -> battle_networks, in a temporary Storage directory."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family", ["alphasame", "aux"])
def test_selfplay_train_gate_loop(tmp_path, monkeypatch, family):
    import torch
    from tetris_reinforcement_learning_b200 import ai, training
    from tetris_reinforcement_learning_b200 import architectures as arch
    monkeypatch.setenv("TRL_STORAGE", str(tmp_path / "Storage"))
    mc = arch.AlphaSameConfig(blocks=2, filters=16) if family == "alphasame" else arch.AuxBaseResNetConfig(blocks=2, filters=16)
    cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=6, training_games=6,
                    battle_games=6, epochs=1, batch_size=32, use_playout_cap_randomization=False)
    torch.manual_seed(0)
    net = ai.instantiate_network(cfg, show_summary=False, save_network=True)
    assert ai.highest_model_number(cfg) == 0 and os.path.exists(f"{cfg.model_dir}/0.pt")
    assert os.path.exists(f"{cfg.model_dir}/versions.jsonl")

    train_cfg = cfg.copy()
    train_cfg.training = True
    infer = ai.get_interference_network(cfg, net)
    os.makedirs(cfg.data_dir, exist_ok=True)
    assert ai.make_training_set(train_cfg, infer, num_games=6, save_game=True, save_stats=True, seed=1) is None
    assert ai.highest_data_number(cfg) == 0
    data = json.load(open(f"{cfg.data_dir}/0.txt"))
    assert len(data) > 0 and len(data) % 4 == 0                      # x4 mirror augmentation (ai.py:1613-1666)
    s = data[0]
    assert len(s) == 13                                              # 11 features + outcome + policy (ai.py:1675-1699)
    assert np.asarray(s[0]).shape == (40, 10) and np.asarray(s[1]).shape == (7, 7)
    assert s[-2] in (cfg.value_min, cfg.value_mid, cfg.value_max)
    pol = np.asarray(s[-1])
    assert pol.shape == (27, 39, 11) and abs(pol.sum() - 1.0) < 0.05 and (pol >= 0).all()
    stats = [json.loads(l) for l in open(ai.logs_dir() / "stats.jsonl")]
    assert len(stats) == 1 and {"app", "dspp", "model_number", "data_number"} <= set(stats[0])

    before = {k: v.clone() for k, v in net.state_dict().items()}
    training.load_data_and_train_model(cfg, net)
    changed = sum(int(not torch.equal(before[k].cpu(), v.cpu())) for k, v in net.state_dict().items())
    assert changed > 0

    challenger = ai.get_interference_network(cfg, net)
    wins, accepted = training.battle_networks(challenger, cfg, infer, cfg, cfg.gating_threshold, cfg.gating_threshold_type,
                                              cfg.battle_games, seed=2)
    assert wins.sum() == cfg.battle_games and accepted in (True, False)


def test_compact_data_format_holds_the_same_samples_and_trains(tmp_path, monkeypatch):
    """make_training_set(data_format="compact") with the same seed: <n>.npz expands to exactly the tensors of the JSON set
    <n>.txt (same games, same order), and load_data_and_train_model / the training step consume it on the GPU."""
    import torch
    from tetris_reinforcement_learning_b200 import ai, training
    from tetris_reinforcement_learning_b200 import architectures as arch
    from tetris_reinforcement_learning_b200.compact import CompactSet
    monkeypatch.setenv("TRL_STORAGE", str(tmp_path / "Storage"))
    cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=arch.AlphaSameConfig(blocks=2, filters=16),
                    MAX_ITER=6, training_games=8, epochs=1, batch_size=64, use_playout_cap_randomization=False, training=True)
    torch.manual_seed(0)
    net = ai.instantiate_network(cfg, show_summary=False, save_network=True)
    infer = ai.get_interference_network(cfg, net)
    os.makedirs(cfg.data_dir, exist_ok=True)
    ai.make_training_set(cfg, infer, num_games=8, save_game=True, save_stats=True, seed=5)
    ccfg = cfg.copy()
    ccfg.engine_data_format = "compact"
    assert ccfg.copy().engine_data_format == "compact"
    ai.make_training_set(ccfg, infer, num_games=8, save_game=True, save_stats=True, seed=5)
    assert ai.highest_data_number(cfg) == 1 and os.path.exists(f"{cfg.data_dir}/1.npz")
    want = training._to_tensors(json.load(open(f"{cfg.data_dir}/0.txt")))
    cs = CompactSet.load(f"{cfg.data_dir}/1.npz")
    got = cs.batch_tensors(np.arange(len(cs)), device="cuda:0")
    assert len(cs) == len(want[0]) > 0 and os.path.getsize(f"{cfg.data_dir}/1.npz") * 50 < os.path.getsize(f"{cfg.data_dir}/0.txt")
    for col, (w, t) in enumerate(zip(want, got)):
        if col == 11:          # outcomes: torch.tensor() of the JSON column is int64 when no game was drawn; training casts to float
            w = w.float()
        assert w.shape == t.shape and w.dtype == t.dtype and torch.equal(w, t.cpu())
    os.remove(f"{cfg.data_dir}/0.txt")                      # train on the compact set alone
    before = {k: v.clone() for k, v in net.state_dict().items()}
    out = training.load_data_and_train_model(ccfg, net)
    assert np.isfinite(out["loss"]) and sum(int(not torch.equal(before[k], v)) for k, v in net.state_dict().items()) > 0


def test_mcts_dropin_on_a_reference_style_game():
    """ai.MCTS(config, game, net) on a duck-typed Game: returns a legal (plane, col, row) move."""
    import torch
    from tetris_reinforcement_learning_b200 import ai, env
    from tetris_reinforcement_learning_b200 import architectures as arch
    from tetris_reinforcement_learning_b200.const import move_to_index
    from tetris_reinforcement_learning_b200.move_generation import movegen_host
    from types import SimpleNamespace as NS
    from tetris_reinforcement_learning_b200.const import MINOS
    from tetris_reinforcement_learning_b200.state import rows_to_grid

    def unpack_game(rec):   # a duck-typed reference Game (game.py:6-26, player.py:10-27)
        players = []
        for pr in rec["players"]:
            players.append(NS(board=NS(grid=rows_to_grid(pr["rows"])), queue=NS(pieces=[MINOS[int(v)] for v in pr["queue"][:int(pr["qlen"])]]),
                              piece=None if int(pr["piece"]) == 255 else NS(type=MINOS[int(pr["piece"])]),
                              held_piece=None if int(pr["held"]) == 255 else MINOS[int(pr["held"])],
                              game_over=bool(pr["game_over"]), garbage_to_receive=[int(v) for v in pr["recv"][:int(pr["n_recv"])]],
                              stats=NS(pieces=int(pr["pieces"]), b2b=int(pr["b2b"]), combo=int(pr["combo"]), b2b_level=int(pr["b2b_level"]))))
        return NS(players=players, turn=int(rec["turn"]), ruleset="s2")
    mc = arch.AlphaSameConfig(blocks=2, filters=16)
    cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=12, training=False)
    torch.manual_seed(0)
    net = arch.AlphaSame(mc).to("cuda:0").eval()
    rec = env.game_setup_host(1, 5, 9)
    game = unpack_game(rec[0])
    move, tree, save = ai.MCTS(cfg, game, net, seed=3)
    pl = rec[0]["players"][0]
    legal = movegen_host(pl["rows"][None, :], np.array([pl["piece"]], np.uint8), np.array([pl["queue"][0]], np.uint8),
                         want_mask=False, want_moves=True)
    assert move_to_index(move) in set(int(m) for m in legal["moves"][0][: int(legal["n_moves"][0])])
    assert int(tree.visits_pre.sum()) == cfg.MAX_ITER - 1 and save is True
