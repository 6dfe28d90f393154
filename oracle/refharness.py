"""Import the UNMODIFIED reference from /root/reference and drive it with tape RNG.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU box,
so nothing in `-m gpu` tests, smoke() or bench.py imports this module.  It is used by
oracle/pin_against_reference.py (oracle vs reference sweeps) and oracle/gen_golden.py
(reference-generated fixtures committed under tests/golden/).

Recipe: put /root/reference on sys.path and pre-seed sys.modules with a dummy
`pygame` exposing init() and the K_* names const.py:15-30 reads; then const, board, piece,
player, stats, piece_queue, game and move_generation import and run unmodified.
"""
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("TRL_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "move_generation.py"))


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return 0


_mods = None


def modules():
    """-> namespace with the reference's game modules (imported once)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    if "pygame" not in sys.modules:
        pg = _Permissive("pygame")
        pg.init = lambda *a, **k: None
        sys.modules["pygame"] = pg
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import board, const, game, move_generation, piece, piece_queue, player, stats  # noqa: E401
    _mods = types.SimpleNamespace(const=const, board=board, piece=piece, player=player, stats=stats,
                                  piece_queue=piece_queue, game=game, move_generation=move_generation)
    return _mods


class _Dummy:
    """Callable, attribute-able stand-in for anything reached inside a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _DummyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


_full = None


def full_modules(workdir="/tmp/trl_ref_work"):
    """Import ALL of the reference incl. ai.py / architectures.py (MCTS, play_game, networks).

    SURVEY §8c recipe: permissive dummies for pygame / tensorflow / keras / matplotlib, ujson ->
    stdlib json, and a cwd whose parent holds a `Storage/` directory (ai.py:57-59 mkdirs
    `Storage/logs` at import time).  NOTE: changes the process cwd to `workdir`/run."""
    global _full
    if _full is not None:
        return _full
    m = modules()
    import json
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.python", "tensorflow.python.ops",
                 "tensorflow.python.ops.math_ops", "keras", "keras.backend", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = _DummyModule(name)
    sys.modules.setdefault("ujson", json)
    os.makedirs(os.path.join(workdir, "Storage"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "run"), exist_ok=True)
    os.chdir(os.path.join(workdir, "run"))
    import ai
    import architectures
    m.ai, m.architectures = ai, architectures
    _full = m
    return m


# ---------------------------------------------------------------------------------------
# tape RNG: the reference's `random` draws replaced by the same Philox streams the oracle
# and the CUDA kernels use
# ---------------------------------------------------------------------------------------

class TapeRandom:
    """Stands in for the `random` module inside player.py and piece_queue.py."""

    def __init__(self, seed):
        from oracle import oracle
        self._o = oracle
        self.seed = seed
        self.game_id = 0
        self.rng_ctr = 0
        self.bag_ctr = 0
        self._bag_player = 0
        self.stream = 0

    def bind(self, game_id, rng_ctr, bag_ctr, stream=0):
        self.game_id, self.rng_ctr, self.bag_ctr, self._bag_player = game_id, rng_ctr, bag_ctr, 0
        self.stream = stream

    def randint(self, a, b):  # player.py:185
        assert (a, b) == (0, 9)
        col = self._o.garbage_column(self.seed, self.game_id, self.rng_ctr, self.stream)
        self.rng_ctr += 1
        return col

    def shuffle(self, lst):  # piece_queue.py:20 (called for player 0 then 1, game.py:34-38)
        assert lst == list("ZLOSIJT")
        bag = self._o.generate_bag(self.seed, self.game_id, self.bag_ctr, self._bag_player)
        lst[:] = ["ZLOSIJT"[int(i)] for i in bag]
        self._bag_player += 1
        if self._bag_player == 2:
            self._bag_player = 0
            self.bag_ctr += 1


def install_tape(seed):
    m = modules()
    tape = TapeRandom(seed)
    m.player.random = tape
    m.piece_queue.random = tape
    return tape


# ---------------------------------------------------------------------------------------
# building reference objects from packed state
# ---------------------------------------------------------------------------------------

def _set_board(player, rows):
    from tetris_reinforcement_learning_b200.state import rows_to_grid
    grid = rows_to_grid(rows)
    player.board.grid = np.where(grid != 0, 1, 0).astype(object)


def make_player(rows, piece, held, queue, ruleset="s2"):
    """Reference Player with the given board / active piece (at spawn) / hold / queue.
    piece, held: 'Z'..'T' or None; queue: list of letters."""
    m = modules()
    p = m.player.Player(ruleset)
    p.color = 0
    _set_board(p, rows)
    p.queue.pieces = list(queue)
    p.held_piece = held
    if piece is not None:
        pc = m.piece.Piece(m.const.piece_dict[piece], type=piece)
        pc.move_to_spawn()
        p.piece = pc
    return p


def movegen(player, algo="convolutional"):
    """bool (27,39,11) from the reference's get_move_matrix."""
    return np.asarray(modules().move_generation.get_move_matrix(player, algo=algo)).astype(bool)


def movegen_packed(rows, cur, alt, alt_is_held=True):
    """Reference mask for the packed (rows, cur, alt) triple the C ABI takes.
    alt is presented to the reference as the held piece (alt_is_held) or as queue[0]."""
    from tetris_reinforcement_learning_b200.state import piece_name
    c, a = piece_name(cur), piece_name(alt)
    if alt_is_held:
        pl = make_player(rows, c, a, [])
    else:
        pl = make_player(rows, c, None, [a] if a is not None else [])
    return movegen(pl)


def make_game(rec, ruleset=None):
    """Reference Game from a GAME_DTYPE scalar (ruleset from the record unless given)."""
    from tetris_reinforcement_learning_b200.state import piece_name
    m = modules()
    if ruleset is None:
        ruleset = "s1" if int(rec["ruleset"]) == 1 else "s2"
    g = m.game.Game(ruleset)
    for i, pl in enumerate(g.players):
        pr = rec["players"][i]
        _set_board(pl, pr["rows"])
        pl.queue.pieces = ["ZLOSIJT"[int(v)] for v in pr["queue"][:int(pr["qlen"])]]
        pl.held_piece = piece_name(pr["held"])
        pl.piece = None
        if int(pr["piece"]) != 255:
            t = piece_name(pr["piece"])
            pc = m.piece.Piece(m.const.piece_dict[t], type=t)
            pc.move_to_spawn()
            pl.piece = pc
        pl.game_over = bool(pr["game_over"])
        pl.garbage_to_receive = [int(v) for v in pr["recv"][:int(pr["n_recv"])]]
        pl.stats.pieces = int(pr["pieces"])
        pl.stats.b2b = int(pr["b2b"])
        pl.stats.b2b_level = int(pr["b2b_level"])
        pl.stats.combo = int(pr["combo"])
    g.turn = int(rec["turn"])
    return g


def step(game, move_index, add_bag, tape, game_id, rng_ctr, bag_ctr):
    """Game.make_move(move, add_bag, add_history=False) with tape RNG; returns new counters."""
    from tetris_reinforcement_learning_b200.const import index_to_move
    tape.bind(game_id, rng_ctr, bag_ctr)
    game.make_move(index_to_move(move_index), add_bag=add_bag, add_history=False)
    return tape.rng_ctr, tape.bag_ctr


def fixture_boards():
    """The four hand-made boards of util.py:34-37, top-padded to 40 rows, plus
    the empty board.  util.py itself cannot be imported (imageio/matplotlib), so the
    literals are extracted with ast."""
    import ast
    src = open(os.path.join(REFERENCE_DIR, "util.py")).read().splitlines()
    out = {"empty": np.zeros(40, dtype=np.uint16)}
    from tetris_reinforcement_learning_b200.state import grid_to_rows
    for line in src[30:40]:
        line = line.strip()
        for name in ("util_t_spin_board", "util_z_spin_board", "util_move_algo_board_2", "util_move_algo_board"):
            if line.startswith(name + " ="):
                grid = ast.literal_eval(line.split("=", 1)[1].strip())
                arr = np.array([[0 if c == 0 else 1 for c in row] for row in grid], dtype=np.int8)
                pad = np.zeros((40 - arr.shape[0], 10), dtype=np.int8)
                out[name] = grid_to_rows(np.vstack([pad, arr]))
                break
    return out


# ---------------------------------------------------------------------------------------
# driving the reference's MCTS with an injected evaluator and tape RNG
# ---------------------------------------------------------------------------------------

class _AiRandomShim:
    """Stands in for the `random` module inside ai.py: random() is the playout-cap coin
    (ai.py:324); choices() restates CPython's random.choices with the tape's single uniform."""

    def __init__(self, tape):
        self.tape = tape

    def random(self):
        return self.tape.coin()

    def choices(self, population, weights=None, k=1):
        from bisect import bisect
        from oracle.mcts_oracle import _accumulate
        cum = _accumulate(weights)
        total = cum[-1] + 0.0
        return [population[bisect(cum, self.tape.choice_uniform() * total, 0, len(population) - 1)]]


class _NumpyProxy:
    """numpy with random.gamma redirected to the tape (ai.py:489)."""

    def __init__(self, tape):
        import numpy
        self._np = numpy
        self.random = types.SimpleNamespace(gamma=lambda shape, scale, size: tape.gamma(shape, size))

    def __getattr__(self, name):
        return getattr(self._np, name)


class _SearchGarbageShim:
    """`random` inside player.py during a search: sequential draws from the search stream."""

    def __init__(self, tape):
        self.tape = tape

    def randint(self, a, b):
        from oracle import oracle
        col = oracle.garbage_column(self.tape.seed, self.tape.game_id, self.tape.garbage_ctr, self.tape.garbage_stream)
        self.tape.garbage_ctr += 1
        return col


def reference_mcts(config, game_rec, evaluate, tape):
    """Run the reference's own MCTS (ai.py:299-659) from a packed game with `evaluate(rec)` as
    the network and `tape` (oracle.mcts_oracle.SearchTape) as every RNG.  Returns the same
    dict layout as mcts_oracle.search (without pre-prune visits)."""
    from tetris_reinforcement_learning_b200.const import move_to_index
    from tetris_reinforcement_learning_b200.state import GAME_DTYPE, pack_game
    m = full_modules()
    ai = m.ai
    game = make_game(game_rec)

    def evaluate_hook(cfg, g, net):
        rec = np.zeros((), dtype=GAME_DTYPE)
        pack_game(g, out=rec)
        return evaluate(rec)

    saved = (ai.evaluate, ai.random, ai.np, m.player.random)
    ai.evaluate, ai.random, ai.np = evaluate_hook, _AiRandomShim(tape), _NumpyProxy(tape)
    m.player.random = _SearchGarbageShim(tape)
    try:
        move, tree, save = ai.MCTS(config, game, None)
    finally:
        ai.evaluate, ai.random, ai.np, m.player.random = saved
    root = tree.get_node("root")
    kids = [tree.get_node(c).data for c in root.successors(tree.identifier)]
    return {"move": move_to_index(move), "moves": [move_to_index(k.move) for k in kids],
            "visits_post": [int(k.visit_count) for k in kids], "priors": [float(k.policy) for k in kids],
            "save": bool(save), "n_nodes": tree._counter + 1, "root_visits": int(root.data.visit_count),
            "root_value_avg": float(root.data.value_avg), "garbage_draws": tape.garbage_ctr}


def reference_game_to_X(game_rec):
    """ai.game_to_X on a packed game -> (grids [2,40,10], extras [105]) in this repo's layout."""
    m = full_modules()
    x = m.ai.game_to_X(make_game(game_rec))
    grids = np.stack([np.asarray(x[0], dtype=np.float32), np.asarray(x[5], dtype=np.float32)])
    extras = np.concatenate([np.asarray(x[1], np.float32).reshape(-1), [x[2], x[3], x[4]],
                             np.asarray(x[6], np.float32).reshape(-1), [x[7], x[8], x[9]], [x[10]]]).astype(np.float32)
    return grids, extras
