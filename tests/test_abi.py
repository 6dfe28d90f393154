"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/trl.h declares (no compute calls here — there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_so():
    from tetris_reinforcement_learning_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "trl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for name in ("trl_movegen", "trl_movegen_host", "trl_movegen_games", "trl_env_step",
                 "trl_env_step_host", "trl_game_setup", "trl_abi_version", "trl_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol(built_so):
    L = ctypes.CDLL(built_so)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/trl.h but not exported"


def test_binding_signatures_cover_header(built_so):
    from tetris_reinforcement_learning_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()
    L = _native.lib()  # checks ABI version and struct sizes against state.py
    assert L.trl_abi_version() == _native.ABI_VERSION
    assert L.trl_sizeof_player() == 192 and L.trl_sizeof_game() == 400


def test_bad_arguments_are_rejected_without_a_gpu(built_so):
    from tetris_reinforcement_learning_b200 import _native
    L = _native.lib()
    assert L.trl_movegen(None, None, None, 4, None, None, 0, None, None, None) == -1
    assert L.trl_movegen_host(None, None, None, 4, None, None, 0, None, None) == -1
    assert L.trl_env_step(None, None, 1, None, 0, 0, None) == -1
    assert L.trl_movegen(None, None, None, 0, None, None, 0, None, None, None) == -1  # NULL inputs
    assert L.trl_game_setup(None, 1, 0, 1, 0, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from tetris_reinforcement_learning_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "_SO", "/nonexistent/libtrl_b200.so")
    with pytest.raises(_native.NativeLibraryError):
        _native.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tetris_reinforcement_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "trl_oracle" not in text, f


def test_games_equal_detects_differences():
    from tetris_reinforcement_learning_b200.state import GAME_DTYPE, games_equal
    a = np.zeros(3, dtype=GAME_DTYPE)
    b = a.copy()
    b[1]["players"][0]["queue"][5] = 3      # beyond qlen: don't care
    assert games_equal(a, b).all()
    b[1]["players"][0]["qlen"] = 6          # now it matters
    assert list(games_equal(a, b)) == [True, False, True]
    b = a.copy(); b[2]["players"][1]["rows"][39] = 1
    assert list(games_equal(a, b)) == [True, True, False]


def test_policy_tables_and_move_roundtrip():
    from tetris_reinforcement_learning_b200 import const
    assert const.POLICY_SHAPE == (27, 39, 11) and const.POLICY_SIZE == 11583 and const.MASK_WORDS == 362
    assert const.policy_index_to_piece[0] == ["O", 0, 0]
    assert const.policy_index_to_piece[6] == ["I", 1, 0]
    assert const.policy_index_to_piece[10] == ["L", 3, 0]
    assert const.policy_index_to_piece[22] == ["T", 3, 1]
    assert const.policy_index_to_piece[26] == ["T", 3, 2]
    assert const.policy_piece_to_index["T"][2] == {0: 17, 1: 21, 2: 25}
    for idx in (0, 1, 428, 429, 5000, 11582):
        assert const.move_to_index(const.index_to_move(idx)) == idx
    assert const.index_to_move(0) == (0, -2, 0)
