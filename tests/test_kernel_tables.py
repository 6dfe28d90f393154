"""CPU checks of the compile-time tables in csrc/movegen_warp.cu: the closure kernel's specialised kick passes and
validity rows use constexpr COPIES of the constant-memory tables of csrc/trl_tables.cuh (reference const.py:191-281);
the copies must stay identical to the originals, which the oracle / golden tests pin against the reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tetris_reinforcement_learning_b200", "csrc")


def _kick_lists(text):
    out = []
    for m in re.finditer(r"\{(\d), \{((?:\{-?\d+, -?\d+\},? ?)+)\}\}", text):
        n = int(m.group(1))
        pairs = [(int(a), int(b)) for a, b in re.findall(r"\{(-?\d+), (-?\d+)\}", m.group(2))]
        out.append((n, pairs[:n]))
    return out


def _minos(text):
    return [tuple(int(x) for x in m.groups()) for m in re.finditer(
        r"TRL_PK\((\d), (\d), (\d), (\d), (\d), (\d), (\d), (\d)\)", text)]


def test_constexpr_kick_tables_equal_the_constant_memory_tables():
    tables = open(os.path.join(CSRC, "trl_tables.cuh")).read()
    warp = open(os.path.join(CSRC, "movegen_warp.cu")).read()
    ref = _kick_lists(tables[tables.index("c_kicks[2][4][3]"):tables.index("// policy planes")])
    assert len(ref) == 24
    wall = _kick_lists(warp[warp.index("constexpr KickList kWallKicks"):warp.index("constexpr KickList kIKicks")])
    ikick = _kick_lists(warp[warp.index("constexpr KickList kIKicks"):warp.index("template <int TAB, int R, int KD, class St>")])
    assert wall == ref[:12]
    assert ikick == ref[12:]
    for n, pairs in ref:
        assert pairs[0] == (0, 0)      # the closure search tests the in-place kick from registers


def test_constexpr_minos_equal_the_constant_memory_table():
    tables = open(os.path.join(CSRC, "trl_tables.cuh")).read()
    warp = open(os.path.join(CSRC, "movegen_warp.cu")).read()
    ref = _minos(tables[tables.index("c_minos[7][4]"):tables.index("// Kick lists")])
    got = _minos(warp[warp.index("constexpr uint32_t kMinos[7][4]"):warp.index("template <int TYPE, int R>")])
    assert len(ref) == 28 and got == ref
