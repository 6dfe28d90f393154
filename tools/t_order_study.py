"""Research tool (build with TRL_NVCC_EXTRA=-DTRL_T_TRACE): which arrivals decide the T-spin flag of cells that
receive both flag values?  Runs T-only calls, fetches the per-(round, target rotation, direction) arrival planes of
the closure search and compares candidate ordering rules with the exact answer (the FIFO form's mask)."""
import collections
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, move_generation, synth  # noqa: E402
from tetris_reinforcement_learning_b200.const import MASK_WORDS  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
L = _native.lib()
dev = torch.device("cuda:0")
boards, cur, alt = synth.movegen_workload(nb)
boards = boards[::7].copy()
n = boards.shape[0]
cur = np.full(n, 6, np.uint8)
d_b = torch.from_numpy(boards.view(np.int16)).to(dev)
d_c = torch.from_numpy(cur).to(dev)
d_mask = torch.zeros((n, MASK_WORDS), dtype=torch.int32, device=dev)
d_n = torch.zeros(n, dtype=torch.int16, device=dev)
L.trl_movegen_select_kernel(1); L.trl_movegen_warp_form(1)
buf = torch.zeros((n + 16, 8 * 4 * 3 * 32 + 4 * 32 + 32), dtype=torch.int32, device=dev)
L.trl_debug_movegen_t_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint]
words = L.trl_debug_movegen_t_trace(buf.data_ptr(), n + 16)
assert words == buf.shape[1], words
move_generation.movegen_device(d_b, d_c, d_c, d_mask, None, d_n, None)
torch.cuda.synchronize()
L.trl_debug_movegen_t_trace(None, 0)
tr = buf.cpu().numpy().view(np.uint32)
masks = d_mask.cpu().numpy().view(np.uint32)
bits = np.unpackbits(masks.view(np.uint8), axis=1, bitorder="little")[:, :27 * 39 * 11].reshape(n, 27, 39, 11)
by_board = {boards[i].tobytes(): i for i in range(n)}
arr = tr[:, :8 * 4 * 3 * 32].reshape(-1, 8, 4, 3, 32)       # [search][round][target rot][kd][lane]
placed = tr[:, 8 * 4 * 3 * 32: 8 * 4 * 3 * 32 + 128].reshape(-1, 4, 32)
brd = tr[:, 8 * 4 * 3 * 32 + 128:]
stats = collections.Counter()
rules = collections.Counter()
n_mixed_cells = 0
for s in range(n):
    rows = np.zeros(40, np.uint16)
    rows[:32] = brd[s] & 0xFFFF
    rows[32:] = (brd[s][:8] >> 16) & 0xFFFF
    i = by_board.get(rows.tobytes())
    if i is None:
        continue
    N = arr[s] & 0xFFFF
    U = arr[s] >> 16
    anyN = np.bitwise_or.reduce(N, axis=(0, 2))       # [rot][lane]
    anyU = np.bitwise_or.reduce(U, axis=(0, 2))
    mixed = anyN & anyU & placed[s]
    if not mixed.any():
        continue
    stats["searches with mixed cells"] += 1
    for rot in range(4):
        for lane in range(32):
            m = int(mixed[rot, lane])
            while m:
                b = (m & -m).bit_length() - 1
                m &= m - 1
                n_mixed_cells += 1
                y, x = lane + 12 - 2, b               # policy row, policy column
                truth_u = bool(bits[i, 23 + rot, y, x])
                assert truth_u or bits[i, 19 + rot, y, x], "mixed cell must be flagged"
                ev = [(r, kd, "U") for r in range(8) for kd in range(3) if (U[r, rot, kd, lane] >> b) & 1]
                ev += [(r, kd, "N") for r in range(8) for kd in range(3) if (N[r, rot, kd, lane] >> b) & 1]
                ru = sorted({e[0] for e in ev if e[2] == "U"}); rn = sorted({e[0] for e in ev if e[2] == "N"})
                stats[f"rounds U {ru} N {rn} -> {'U' if truth_u else 'N'}"] += 1
                # rule A: the kind with the later last round wins (tie: undecided)
                if ru[-1] != rn[-1]:
                    rules["A decided"] += 1
                    rules["A correct"] += int((ru[-1] > rn[-1]) == truth_u)
                else:
                    # same last round: compare the source rotation of the last-round arrivals (kd 0 -> source rot-1, kd 1 -> rot-2, kd 2 -> rot-3)
                    last = ru[-1]
                    su = sorted({(rot - kd - 1) % 4 for (r, kd, k) in ev if r == last and k == "U"})
                    sn = sorted({(rot - kd - 1) % 4 for (r, kd, k) in ev if r == last and k == "N"})
                    stats[f"  same last round {last}: target rot {rot} U from rot {su} N from rot {sn} -> {'U' if truth_u else 'N'}"] += 1
print("T searches", n, "mixed cells", n_mixed_cells)
for k, v in sorted(stats.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{v:7d}  {k}")
print(dict(rules))
