"""Times both movegen kernels (device-resident inputs): python tools/movegen_bench.py [n_boards]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, move_generation, synth  # noqa: E402
from tetris_reinforcement_learning_b200.const import MASK_WORDS  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
L = _native.lib()
dev = torch.device("cuda:0")
for n_boards in (586, nb):   # 586 x 7 = 4102 calls: the size of one self-play step
    boards, cur, alt = synth.movegen_workload(n_boards)
    n = boards.shape[0]
    d_boards = torch.from_numpy(boards.view(np.int16)).to(dev)
    d_cur, d_alt = torch.from_numpy(cur).to(dev), torch.from_numpy(alt).to(dev)
    res = {}
    for name, k, form in (("thread", 0, -1), ("warp", 1, 0), ("solo", 1, 1), ("rows", 1, 2)):
        L.trl_movegen_select_kernel(k)
        L.trl_movegen_warp_form(form)
        for want_mask in (True, False):
            d_mask = torch.zeros((n, MASK_WORDS), dtype=torch.int32, device=dev) if want_mask else None
            d_moves = None if want_mask else torch.zeros((n, 512), dtype=torch.int16, device=dev)
            d_n = torch.zeros(n, dtype=torch.int16, device=dev)
            d_st = torch.zeros(n, dtype=torch.int32, device=dev)
            for _ in range(2):
                move_generation.movegen_device(d_boards, d_cur, d_alt, d_mask, d_moves, d_n, d_st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                move_generation.movegen_device(d_boards, d_cur, d_alt, d_mask, d_moves, d_n, d_st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            tot = int(d_n.to(torch.int64).sum())
            if k == 1:
                import ctypes
                st2 = (ctypes.c_uint64 * 8)()
                L.trl_debug_movegen_fast_stats(st2)
                print(f"    closure form answered {st2[0]} piece searches, FIFO form {st2[1]} ({100.0 * st2[1] / max(st2[0] + st2[1], 1):.3f} %)")
            print(f"n_calls={n:8d} kernel={name:6s} out={'mask' if want_mask else 'list'}: {ms * 1e3:9.1f} us  "
                  f"{tot / ms / 1e6:8.1f} G placements/s  status!=0: {int((d_st != 0).sum())}")
            res[(name, want_mask)] = (d_n.clone(), None if d_mask is None else d_mask.clone())
    for other in ("warp", "solo", "rows"):
        print(f"  thread vs {other}: counts equal:", bool((res[("thread", True)][0] == res[(other, True)][0]).all()),
              " masks equal:", bool((res[("thread", True)][1] == res[(other, True)][1]).all()))
L.trl_movegen_select_kernel(-1)
L.trl_movegen_warp_form(-1)
