"""One-off soak of the T emission-order analysis: T searches (current and hold both T-heavy) on fresh seeds of all four
board families, closure kernel + clean-up (form 2) and the one-kernel form against the one-thread-per-call kernel
(literal FIFO).  python tools/t_soak.py [n_boards] [seed ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, move_generation, synth  # noqa: E402
from tetris_reinforcement_learning_b200.const import MASK_WORDS  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
seeds = [int(x) for x in sys.argv[2:]] or [101, 202, 303]
L = _native.lib()
dev = torch.device("cuda:0")
bad_total = 0
for seed in seeds:
    boards, cur, alt = synth.movegen_workload(nb, seed=seed, caves=True)
    cur = cur.copy(); alt = alt.copy()
    cur[::2] = 6                       # every second call searches T as the current piece
    alt[1::3] = 6                      # ... and every third holds a T
    n = cur.size
    d_b = torch.from_numpy(boards.view(np.int16)).to(dev)
    d_c, d_a = torch.from_numpy(cur).to(dev), torch.from_numpy(alt).to(dev)
    outs = []
    for kernel, form in ((0, -1), (1, 2), (1, 1)):
        L.trl_movegen_select_kernel(kernel); L.trl_movegen_warp_form(form)
        mask = torch.zeros((n, MASK_WORDS), dtype=torch.int32, device=dev)
        cnt = torch.zeros(n, dtype=torch.int16, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev)
        move_generation.movegen_device(d_b, d_c, d_a, mask, None, cnt, st)
        torch.cuda.synchronize()
        outs.append((mask, cnt, st))
    for name, o in (("closure + clean-up", outs[1]), ("one kernel", outs[2])):
        diff = int((outs[0][0] != o[0]).any(dim=1).sum()) + int((outs[0][1] != o[1]).sum()) + int((o[2] != 0).sum())
        bad_total += diff
        print(f"seed {seed}: {n} calls, {name} vs literal FIFO: {diff} differences", flush=True)
L.trl_movegen_select_kernel(-1); L.trl_movegen_warp_form(-1)
print("TOTAL differences", bad_total)
sys.exit(1 if bad_total else 0)
