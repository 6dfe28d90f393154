"""Short, single-purpose workloads to run under ncu (one target per invocation, few launches each).

    python tools/profile_targets.py movegen_warp|movegen_thread [n_boards]     # the sweep kernel / the scalar restatement
    python tools/profile_targets.py wide64|wide32|wide64post [n_boards]       # csrc/trunk_wide.cu alone
    python tools/profile_targets.py step16|step64|step32 [n_games]            # a few self-play steps (eager launches)

Under ncu use e.g.  ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <count> -o gpurun_out/<name>
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, architectures as arch, move_generation, synth, trunk_wide  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.const import MASK_WORDS  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")

if what.startswith("movegen"):
    n_boards = n or (100000 if what == "movegen_warp" else 10000)
    boards, cur, alt = synth.movegen_workload(n_boards)
    nc = boards.shape[0]
    d_boards = torch.from_numpy(boards.view(np.int16)).to(dev)
    d_cur, d_alt = torch.from_numpy(cur).to(dev), torch.from_numpy(alt).to(dev)
    d_mask = torch.zeros((nc, MASK_WORDS), dtype=torch.int32, device=dev)
    d_n = torch.zeros(nc, dtype=torch.int16, device=dev)
    d_st = torch.zeros(nc, dtype=torch.int32, device=dev)
    _native.lib().trl_movegen_select_kernel(0 if what == "movegen_thread" else 1)
    _native.lib().trl_movegen_warp_form({"movegen_warp2": 0, "movegen_solo": 1, "movegen_rows": 2, "movegen_rows_list": 2}.get(what, -1))
    d_moves = torch.zeros((nc, 128), dtype=torch.int16, device=dev) if what.endswith("_list") else None
    for _ in range(3):
        move_generation.movegen_device(d_boards, d_cur, d_alt, None if d_moves is not None else d_mask, d_moves, d_n, d_st)
    torch.cuda.synchronize()
    print(what, "calls", nc, "placements", int(d_n.to(torch.int64).sum()), "status", int((d_st != 0).sum()))
elif what.startswith("wide"):
    f = 64 if "64" in what else 32
    torch.manual_seed(0)
    if what.endswith("post") or f == 32:
        net = arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=20 if f == 64 else 8, filters=f))
    else:
        net = arch.AlphaSame(arch.AlphaSameConfig(blocks=20, filters=f))
    net = net.to(dev).eval()
    nb = n or 4096
    grids = (torch.rand((nb, 1, 40, 10), device=dev) < 0.35).to(torch.bfloat16)
    wt = trunk_wide.WideTrunk(trunk_wide.pack_wide_trunk(net), dev)
    out = torch.empty((nb, wt.row_elems), dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        wt(grids, out, n_images=nb)
    torch.cuda.synchronize()
    wt.check()
    print(what, "boards", nb, "ok")
else:
    from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, best_evaluator  # noqa: E402
    f = int(what[4:])
    torch.manual_seed(0)
    mc = {16: arch.AlphaSameConfig(blocks=10, filters=16), 64: arch.AlphaSameConfig(blocks=20, filters=64),
          32: arch.AuxBaseResNetConfig()}[f]
    net = arch.build_network(mc).to(dev)
    cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
                 use_playout_cap_randomization=False)
    eng = SelfPlayEngine(cfg, best_evaluator(net), n or 4096, device=dev, seed=20261018, feature_dtype=torch.bfloat16,
                         use_cuda_graph=False)
    eng.step(24)
    torch.cuda.synchronize()
    print(what, "steps 24 status", eng.status_bits())
