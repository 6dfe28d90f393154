"""In-graph timeline of a self-play step (BASELINE config 3) from %globaltimer stamps between the kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, best_evaluator  # noqa: E402

torch.manual_seed(0)
mc = arch.AlphaSameConfig(blocks=10, filters=16)
net = arch.AlphaSame(mc).to("cuda:0")
ev = best_evaluator(net, torch.bfloat16)
cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
             use_playout_cap_randomization=False, use_dirichlet_noise=True, FpuStrategy="reduction")
mode = sys.argv[1] if len(sys.argv) > 1 else "trunk"
print("overlap_movegen =", mode)
if mode == "serial":
    mode = False
eng = SelfPlayEngine(cfg, ev, 4096, seed=20261018, feature_dtype=torch.bfloat16, overlap_movegen=mode)
eng.enable_timeline()
eng.step(40)
rows = []
for _ in range(200):
    eng.step(1)
    torch.cuda.synchronize()
    rows.append(eng._stamps.cpu().numpy().copy())
r = np.array(rows).astype(np.float64)
it = (np.arange(len(r)) + 40) % 160
keep = (it > 3)           # skip the first iterations of a search (all leaves enumerate, roots take two boards)
t0 = r[:, 0:1]
names = {1: "select (0 when fused)", 2: "encode", 3: "trunk", 4: "heads", 5: "policy GEMM", 6: "expand(+select)"}
prev = 0
print("mean over", int(keep.sum()), "steps, us after step start / duration")
for k in (1, 2, 3, 4, 5, 6):
    print(f"  {names[k]:24s} end {np.mean((r[keep, k] - r[keep, 0]) / 1e3):8.1f}   dur {np.mean((r[keep, k] - r[keep, prev]) / 1e3):8.1f}")
    prev = k
print(f"  movegen (forked)         start {np.mean((r[keep, 8] - r[keep, 0]) / 1e3):8.1f} end {np.mean((r[keep, 9] - r[keep, 0]) / 1e3):8.1f}")
cnt = None
print("step-to-step (stamp 0 to next stamp 0):", np.mean(np.diff(r[:, 0])[keep[1:]]) / 1e3, "us (includes the host sync between replays)")
