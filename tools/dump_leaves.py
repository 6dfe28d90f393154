"""Dump the leaf states of one self-play step (for offline analysis of the leaf movegen workload)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch, trunk  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine  # noqa: E402
from tetris_reinforcement_learning_b200.state import GAME_DTYPE  # noqa: E402

torch.manual_seed(0)
mc = arch.AlphaSameConfig(blocks=10, filters=16)
net = arch.AlphaSame(mc).to("cuda:0")
cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
             use_playout_cap_randomization=False)
eng = SelfPlayEngine(cfg, trunk.make_fused_evaluator(net), 4096, seed=20261018, feature_dtype=torch.bfloat16)
eng.step(int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
torch.cuda.synchronize()
ls = eng.t["leaf_state"].cpu().numpy()
states = eng.t["states"].view(-1, 400)
sel = torch.from_numpy(ls[ls >= 0].astype(np.int64)).to("cuda:0")
leaves = states[sel].cpu().numpy().view(GAME_DTYPE).reshape(-1)
np.save(os.path.join("gpurun_out", "leaf_states.npy"), leaves.view(np.uint8).reshape(len(leaves), -1))
print("dumped", len(leaves), "leaf states")
