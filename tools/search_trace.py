"""Phase trace of search_expand_select_encode_kernel (needs a build with TRL_NVCC_EXTRA=-DTRL_SEARCH_TRACE):
mean clocks between the stamps of lane 0 of every game's warp, over a few steps of BASELINE config 3."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, architectures as arch  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, best_evaluator  # noqa: E402

NAMES = ["start", "ctl+path loaded", "legal list located", "logits gathered", "exp + prior store", "children created", "backup",
         "FPU refresh", "(end of search)", "selection walk", "parent state staged", "env step", "leaf classified", "features encoded"]
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
mc = arch.AlphaSameConfig(blocks=10, filters=16)
net = arch.AlphaSame(mc).to("cuda:0")
cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
             use_playout_cap_randomization=False)
eng = SelfPlayEngine(cfg, best_evaluator(net), G, device=torch.device("cuda:0"), seed=20261018, feature_dtype=torch.bfloat16,
                     use_cuda_graph=False)
eng.step(100)          # mid-search: trees have depth
buf = torch.zeros((G, 16), dtype=torch.int64, device="cuda:0")
L = _native.lib()
L.trl_debug_search_trace.argtypes = [ctypes.c_void_p]
assert L.trl_debug_search_trace(buf.data_ptr()) == 0, "build with TRL_NVCC_EXTRA=-DTRL_SEARCH_TRACE"
acc = np.zeros(14)
n = 0
for _ in range(8):
    buf.zero_()
    eng.step(1)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype(np.float64)
    ok = (t[:, :14] > 0).all(axis=1)
    d = np.diff(t[ok][:, :14], axis=1)
    acc[1:] += d.mean(axis=0)
    acc[0] += (t[ok][:, 13] - t[ok][:, 0]).mean()
    n += 1
L.trl_debug_search_trace(None)
print(f"games {G}: mean clocks per game from kernel start to end: {acc[0] / n:.0f}")
for k in range(1, 14):
    print(f"  {NAMES[k]:24s} {acc[k] / n:8.0f}")
