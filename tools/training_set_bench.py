"""End-to-end throughput of the drop-in data-generation call (ai.generate_games = the body of make_training_set):
complete games at BASELINE config 3 through the public API, records drained to the host and assembled into a data set.
compact format at full size, the reference's JSON sample lists on a small number of games."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import ai  # noqa: E402
from tetris_reinforcement_learning_b200 import architectures as arch  # noqa: E402

os.environ.setdefault("TRL_STORAGE", "/tmp/trl_storage_bench")
torch.manual_seed(0)
mc = arch.AlphaSameConfig(blocks=10, filters=16)
cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
                use_playout_cap_randomization=False, use_dirichlet_noise=True, FpuStrategy="reduction")
os.makedirs(cfg.model_dir, exist_ok=True)
os.makedirs(cfg.data_dir, exist_ok=True)
net = arch.AlphaSame(mc).to("cuda:0").eval()
for fmt, games in (("compact", 4096), ("compact", 4096), ("json", 64)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    data, stats = ai.generate_games(cfg, net, games, seed=20261018, compact=(fmt == "compact"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = len(data)
    print(f"{fmt:8s} {games} games: {dt:6.2f} s wall -> {games / dt * 3600:.3g} games/h, {n} samples ({n / dt:.3g} samples/s)", flush=True)
