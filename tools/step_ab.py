"""A/B timing of self-play step organisations on BASELINE config 3 (4096 games, AlphaSame(10,16), MAX_ITER 160).

    python tools/step_ab.py [--steps 320] [--desync 0] variant ...
variant = comma separated key=value engine flags, e.g. "compact_movegen=0,fuse_expand_select=0"
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, best_evaluator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=320)
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--desync", type=int, default=0, help="warm-up steps before timing (games drift apart after ~10k)")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--net", default="alphasame", choices=["alphasame", "alphasame64", "base", "aux"],
                help="alphasame = AlphaSame(10,16) (fused trunk); alphasame64 = AlphaSame(20,64), base / aux = BaseResNet / "
                     "AuxBaseResNet(8,32) (the Config default) through the PyTorch evaluator")
ap.add_argument("--forced", action="store_true", help="forced playouts + policy-target pruning (BASELINE config 4)")
ap.add_argument("--cudnn-benchmark", action="store_true")
ap.add_argument("--nchw", action="store_true", help="PyTorch evaluator without channels-last")
ap.add_argument("variants", nargs="*", default=["", "compact_movegen=0", "fuse_expand_select=0", "compact_movegen=0,fuse_expand_select=0"])
args = ap.parse_args()

torch.manual_seed(0)
if args.net == "alphasame":
    mc = arch.AlphaSameConfig(blocks=10, filters=16); net = arch.AlphaSame(mc)
elif args.net == "alphasame64":
    mc = arch.AlphaSameConfig(blocks=20, filters=64); net = arch.AlphaSame(mc)
elif args.net == "base":
    mc = arch.BaseResNetConfig(); net = arch.BaseResNet(mc)
else:
    mc = arch.AuxBaseResNetConfig(); net = arch.AuxBaseResNet(mc)
net = net.to("cuda:0")
torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
if args.nchw:
    from tetris_reinforcement_learning_b200.selfplay import make_net_evaluator
    ev = make_net_evaluator(net, torch.bfloat16, channels_last=False)
else:
    ev = best_evaluator(net, torch.bfloat16)
cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=160, CPUCT=0.75, training=True,
             use_playout_cap_randomization=False, use_dirichlet_noise=True, FpuStrategy="reduction",
             use_forced_playouts_and_policy_target_pruning=args.forced)
print("net", args.net, "forced", args.forced, flush=True)
for v in args.variants:
    kw = {}
    for item in filter(None, v.split(",")):
        k, val = item.split("=")
        if k == "pdl":
            from tetris_reinforcement_learning_b200 import _native
            _native.lib().trl_set_pdl(int(val))
            continue
        if k == "rounds":
            from tetris_reinforcement_learning_b200 import _native
            _native.lib().trl_search_movegen_rounds(int(val))
            continue
        kw[k] = val if k == "overlap_movegen" and not val.isdigit() else (int(val) if k == "steps_per_graph" else bool(int(val)))
    eng = SelfPlayEngine(cfg, ev, args.games, seed=20261018, feature_dtype=torch.bfloat16, **kw)
    eng.step(12 + args.desync)
    eng.drain()
    best = []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(args.steps)
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / args.steps)
        eng.drain()
    st = int((eng.get_ctl()["status"] != 0).sum())
    print(f"variant [{v or 'default'}] ms/step {' '.join(f'{b:.4f}' for b in best)}  sims/s {args.games / (min(best) * 1e-3):.4g}  status_nonzero {st}", flush=True)
    del eng
    torch.cuda.empty_cache()
