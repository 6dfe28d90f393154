"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py [movegen|trunk|wide|engine|wide_engine]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch, move_generation, synth, trunk, trunk_wide  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("movegen", "all"):
    from tetris_reinforcement_learning_b200 import _native
    boards, cur, alt = synth.movegen_workload(40, seed=3, caves=True)
    # every form of the warp-cooperative enumeration (two warps per call / one warp per call / closure-search
    # kernel + clean-up pass), then every search through the exact FIFO form, then the one-thread kernel
    for kernel, form, fast in ((1, 0, 1), (1, 1, 1), (1, 2, 1), (1, 0, 0), (0, -1, 1)):
        _native.lib().trl_movegen_select_kernel(kernel)
        _native.lib().trl_movegen_warp_form(form)
        _native.lib().trl_debug_movegen_fast_path(fast)
        res = move_generation.movegen_host(boards, cur, alt, want_mask=True, want_moves=True)
        res2 = move_generation.movegen_host_compact(boards, cur, alt)
        print("movegen ok", (kernel, form, fast), int(res["n_moves"].sum()), res2["total"])
    _native.lib().trl_movegen_select_kernel(-1)
    _native.lib().trl_movegen_warp_form(-1)
    _native.lib().trl_debug_movegen_fast_path(1)
if what in ("trunk", "all"):
    torch.manual_seed(0)
    net = arch.AlphaSame(arch.AlphaSameConfig(blocks=2)).to("cuda:0").eval()
    packed = trunk.pack_alphasame_trunk(net)
    g = (torch.rand((11, 1, 40, 10), device="cuda:0") < 0.3).to(torch.bfloat16)
    out = trunk.trunk_forward(packed, g)
    torch.cuda.synchronize()
    print("trunk ok", float(out.float().abs().sum()))
if what in ("engine", "all"):
    torch.manual_seed(0)
    mc = arch.AlphaSameConfig(blocks=2)
    net = arch.AlphaSame(mc).to("cuda:0")
    cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=4, training=True)
    eng = SelfPlayEngine(cfg, trunk.make_fused_evaluator(net), 16, seed=1, feature_dtype=torch.bfloat16, max_rounds=3,
                         use_cuda_graph=False)
    eng.step(40)
    s, e = eng.drain()
    print("engine ok", len(s), len(e), int(eng.get_ctl()["status"].max()))
if what in ("wide", "all"):
    for net_cfg in (arch.AlphaSameConfig(blocks=1, filters=64), arch.AuxBaseResNetConfig(blocks=1, filters=32)):
        torch.manual_seed(0)
        net = arch.build_network(net_cfg).to("cuda:0").eval()
        wt = trunk_wide.WideTrunk(trunk_wide.pack_wide_trunk(net), "cuda:0")
        g = (torch.rand((11, 1, 40, 10), device="cuda:0") < 0.3).to(torch.bfloat16)
        out = trunk_wide.wide_trunk_forward(wt, g)
        torch.cuda.synchronize()
        wt.check()
        print("wide trunk ok", type(net).__name__, float(out.float().abs().sum()))
if what in ("wide_engine", "all"):
    from tetris_reinforcement_learning_b200.selfplay import best_evaluator  # noqa: E402
    torch.manual_seed(0)
    mc = arch.AuxBaseResNetConfig(blocks=1, filters=32)
    net = arch.build_network(mc).to("cuda:0")
    cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=4, training=True)
    eng = SelfPlayEngine(cfg, best_evaluator(net), 16, seed=1, feature_dtype=torch.bfloat16, max_rounds=3, use_cuda_graph=False)
    eng.step(40)
    s, e = eng.drain()
    print("wide engine ok", len(s), len(e), int(eng.get_ctl()["status"].max()))
