#!/usr/bin/env python
"""Time the self-play step and its parts on one GPU (development probe, not the bench)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch  # noqa: E402
from tetris_reinforcement_learning_b200.config import Config  # noqa: E402
from tetris_reinforcement_learning_b200.selfplay import SelfPlayEngine, make_net_evaluator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=160)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--blocks", type=int, default=10)
    ap.add_argument("--filters", type=int, default=16)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--overlap", default="trunk")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    mc = arch.AlphaSameConfig(blocks=args.blocks, filters=args.filters)
    net = arch.AlphaSame(mc).to(dev)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    cfg = Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=args.iters, CPUCT=0.75,
                 training=True, use_playout_cap_randomization=False)
    if args.fused:
        from tetris_reinforcement_learning_b200 import trunk
        ev = trunk.make_fused_evaluator(net)
        g2 = torch.zeros((2 * args.games, 1, 40, 10), dtype=dt, device=dev)
        for _ in range(3):
            trunk.trunk_forward(ev.packed, g2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            trunk.trunk_forward(ev.packed, g2)
        b.record(); torch.cuda.synchronize()
        tms = a.elapsed_time(b) / 20
        print(f"fused trunk alone: {tms:.3f} ms for {2 * args.games} images -> {2 * args.games * 37.2e6 / tms / 1e9:.1f} TFLOP/s")
    else:
        ev = make_net_evaluator(net, dt)
    # net alone
    G = args.games
    grids = torch.zeros((2 * G, 1, 40, 10), dtype=dt, device=dev)
    extras = torch.zeros((G, 105), dtype=dt, device=dev)
    with torch.no_grad():
        for _ in range(3):
            ev(grids, extras)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ev(grids, extras)
        e1.record(); torch.cuda.synchronize()
    net_ms = e0.elapsed_time(e1) / 10
    print(f"net forward (eager, {args.dtype}, batch {G}): {net_ms:.3f} ms  -> {G / net_ms * 1e3:.3e} evals/s, "
          f"{86.5e6 * G / net_ms / 1e9:.1f} TFLOP/s")
    for graph in (False, True):
        eng = SelfPlayEngine(cfg, ev, G, seed=1, feature_dtype=dt, use_cuda_graph=graph, overlap_movegen=False if args.no_overlap else args.overlap)
        eng.step(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        eng.step(args.steps)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        samples, ends = eng.drain()
        ctl = eng.get_ctl()
        print(f"engine step (graph={graph}): {ms:.3f} ms device, {wall:.3f} ms wall -> {G / ms * 1e3:.3e} sims/s; "
              f"samples {len(samples)}, game ends {len(ends)}, status {int(ctl['status'].max())}, "
              f"max nodes {int(ctl['n_nodes'].max())}, max depth {int(ctl['max_depth'].max())}")
    # parts, eager with events
    import ctypes
    from tetris_reinforcement_learning_b200 import _native
    eng = SelfPlayEngine(cfg, ev, G, seed=1, feature_dtype=dt, use_cuda_graph=False)
    eng.step(60)
    lib, st = eng.lib, torch.cuda.current_stream().cuda_stream
    bp, pp = ctypes.byref(eng.buf), ctypes.byref(eng.params)
    names = ["select", "movegen", "features", "net", "expand"]
    acc = dict.fromkeys(names, 0.0)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    reps = 40
    for _ in range(reps):
        evs[0].record()
        lib.trl_search_select(bp, pp, st); evs[1].record()
        lib.trl_search_movegen(bp, st); evs[2].record()
        lib.trl_encode_features(eng.t["states"].data_ptr(), eng.t["leaf_state"].data_ptr(), G, eng.grids.data_ptr(),
                                eng.extras.data_ptr(), 0 if dt == torch.float32 else 1, st); evs[3].record()
        with torch.no_grad():
            v, l = ev(eng.grids, eng.extras)
        v = v.reshape(-1).to(l.dtype); evs[4].record()
        lib.trl_search_expand(bp, pp, v.data_ptr(), l.data_ptr(), l.stride(0), 0 if l.dtype == torch.float32 else 1, st); evs[5].record()
        torch.cuda.synchronize()
        for i, nme in enumerate(names):
            acc[nme] += evs[i].elapsed_time(evs[i + 1])
    print("parts (ms):", {k: round(v / reps, 4) for k, v in acc.items()})


if __name__ == "__main__":
    main()
