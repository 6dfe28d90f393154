"""Times the wide fused trunk (csrc/trunk_wide.cu) against the same trunk through PyTorch / cuDNN (bf16).

    python tools/trunk_wide_bench.py [n_boards]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tetris_reinforcement_learning_b200 import architectures as arch, trunk_wide


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = "cuda:0"
    cases = [("alphasame", 20, 64), ("aux", 8, 32), ("alphasame", 10, 32), ("base", 20, 64)]
    for family, blocks, f in cases:
        torch.manual_seed(0)
        if family == "alphasame":
            net = arch.AlphaSame(arch.AlphaSameConfig(blocks=blocks, filters=f))
            ref = lambda g, net=net: net.grid_features(g)
            stem = 25
        else:
            net = (arch.AuxBaseResNet(arch.AuxBaseResNetConfig(blocks=blocks, filters=f)) if family == "aux"
                   else arch.BaseResNet(arch.BaseResNetConfig(blocks=blocks, filters=f)))
            ref = lambda g, net=net: net._process_grid(g)
            stem = 9
        net = net.to(dev).eval()
        flops = 2.0 * n * 400 * (stem * f + 2 * blocks * 9 * f * f)
        grids = (torch.rand((n, 1, 40, 10), device=dev) < 0.35).to(torch.bfloat16)
        wt = trunk_wide.WideTrunk(trunk_wide.pack_wide_trunk(net), dev)
        out = torch.empty((n, wt.row_elems), dtype=torch.bfloat16, device=dev)
        from tetris_reinforcement_learning_b200 import _native
        _native.lib().trl_debug_trunk_wide_l2_window(0)           # A/B: without the persisting-L2 window over the scratch
        best0, _ = timed(lambda: wt(grids, out, n_images=n))
        _native.lib().trl_debug_trunk_wide_l2_window(1)
        best, med = timed(lambda: wt(grids, out, n_images=n))
        wt.check()
        net16 = net.to(torch.bfloat16)
        cl = f >= 64
        g2 = grids.contiguous(memory_format=torch.channels_last) if cl else grids
        if cl:
            net16 = net16.to(memory_format=torch.channels_last)
        with torch.no_grad():
            tb, tm = timed(lambda: ref(g2), reps=3, warm=1)
        print(json.dumps({"net": f"{family}({blocks},{f})", "boards": n, "fused_ms_best": round(best, 3), "fused_ms_median": round(med, 3), "fused_ms_best_without_l2_window": round(best0, 3),
                          "fused_tflops": round(flops / best / 1e9, 1), "cudnn_ms_best": round(tb, 3), "speedup": round(tb / best, 1)}), flush=True)


if __name__ == "__main__":
    main()
