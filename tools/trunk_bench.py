"""Times the fused trunk kernels (both layouts) on random boards: python tools/trunk_bench.py [n_images]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import architectures as arch, trunk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
net = arch.AlphaSame(arch.AlphaSameConfig(blocks=10, filters=16)).to("cuda:0").eval()
grids = (torch.rand((n, 1, 40, 10), device="cuda:0") < 0.3).to(torch.bfloat16)
outs = {}
for layout in ("rows", "taps"):
    packed = trunk.pack_alphasame_trunk(net, layout=layout)
    out = torch.empty((n, 400), dtype=torch.bfloat16, device="cuda:0")
    for _ in range(3):
        trunk.trunk_forward(packed, grids, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        trunk.trunk_forward(packed, grids, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flops = n * 37196800 / 2 * 2  # conv MACs per grid x 2
    print(f"{layout}: {ms * 1e3:.1f} us / {n} images  ({flops / ms / 1e9:.1f} TFLOP/s useful)")
    outs[layout] = out.float()
print("max |rows - taps| =", (outs["rows"] - outs["taps"]).abs().max().item())
