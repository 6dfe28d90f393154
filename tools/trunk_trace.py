"""Clock-stamp trace of the row-Toeplitz trunk kernel (CTA 0, first group): python tools/trunk_trace.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import _native, architectures as arch, trunk  # noqa: E402

n = 8192
torch.manual_seed(0)
net = arch.AlphaSame(arch.AlphaSameConfig(blocks=10, filters=16)).to("cuda:0").eval()
grids = (torch.rand((n, 1, 40, 10), device="cuda:0") < 0.3).to(torch.bfloat16)
packed = trunk.pack_alphasame_trunk(net, layout="rows")
L = _native.lib()
trunk.trunk_forward(packed, grids)
buf = torch.zeros(21 * 10 * 4, dtype=torch.int64, device="cuda:0")
L.trl_debug_trunk_rows_trace.argtypes = [ctypes.c_void_p]
L.trl_debug_trunk_rows_trace(buf.data_ptr())
trunk.trunk_forward(packed, grids)
torch.cuda.synchronize()
L.trl_debug_trunk_rows_trace(None)
t = buf.cpu().view(21, 10, 4)
t0 = int(t[0, 0, 0])
print("layer col :  issue  epi_wake  ld_done  published   (clk since first issue)")
for layer in (0, 1, 2, 11, 12):
    for c in range(0, 10, 2):
        r = [int(v) - t0 if int(v) else -1 for v in t[layer, c]]
        q = [int(v) - t0 if int(v) else -1 for v in t[layer, c + 1]]
        print(f"{layer:3d} {c:2d} : issue {r[0]:7d} (c+1 {q[0]:7d})  wake {r[1]:7d}  col0 +{r[2]-r[1]:5d}  col1 +{q[1]-r[2]:5d}  "
              f"wait_st +{q[2]-q[1]:5d}  proxyfence +{q[3]-q[2]:5d}  publish +{r[3]-q[3]:5d}   epi total {r[3]-r[1]:6d}")
print("clk per layer (issue col0 L -> L+1):", [int(t[l + 1, 0, 0] - t[l, 0, 0]) for l in range(1, 20)])
