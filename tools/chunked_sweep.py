import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from tetris_reinforcement_learning_b200 import _native, move_generation, synth
from tetris_reinforcement_learning_b200.const import MASK_WORDS
dev=torch.device("cuda:0")
boards,cur,alt=synth.movegen_workload(1000000)
n=boards.shape[0]
d_b=torch.from_numpy(boards.view(np.int16)).to(dev); d_c=torch.from_numpy(cur).to(dev); d_a=torch.from_numpy(alt).to(dev)
d_n=torch.zeros(n,dtype=torch.int16,device=dev); d_st=torch.zeros(n,dtype=torch.int32,device=dev)
d_moves=torch.zeros((n,128),dtype=torch.int16,device=dev)
def run(chunk):
    for lo in range(0,n,chunk):
        hi=min(n,lo+chunk)
        move_generation.movegen_device(d_b[lo:hi],d_c[lo:hi],d_a[lo:hi],None,d_moves[lo:hi],d_n[lo:hi],d_st[lo:hi])
for chunk in (n, 1<<19, 1<<17, 1<<16):
    run(chunk); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run(chunk)
    e1.record(); torch.cuda.synchronize()
    print("chunk",chunk,"ms per sweep (lists, moves_cap 128): %.2f"%(e0.elapsed_time(e1)/3))
