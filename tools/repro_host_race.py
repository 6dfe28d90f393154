"""Repro of a race found in round 2 (fixed in capi.cu: cudaMemset on the legacy stream vs. the first H2D copy on a non-blocking stream after the host workspace is re-allocated): host and compact entry points against the oracle, all launch forms."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle
from tetris_reinforcement_learning_b200 import _native, move_generation as mg, synth
L=_native.lib()
boards, cur, alt = synth.movegen_workload(43000, seed=5, caves=True)
cur = cur.copy(); alt = alt.copy()
cur[::97] = 255; alt[::97] = 255
_, want_n, _, _ = oracle.movegen_batch(boards, cur, alt, n_threads=os.cpu_count() or 1)
for form in (0, 1, 2, 0, 2):
    L.trl_movegen_select_kernel(1); L.trl_movegen_warp_form(form)
    for rep in range(4):
        ref = mg.movegen_host(boards, cur, alt, want_mask=True, want_moves=False)
        res = mg.movegen_host_compact(boards, cur, alt)
        b1 = np.flatnonzero(ref["n_moves"] != want_n); b2 = np.flatnonzero(res["n_moves"] != want_n)
        print("form", form, "rep", rep, "host bad", b1.size, b1[:6], "compact bad", b2.size, b2[:6], flush=True)
