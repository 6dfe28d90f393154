#!/usr/bin/env python
"""BASELINE config 5 on N GPUs of one box: larger net, self-play -> train -> gate with the NCCL
gradient all-reduce (the only collective of the whole system; the reference has none).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        tools/config5_loop.py [--blocks 20 --filters 64 --max-iter 800 --games 256 --battle-games 64]

Every rank plays its own shard of self-play games (ids rank, rank+N, ...: no collective on the data
path), trains on its own samples with gradients averaged over NVLink (training.allreduce_gradients),
and plays its shard of the gating battles; win counts are summed with one all-reduce.  Rank 0 prints
one JSON line."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_reinforcement_learning_b200 import ai, training  # noqa: E402
from tetris_reinforcement_learning_b200 import architectures as arch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--filters", type=int, default=64)
    ap.add_argument("--max-iter", type=int, default=800)
    ap.add_argument("--games", type=int, default=256, help="self-play games per GPU")
    ap.add_argument("--battle-games", type=int, default=64, help="gating games per GPU")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--train-samples", type=int, default=0, help="cap on the samples per rank used for the training step (0 = all)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.setdefault("TRL_STORAGE", f"/tmp/trl_config5_{os.getpid()}")
    mc = arch.AlphaSameConfig(blocks=args.blocks, filters=args.filters)
    cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=args.max_iter, CPUCT=0.75,
                    training=True, batch_size=args.batch_size, epochs=1)
    torch.manual_seed(0)                                  # identical initial weights on every rank
    best = arch.AlphaSame(mc).cuda().eval()
    challenger = arch.AlphaSame(mc).cuda()
    challenger.load_state_dict(best.state_dict())

    # ---- self-play: independent shards ----
    t0 = time.perf_counter()
    data, stats = ai.generate_games(cfg, best, args.games, seed=20261018, first_game_id=rank, game_id_stride=world, compact=True)
    torch.cuda.synchronize()
    t_play = time.perf_counter() - t0

    # ---- training: same number of optimizer steps on every rank, gradients averaged over NCCL ----
    n = torch.tensor([len(data)], device="cuda")
    if world > 1:
        dist.all_reduce(n, op=dist.ReduceOp.MIN)
    n_common = int(n.item()) // args.batch_size * args.batch_size
    if args.train_samples:
        n_common = min(n_common, args.train_samples // args.batch_size * args.batch_size)
    mem_play = torch.cuda.max_memory_allocated()
    t0 = time.perf_counter()
    import numpy as np
    losses = training.train_network_pytorch(cfg, challenger, training._CompactSelection(data, np.arange(n_common)), log=False) if n_common else {}
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    flat = torch.cat([p.detach().reshape(-1).float() for p in challenger.parameters()])
    spread = torch.zeros(1, device="cuda")
    if world > 1:                                         # replicas must stay bit-identical
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = (hi - lo).abs().max().reshape(1)

    # ---- gating: the battles shard like self-play ----
    gate_cfg = cfg.copy()
    gate_cfg.training = False
    t0 = time.perf_counter()
    wins, _ = training.battle_networks(challenger.eval(), gate_cfg, best, gate_cfg, None, "moreorequal", args.battle_games,
                                       seed=7, first_game_id=rank, game_id_stride=world)
    t_gate = time.perf_counter() - t0
    w = torch.tensor(wins, device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(len(data)), float(args.games), t_play, t_train, t_gate], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(w)
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm)
        samples, games, t_play, t_train, t_gate = float(sm[0]), float(sm[1]), float(mx[2]), float(mx[3]), float(mx[4])
    else:
        samples, games = float(tot[0]), float(tot[1])
    total_games = args.battle_games * world
    accepted = training._check_threshold(w.cpu().numpy(), total_games, cfg.gating_threshold, cfg.gating_threshold_type)
    if rank == 0:
        print(json.dumps({
            "config": f"AlphaSame(blocks={args.blocks}, filters={args.filters}) MAX_ITER={args.max_iter}, {world} GPU(s)",
            "selfplay_games": games, "samples": samples, "selfplay_s": t_play, "games_per_hour": games / t_play * 3600,
            "train_samples_per_rank": n_common, "train_s": t_train, "losses": losses,
            "replica_weight_spread_after_allreduce_training": float(spread.item()),
            "gating": {"games": total_games, "challenger_wins": float(w[0]), "best_wins": float(w[1]), "accepted": accepted,
                       "seconds": t_gate},
            "params": int(flat.numel()),
            "hbm_peak_bytes_selfplay_rank0": int(mem_play), "hbm_peak_bytes_total_rank0": int(torch.cuda.max_memory_allocated())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
