"""2-GPU NCCL test of BASELINE config 5's only collective (skipped with fewer than two GPUs; run it with
`gpurun --gpus 2`): data-parallel training step with the flat gradient all-reduce -> bit-identical replicas
(parameters AND BatchNorm buffers), and gating battles sharded over ranks -> the summed win counts equal the
one-rank run.  Reference: the training step ai.py:1139-1220 / 1871-1921 and the gating loop ai.py:1975-2114; the
all-reduce itself is new (the reference is single process)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      TRL_STORAGE=f"/tmp/trl_dist_test_{os.getpid()}")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        from test_gpu_search import fake_evaluator_torch
        from tetris_reinforcement_learning_b200 import ai, training
        from tetris_reinforcement_learning_b200 import architectures as arch
        mc = arch.AuxBaseResNetConfig(blocks=1, filters=32)
        cfg = ai.Config(visual=False, ruleset="s2", model="pytorch", model_config=mc, MAX_ITER=6, training=True,
                        use_playout_cap_randomization=False, batch_size=32, epochs=1)
        torch.manual_seed(0)                                   # identical initial weights on every rank
        net = arch.build_network(mc).cuda()
        # every rank generates its own shard of games (ids rank, rank + world, ...) and trains on its own samples
        data, _ = ai.generate_games(cfg, net, 8, seed=11, first_game_id=rank, game_id_stride=world)
        n = torch.tensor([len(data)], device="cuda")
        dist.all_reduce(n, op=dist.ReduceOp.MIN)
        n_common = int(n.item()) // cfg.batch_size * cfg.batch_size
        assert n_common >= cfg.batch_size
        training.train_network_pytorch(cfg, net, data[:n_common], log=False)
        flat = torch.cat([t.detach().reshape(-1).double() for t in list(net.parameters()) + list(net.buffers())])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = float((hi - lo).abs().max())
        # sharded gating with deterministic evaluators: the union of the shards is the one-rank set of games
        ev1 = fake_evaluator_torch(torch.device("cuda", rank))

        def ev2(grids, extras):
            v, l = ev1(grids, extras)
            return (1.0 - v).contiguous(), l.roll(17, dims=1).contiguous()
        gate = cfg.copy()
        gate.training = False
        wins, _ = training.battle_networks(ev1, gate, ev2, gate, None, "more", 8, seed=5, first_game_id=rank, game_id_stride=world)
        w = torch.tensor(wins, device="cuda", dtype=torch.float64)
        dist.all_reduce(w)
        full = None
        if rank == 0:
            full, _ = training.battle_networks(ev1, gate, ev2, gate, None, "more", 8 * world, seed=5)
        q.put((rank, spread, w.cpu().tolist(), None if full is None else full.tolist(), n_common))
    finally:
        dist.destroy_process_group()


def test_nccl_allreduce_training_and_sharded_gating_world2():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, spread, wins, full, n_common = q.get(timeout=600)
        res[rank] = (spread, wins, full, n_common)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][0] == 0.0 and res[1][0] == 0.0          # replicas bit-identical after the all-reduced steps
    assert res[0][1] == res[1][1] and sum(res[0][1]) == 16
    assert res[0][2] == res[0][1]                         # sharded sums == the one-rank run
