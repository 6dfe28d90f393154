"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: game-id sharding without any
data-path collective, and the flat gradient all-reduce used by the training step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tetris_reinforcement_learning_b200.selfplay import shard_for_rank
        from tetris_reinforcement_learning_b200.training import allreduce_gradients
        sh = shard_for_rank(rank, world, 8)
        ids = torch.tensor([sh["first_game_id"] + k * sh["game_id_stride"] for k in range(24)])
        gathered = [torch.zeros_like(ids) for _ in range(world)]
        dist.all_gather(gathered, ids)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        x = torch.full((5, 4), float(rank + 1))
        model(x).sum().backward()
        local = [p.grad.clone() for p in model.parameters()]
        allreduce_gradients(model)
        q.put((rank, [g.tolist() for g in gathered], [g.tolist() for g in local],
               [p.grad.tolist() for p in model.parameters()]))
    finally:
        dist.destroy_process_group()


def test_sharding_and_gradient_allreduce_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, gathered, local, reduced = q.get(timeout=120)
        results[rank] = (gathered, local, reduced)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_ids = sorted(i for ids in results[0][0] for i in ids)
    assert all_ids == list(range(48))                      # disjoint and complete cover of the id space
    for k in range(len(results[0][1])):
        mean = (np.array(results[0][1][k]) + np.array(results[1][1][k])) / 2
        assert np.allclose(results[0][2][k], mean) and np.allclose(results[1][2][k], mean)


def test_shard_validation():
    from tetris_reinforcement_learning_b200.selfplay import shard_for_rank
    assert shard_for_rank(3, 8, 4096) == {"first_game_id": 3, "game_id_stride": 8, "n_games": 4096}
    with pytest.raises(ValueError):
        shard_for_rank(8, 8, 1)
