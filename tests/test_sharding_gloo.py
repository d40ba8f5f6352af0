"""world_size-2 gloo test of the multi-GPU path's only collectives (eval_utils.py:523-527,565-568; value_net.py:195-201)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deer_vla_b200.sharding import all_gather_values, gather_results, shard_sequences
    seqs = list(range(8))
    mine = shard_sequences(seqs, rank, world)
    T = 3
    actions = torch.tensor([[[float(s * 10 + t)] * 7 for t in range(T)] for s in mine])        # [k, T, 7]
    exits = torch.tensor([[s + t for t in range(T)] for s in mine], dtype=torch.int32)         # [k, T]
    res = gather_results(actions, exits, dst=0)
    vals = all_gather_values(torch.full((2, 4), float(rank)))
    ok = vals.shape == (2, 8) and bool((vals[:, :4] == 0).all()) and bool((vals[:, 4:] == 1).all())
    if rank == 0:
        acts, ex = res
        ok = ok and acts.shape == (8, T, 7) and ex.shape == (8, T) and ex.dtype == torch.int32
        ok = ok and all(float(acts[s, t, 0]) == s * 10 + t and int(ex[s, t]) == s + t for s in seqs for t in range(T))
    else:
        ok = ok and res is None
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
