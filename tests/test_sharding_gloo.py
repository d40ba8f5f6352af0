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
    # ExitController.set_threshold(values=None): each rank calibrates on its own shard, the solver sees the gathered table
    # (value_net.py:192-201) -> identical thresholds on every rank, equal to the single-process solve over all samples
    from deer_vla_b200.value_net import ExitController, solve_thresholds

    class _Eng:                                  # stands in for the engine's full-depth step: rank-dependent delta columns
        def __init__(self, r):
            self.r = r
            self.got = None

        def reset(self):
            pass

        def step(self, full_depth, **b):
            g = torch.Generator().manual_seed(100 + self.r)
            return {"delta": torch.rand(16, 3, generator=g)}

        def set_thresholds(self, t):
            self.got = list(t)
    ctrl = ExitController(None, exit_id_list=[1, 3, 5], steps_per_stage=1, leq=True, max_layer=6)
    ctrl.engine = _Eng(rank)
    ctrl.set_threshold(None, None, [dict()], 1.0, "mpt_dolly_3b", values=None)
    full = torch.cat([torch.rand(16, 3, generator=torch.Generator().manual_seed(100 + r)).t() for r in range(world)], dim=1)
    want = solve_thresholds(full, 1.0, "exp", "mpt_dolly_3b", True)
    ok = ok and all(abs(ctrl.thresholds[e] - float(want[i])) < 1e-12 for i, e in enumerate([1, 3, 5])) and ctrl.engine.got is not None
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
