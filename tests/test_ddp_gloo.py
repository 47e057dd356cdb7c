"""world_size-2 gloo test (CPU) of the bucketed gradient reducer: N-rank reduced grads == single-process grads on the
concatenated batch (SURVEY.md §8e verification), incl. multi-bucket splitting and no_sync accumulation."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.Tanh(), torch.nn.Linear(64, 32), torch.nn.Tanh(),
                               torch.nn.Linear(32, 1))


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neurosis_b200.ddp import BucketedGradReducer
    m = _model()
    red = BucketedGradReducer(m.parameters(), bucket_mb=0.004)  # ~1k floats per bucket -> several buckets
    assert len(red.buckets) >= 3
    torch.manual_seed(123)
    data = torch.randn(8, 16)
    mine = data[rank * 4:(rank + 1) * 4]
    red.zero_grad()
    m(mine).pow(2).mean().backward()
    red.finish()
    g1 = [p.grad.clone() for p in m.parameters()]
    # gradient accumulation: first micro-step local only, second reduced
    red.zero_grad()
    with red.no_sync():
        m(mine[:2]).pow(2).mean().backward()
    m(mine[2:]).pow(2).mean().backward()
    red.finish()
    g2 = [p.grad.clone() for p in m.parameters()]
    if rank == 0:
        q.put(([g.numpy() for g in g1], [g.numpy() for g in g2]))
    dist.barrier()
    dist.destroy_process_group()


def test_reducer_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g1, g2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _model()
    torch.manual_seed(123)
    data = torch.randn(8, 16)
    # mean over ranks of per-rank mean losses == mean over the concatenated batch (equal shard sizes)
    m(data).pow(2).mean().backward()
    for got, p in zip(g1, m.parameters()):
        assert torch.allclose(torch.from_numpy(got), p.grad, rtol=1e-5, atol=1e-7)
    m.zero_grad()
    (sum(m(data[i:i + 2]).pow(2).mean() for i in (0, 2, 4, 6)) / 2).backward()  # two micro-steps per rank, averaged over ranks
    for got, p in zip(g2, m.parameters()):
        assert torch.allclose(torch.from_numpy(got), p.grad, rtol=1e-5, atol=1e-7)
