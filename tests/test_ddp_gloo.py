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


# ---------------------------------------------------------------------------------------------------------------
# optimizer-state sharding (ZeRO-1 style): reduce to the bucket owner -> owner's optimizer step -> broadcast
# ---------------------------------------------------------------------------------------------------------------
def _sharded_worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neurosis_b200.ddp import ShardedOptimizerReducer
    m = _model()
    names = [n for n, _ in m.named_parameters()]
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    red = ShardedOptimizerReducer(m.parameters(), bucket_mb=0.004)
    assert len(red.buckets) >= 3
    assert all(torch.equal(before[n], p.detach()) for n, p in m.named_parameters()), "re-homing keeps the values"
    assert [n for n, _ in m.named_parameters()] == names and set(m.state_dict()) == set(before)
    owned = red.owned_params()
    assert 0 < len(owned) < len(names), "each rank owns a strict subset"
    opt = torch.optim.Adam(owned, lr=1e-2)  # optimizer state only for the owned parameters
    torch.manual_seed(123)
    data = torch.randn(3, 8, 16)
    for step in range(3):
        red.zero_grad()
        m(data[step, rank * 4:(rank + 1) * 4]).pow(2).mean().backward()
        red.finish()
        opt.step()
        red.broadcast_params()
    n_state = sum(len(opt.state[p]) > 0 for p in owned)
    q.put((rank, {n: p.detach().numpy().copy() for n, p in m.named_parameters()}, len(owned), n_state))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    owned_total = 0
    for _ in range(2):
        rank, params, n_owned, n_state = q.get(timeout=120)
        got[rank] = params
        owned_total += n_owned
        assert n_state == n_owned
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _model()
    assert owned_total == len(list(m.parameters())), "every parameter has exactly one owner"
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    torch.manual_seed(123)
    data = torch.randn(3, 8, 16)
    for step in range(3):
        opt.zero_grad()
        m(data[step]).pow(2).mean().backward()  # mean over the concatenated batch == mean of the two rank means
        opt.step()
    for n, p in m.named_parameters():
        for rank in (0, 1):
            assert torch.allclose(torch.from_numpy(got[rank][n]), p.detach(), rtol=1e-5, atol=1e-6), (rank, n)


# ---- gradients written straight into the bucket storage ("gradient sink"), as the CUDA kernels do -----------------------
class _SinkLinearFn(torch.autograd.Function):
    """y = x W^T with the weight gradient ACCUMULATED INTO THE REDUCER'S BUCKET by the backward itself (what
    neurosis_b200.ops.LinearFn does through nk_linear_wgrad): autograd gets None for the weight and the reducer is told
    through `mark_ready`."""

    @staticmethod
    def forward(ctx, x, w, sink):
        ctx.save_for_backward(x, w)
        ctx.sink = sink
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        buf = ctx.sink.buffer_for(w)
        assert buf is not None
        buf.add_(dy.t() @ x)
        ctx.sink.mark_ready(w)
        return dy @ w, None, None


def _sink_worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neurosis_b200.ddp import BucketedGradReducer
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(24, 24) * 0.2) for _ in range(6)]
    bias = torch.nn.Parameter(torch.zeros(24))  # an ordinary autograd-accumulated gradient in the same buckets
    red = BucketedGradReducer(ws + [bias], bucket_mb=0.005)  # two 24x24 weights per bucket
    assert len(red.buckets) >= 3
    torch.manual_seed(123)
    data = torch.randn(8, 24)
    mine = data[rank * 4:(rank + 1) * 4]
    launched = []
    orig = red._launch
    red._launch = lambda b: (launched.append([bb["pending"] for bb in red.buckets]), orig(b))[1]
    red.zero_grad()
    h = mine
    for w in ws:
        h = torch.tanh(_SinkLinearFn.apply(h, w, red))
    (h + bias).pow(2).mean().backward()
    pending = [b["pending"] for b in red.buckets]
    red.finish()
    if rank == 0:
        q.put(([p.grad.clone().numpy() for p in ws + [bias]], pending, len(launched)))
    dist.barrier()
    dist.destroy_process_group()


def test_sunk_gradients_are_counted_once():
    """Regression for the premature all-reduce found on 2 B200s in round 2: torch fires the post-accumulate hook also for
    parameters whose Function returned None, so a parameter that reported itself through `mark_ready` was counted twice
    and its bucket was reduced when only half of its gradients had been written."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sink_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, pending, n_launched = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(v == 0 for v in pending), pending  # every bucket launched exactly when its last gradient arrived
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(24, 24) * 0.2) for _ in range(6)]
    bias = torch.nn.Parameter(torch.zeros(24))
    torch.manual_seed(123)
    data = torch.randn(8, 24)
    tot = 0
    for half in (data[:4], data[4:]):
        h = half
        for w in ws:
            h = torch.tanh(h @ w.t())
        tot = tot + (h + bias).pow(2).mean() / 2
    tot.backward()
    for g, p in zip(grads, ws + [bias]):
        assert torch.allclose(torch.from_numpy(g), p.grad, rtol=1e-5, atol=1e-7)
