"""Data-parallel gradient reduction for the training step: bucketed all-reduce overlapped with backward.

The reference never calls a collective itself — Lightning's `strategy: auto` wraps the module in torch DDP
(25 MiB buckets, fp32 grads).  Here the reducer is explicit and sized for NVLink 5 / NVSwitch:

  * parameters are grouped, in reverse registration order (the order backward produces gradients), into buckets of
    `bucket_mb` (default 256 MiB — NVSwitch saturates only on large messages, link count is not the limit);
  * every parameter's `.grad` is a VIEW into its bucket's flat fp32 buffer, so no pack/unpack kernels exist;
  * `register_post_accumulate_grad_hook` counts ready gradients; when a bucket is complete its all-reduce is
    enqueued on a side stream (after an event on the compute stream) and runs while backward continues;
  * `finish()` makes the compute stream wait for the outstanding reductions.

Works with NCCL (CUDA tensors, op AVG) and with gloo (CPU tensors, SUM then scale) — the latter is what the
world_size-2 CPU tests exercise.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist
from torch import nn


class BucketedGradReducer:
    def __init__(self, params: Iterable[nn.Parameter], bucket_mb: float = 256.0,
                 process_group: Optional[dist.ProcessGroup] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: list[dict] = []
        cap = int(bucket_mb * 1024 * 1024 / 4)
        cur: list[nn.Parameter] = []
        n = 0
        for p in reversed(self.params):
            if cur and n + p.numel() > cap:
                self._make_bucket(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            self._make_bucket(cur)
        self._index = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._index[p] = bi
        self.enabled = True
        self._works: list = []
        self._timeline = None  # start_timeline(): CUDA events per bucket all-reduce of ONE step (diagnostics, off by default)
        # Two sources report a gradient as complete: autograd's post-accumulate hook (gradients that autograd itself
        # accumulates) and `mark_ready` (kernels of neurosis_b200.ops that wrote straight into the bucket storage).
        # torch fires the post-accumulate hook for EVERY parameter an autograd Function was asked a gradient for, even
        # when the Function returned None because a kernel already wrote the bucket (found by the 2-rank NCCL equality
        # test: every sunk parameter was counted twice, so buckets were all-reduced half-way through backward).  A
        # parameter whose storage was handed out through `buffer_for` this step is therefore counted by `mark_ready`
        # only — its hook may even fire BEFORE the side-stream weight-gradient job has been joined.
        self._sunk: set = set()
        self._counted: set = set()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_hook) for p in self.params]
        dev = self.params[0].device
        self._cuda = dev.type == "cuda"
        self._stream = torch.cuda.Stream(device=dev) if self._cuda else None

    def _make_bucket(self, ps: list) -> None:
        total = sum(p.numel() for p in ps)
        flat = torch.zeros(total, dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append({"params": ps, "flat": flat, "pending": len(ps)})

    # -- per step ---------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        self._sunk.clear()
        self._counted.clear()
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            for p, view in zip(b["params"], self._views(b)):
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view

    def _views(self, b: dict):
        off = 0
        for p in b["params"]:
            yield b["flat"][off: off + p.numel()].view_as(p)
            off += p.numel()

    # -- gradient sink protocol used by neurosis_b200.ops (weight gradients are written straight into the buckets) --
    def buffer_for(self, p: nn.Parameter):
        if p not in self._index:
            return None
        g = p.grad
        if g is None or g.dtype != torch.float32:
            return None
        self._sunk.add(id(p))
        return g

    def mark_ready(self, p: nn.Parameter) -> None:
        self._on_grad(p)

    def _on_hook(self, p: nn.Parameter) -> None:
        if id(p) in self._sunk:
            return  # a kernel owns this gradient; `mark_ready` reports it once the write is ordered on the main stream
        self._on_grad(p)

    def attach_as_grad_sink(self) -> None:
        from . import ops
        ops.GRAD_SINK = self

    def detach_grad_sink(self) -> None:
        from . import ops
        if self._cuda:
            ops.join_side_streams()
        if ops.GRAD_SINK is self:
            ops.GRAD_SINK = None

    def _on_grad(self, p: nn.Parameter) -> None:
        if not self.enabled or self.world == 1:
            return
        if id(p) in self._counted:  # once per parameter and step, whatever the source
            return
        self._counted.add(id(p))
        b = self.buckets[self._index[p]]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    # -- diagnostics: where do the bucket all-reduces sit relative to backward? ------------------------------------
    def start_timeline(self) -> None:
        """record CUDA events around every bucket all-reduce of the NEXT step (call before zero_grad / backward)."""
        if self._cuda:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record()
            self._timeline = {"t0": t0, "buckets": [], "backward_done": None}

    def end_timeline(self) -> Optional[dict]:
        """after finish(): milliseconds relative to start_timeline().  `exposed_ms` = time the compute stream had to wait for
        communication after its own last kernel (end of the last all-reduce minus end of backward, floored at 0)."""
        tl, self._timeline = self._timeline, None
        if not tl or tl["backward_done"] is None:
            return None
        torch.cuda.synchronize()
        t0 = tl["t0"]
        rows = [{"bucket": i, "mbytes": round(n * 4 / 2 ** 20, 1), "start_ms": round(t0.elapsed_time(a), 3),
                 "end_ms": round(t0.elapsed_time(b), 3)} for i, n, a, b in tl["buckets"]]
        bwd = t0.elapsed_time(tl["backward_done"])
        last = max((r["end_ms"] for r in rows), default=bwd)
        busy = sum(r["end_ms"] - r["start_ms"] for r in rows)
        return {"backward_done_ms": round(bwd, 3), "last_allreduce_end_ms": round(last, 3),
                "exposed_ms": round(max(0.0, last - bwd), 3), "allreduce_busy_ms": round(busy, 3), "buckets": rows}

    def _launch(self, b: dict) -> None:
        flat = b["flat"]
        if self._cuda:
            ev = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                if self._timeline is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self._stream)
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                    e1.record(self._stream)
                    self._timeline["buckets"].append((self.buckets.index(b), flat.numel(), e0, e1))
                    return
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((w, flat))

    def finish(self) -> None:
        """call after backward: every bucket has been reduced when this returns (stream-ordered on CUDA)."""
        if self._cuda:
            from . import ops
            ops.join_side_streams()  # weight gradients computed on the side stream (marks their parameters ready)
        if self.world == 1 or not self.enabled:
            # inside no_sync(): a gradient-accumulation micro-step only joins the side streams (mark_ready is a no-op
            # while disabled), nothing is communicated and the buckets keep accumulating
            return
        for b in self.buckets:  # parameters that received no gradient this step still take part
            if b["pending"] > 0:
                self._launch(b)
                b["pending"] = 0
        if self._cuda:
            if self._timeline is not None:  # end of this rank's own backward work, before it waits for communication
                done = torch.cuda.Event(enable_timing=True)
                done.record()
                self._timeline["backward_done"] = done
            torch.cuda.current_stream().wait_stream(self._stream)
        else:
            for w, flat in self._works:
                w.wait()
                flat.div_(self.world)
            self._works.clear()

    def no_sync(self):
        """context manager: accumulate locally, skip the reduction (gradient accumulation micro-steps)."""
        reducer = self

        class _Ctx:
            def __enter__(self):
                reducer.enabled = False

            def __exit__(self, *a):
                if reducer._cuda:
                    from . import ops
                    ops.join_side_streams()  # retire side-stream jobs while mark_ready is still a no-op
                reducer.enabled = True
                reducer._sunk.clear()
                reducer._counted.clear()
                for b in reducer.buckets:  # the boundary micro-step counts every parameter again
                    b["pending"] = len(b["params"])

        return _Ctx()


class ShardedOptimizerReducer(BucketedGradReducer):
    """Optimizer-state sharding over the data-parallel ranks (ZeRO-1 style) on top of the bucketed reducer — the
    multi-GPU form of SURVEY.md §8(f) row 2 ("optimizer step fused with the all-reduce epilogue").

    An all-reduce is a reduce-scatter followed by an all-gather; here the optimizer step sits between the two halves:

      * every bucket has an OWNER rank (round-robin); its gradients are reduced (averaged) TO the owner only —
        `reduce` instead of `all_reduce`, launched per completed bucket on the side stream during backward as before;
      * parameters live in per-bucket flat fp32 buffers (each `nn.Parameter` is re-homed as a view; names, shapes and
        state dicts are unchanged), and each rank builds its optimizer over `owned_params()` only: optimizer compute
        and optimizer state (Adafactor moments, Adam m/v, EMA shadows) shrink by the world size;
      * after `optimizer.step()` the owner broadcasts the updated flat parameter buffer (`broadcast_params()`).

    Bytes on the wire equal the all-reduce's (reduce + broadcast of every bucket); over NVSwitch every GPU has full
    bandwidth to every peer, so the round-robin owners' transfers proceed concurrently.  Adafactor's factored
    statistics need whole tensors, which is why ownership is per bucket of whole parameters and not per element range.
    Works with NCCL (CUDA) and gloo (CPU; the world_size-2 test)."""

    def __init__(self, params: Iterable[nn.Parameter], bucket_mb: float = 256.0,
                 process_group: Optional[dist.ProcessGroup] = None):
        super().__init__(params, bucket_mb, process_group)
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        for bi, b in enumerate(self.buckets):
            b["owner"] = bi % self.world
            pflat = torch.empty_like(b["flat"])
            off = 0
            with torch.no_grad():
                for p in b["params"]:
                    n = p.numel()
                    pflat[off: off + n].copy_(p.detach().reshape(-1))
                    p.data = pflat[off: off + n].view_as(p)  # same Parameter object, storage re-homed
                    off += n
            b["pflat"] = pflat
        self._bworks: list = []

    def owned_params(self) -> list:
        """parameters whose gradients are reduced to this rank: build the optimizer (and EMA) over these."""
        return [p for b in self.buckets if b["owner"] == self.rank for p in b["params"]]

    def _launch(self, b: dict) -> None:
        flat, owner = b["flat"], b["owner"]
        dst = dist.get_global_rank(self.group, owner) if self.group is not None else owner
        if self._cuda:
            ev = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                dist.reduce(flat, dst=dst, op=dist.ReduceOp.AVG, group=self.group)
        else:
            w = dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((w, flat))

    def broadcast_params(self) -> None:
        """after the owners' optimizer steps: every rank receives every bucket's updated parameters.  CUDA: enqueued on
        the side stream behind the optimizer kernels, the compute stream waits for all of them."""
        if self.world == 1:
            return
        changed = []
        if self._cuda:
            ev = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                for b in self.buckets:
                    src = dist.get_global_rank(self.group, b["owner"]) if self.group is not None else b["owner"]
                    dist.broadcast(b["pflat"], src=src, group=self.group)
            torch.cuda.current_stream().wait_stream(self._stream)
            for b in self.buckets:
                if b["owner"] != self.rank:
                    changed.extend((p, False) for p in b["params"])
            from . import ops
            ops.parameters_updated_in_place(changed)  # raw writes into parameter storage: bump the version counters
        else:
            works = []
            for b in self.buckets:
                src = dist.get_global_rank(self.group, b["owner"]) if self.group is not None else b["owner"]
                works.append(dist.broadcast(b["pflat"], src=src, group=self.group, async_op=True))
            for w in works:
                w.wait()
