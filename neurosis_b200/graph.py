"""CUDA-graph capture of the whole training step (VAE encode -> loss -> backward -> gradient all-reduce).

The step launches ~4 000 kernels of this library plus the autograd engine's own bookkeeping; issued eagerly that is
~0.3 s of Python/ctypes work per step, comparable to the GPU time.  Shapes are static in training (aspect buckets give
a handful of shapes — one graph per bucket), so the step is captured once and replayed: the host only refreshes the
static input buffers (images, conditioning, the per-sample sigma draw and loss weights) and launches one graph.

Optionally the optimizer step (`neurosis_b200.optim.Adafactor`) and the EMA update (`optim.LitEma`) are part of the
captured step (`optimizer=`, `ema=`): their per-step scalars (step count -> beta2t / relative step size, EMA decay
warm-up) live in device memory and are advanced by kernels inside the graph, so a replay is a complete training
iteration: refresh -> encode -> loss -> backward -> all-reduce -> parameter update -> EMA.

Host-side semantics stay those of `DiffusionEngine.training_step` (reference models/diffusion.py:205-233): sigma draw
on the CPU generator (`StandardDiffusionLoss.draw_sigmas`), loss hooks as per-sample weights, `loss.mean()`.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .ddp import BucketedGradReducer
from .engine import DiffusionEngine


class GraphedTrainStep:
    def __init__(self, engine: DiffusionEngine, reducer: BucketedGradReducer, image: Tensor, crossattn: Tensor,
                 vector: Optional[Tensor], warmup: int = 3, optimizer=None, ema=None, pool=None):
        self.engine, self.reducer = engine, reducer
        self.optimizer, self.ema = optimizer, ema
        if getattr(engine.loss_fn, "noise_offset", 0.0):
            # apply_noise_offset draws its chance on the host and its offset with CPU torch.randn (loss.py): captured,
            # both would be frozen into the graph and every replay would reuse the same offset
            raise NotImplementedError("GraphedTrainStep: loss_fn.noise_offset > 0 is a per-step host-side random draw "
                                      "and cannot be captured; run the eager DiffusionEngine.training_step instead")
        self._opt_ready = self._ema_ready = False
        dev = image.device
        self.image = image.clone()
        self.crossattn = crossattn.clone()
        self.vector = vector.clone() if vector is not None else None
        n = image.shape[0]
        self.sigmas = torch.ones(n, dtype=torch.float32, device=dev)
        self.weights = torch.ones(n, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.per_sample = torch.zeros(n, dtype=torch.float32, device=dev)
        self._refresh_sigmas()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._core()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the eager warm-up's activations must not stay cached next to the graph's pool
        # the fp32 -> bf16 / packed weight refresh is the first thing `_core` does, so it is captured and replayed every
        # step; run it once here so its span table is built outside the capture
        ops.refresh_weight_copies(force=True)
        if self.optimizer is not None:  # gradients exist now: build the optimizer's device tables outside the capture
            self.optimizer.graph_prepare()
            self._opt_ready = True
        if self.ema is not None:
            self.ema.graph_prepare(engine.model)
            self._ema_ready = True
        l0 = ops.LAUNCHES
        # `pool`: graphs that never replay concurrently (one captured step per aspect bucket) share one memory pool, so
        # only the largest bucket's activations are resident instead of the sum over buckets
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self._core()
        self.pool = self.graph.pool()
        self.launches_per_replay = ops.LAUNCHES - l0

    def _core(self) -> None:
        eng = self.engine
        ops.refresh_weight_copies(force=True)  # the optimizer changed the fp32 weights since the last step
        self.reducer.zero_grad()
        z = eng.encode_first_stage(self.image) if eng.first_stage_model is not None else self.image
        cond = {"crossattn": self.crossattn}
        if self.vector is not None:
            cond["vector"] = self.vector
        # hook weights (tag-frequency scaling) enter the reduction kernel's weight vector, not a multiply afterwards
        loss = eng.loss_fn._forward(eng.model, eng.denoiser, cond, z, {}, sigmas=self.sigmas, sample_weights=self.weights)
        total = loss.mean()
        total.backward()
        self.reducer.finish()
        self.per_sample.copy_(loss.detach())
        self.loss.copy_(total.detach())
        if self._opt_ready:  # (the eager warm-up steps before the tables exist leave the parameters untouched)
            self.optimizer.graph_launch()
            if hasattr(self.reducer, "broadcast_params"):  # ShardedOptimizerReducer: owners publish their updates
                self.reducer.broadcast_params()
        if self._ema_ready:
            self.ema.graph_launch()

    def _refresh_sigmas(self) -> None:
        s = self.engine.loss_fn.draw_sigmas(self.sigmas.shape[0]).float()
        self.sigmas.copy_(s)  # a few bytes, stream-ordered before the replay

    def step(self, image: Optional[Tensor] = None, crossattn: Optional[Tensor] = None, vector: Optional[Tensor] = None,
             weights: Optional[Tensor] = None) -> Tensor:
        """copy the new batch into the static buffers (host or device tensors), draw sigmas, replay.  Returns the
        device scalar loss (read it with .item() when needed)."""
        if image is not None:
            self.image.copy_(image, non_blocking=True)
        if crossattn is not None:
            self.crossattn.copy_(crossattn, non_blocking=True)
        if vector is not None and self.vector is not None:
            self.vector.copy_(vector, non_blocking=True)
        if weights is not None:
            self.weights.copy_(weights, non_blocking=True)
        self._refresh_sigmas()
        self.graph.replay()
        return self.loss

    def sync_host_state(self) -> None:
        """bring the host mirrors of the device-side counters up to date (before checkpointing / leaving graph mode)."""
        if self.optimizer is not None:
            self.optimizer.sync_steps_from_device()
        if self.ema is not None:
            self.ema.sync_from_device()
