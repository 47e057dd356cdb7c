"""Optimizer-side drop-ins: the reference's `Adafactor` / `AdafactorScheduler`
(/root/reference/src/neurosis/optimizers/adafactor.py:13-256, selected by configs/sdxl/sdxl.example.yaml:158-169) and
`LitEma` (modules/ema.py:11-97) on multi-tensor sm_100a kernels (`csrc/optim.cu`).

One `Adafactor.step()` is five stream-ordered launches over ALL parameters (1 680 tensors / 2.57 G elements for SDXL)
instead of ~15 ATen kernels per parameter, and the apply pass also rewrites the bf16 weight mirrors that the GEMMs
consume (`ops.bf16_weight`), so the per-step fp32 -> bf16 refresh pass disappears for those parameters.
Constructor signatures, `param_groups`, and the per-parameter state keys (`step`, `exp_avg_sq_row`, `exp_avg_sq_col`,
`exp_avg_sq`, `exp_avg`, `RMS`) are the reference's, so optimizer state dicts interchange.

There is no CPU fallback: parameters and gradients must be CUDA fp32 tensors.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn
from torch.optim.lr_scheduler import LambdaLR
from torch.optim.optimizer import Optimizer

from . import ops
from ._lib import check, lib

KIND_VEC, KIND_SMALL, KIND_MAT = 0, 1, 2
TILE_R, TILE_C, VEC_CHUNK, SMALL_PER_BLOCK = 64, 128, 4096, 1024
MAX_SPLIT = 4096  # a [Bt, R, C] parameter with large R*C is described as Bt separate matrices up to this many

_TENSOR_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("vr", "<u8"), ("vc", "<u8"), ("exp_avg", "<u8"),
                          ("mirror", "<u8"), ("scratch", "<u8"), ("n", "<i8"), ("kind", "<i4"), ("Bt", "<i4"),
                          ("R", "<i4"), ("C", "<i4"), ("group", "<i4"), ("owner", "<i4")])
assert _TENSOR_DTYPE.itemsize == 88


def factor_dims(shape) -> tuple[int, int, int]:
    """(Bt, R, C) of the reference's factoring over the last two dims (adafactor.py:193-194, 220-221)."""
    shape = tuple(shape)
    return int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1, int(shape[-2]), int(shape[-1])


def blocks_for(kind: int, n: int, Bt: int, R: int, C: int) -> int:
    if kind == KIND_VEC:
        return -(-n // VEC_CHUNK)
    if kind == KIND_SMALL:
        return -(-Bt // SMALL_PER_BLOCK)
    return -(-R // TILE_R) * -(-C // TILE_C)


def plan_tensor(shape) -> list[tuple[int, int, int, int, int]]:
    """[(kind, elem_offset, Bt, R, C)] records describing one parameter (host logic, testable without a GPU)."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if len(shape) < 2:
        return [(KIND_VEC, 0, 1, 1, n)]
    Bt, R, C = factor_dims(shape)
    if R * C <= 64:
        return [(KIND_SMALL, 0, Bt, R, C)]
    if Bt > MAX_SPLIT:
        raise NotImplementedError(f"Adafactor: parameter of shape {shape} has {Bt} large matrices")
    return [(KIND_MAT, b * R * C, 1, R, C) for b in range(Bt)]


class Adafactor(Optimizer):
    """Same arguments and update rule as the reference Adafactor (optimizers/adafactor.py:100-246).

    Not reproduced: `relative_step=False` together with `scale_parameter=True` — there the reference overwrites
    `group["lr"]` with `lr * RMS(p)` after every parameter (adafactor.py:214), so the external learning rate compounds
    across parameters and steps; that combination raises here."""

    def __init__(self, params, lr: Optional[float] = None, eps: tuple[float, float] = (1e-30, 1e-3),
                 clip_threshold: float = 1.0, decay_rate: float = -0.8, beta1: Optional[float] = None,
                 weight_decay: float = 0.0, scale_parameter: bool = True, relative_step: bool = True,
                 warmup_init: bool = False, refresh_mirrors: bool = True):
        if lr is not None and relative_step:
            raise ValueError("Cannot combine manual `lr` and `relative_step=True` options")
        if warmup_init and not relative_step:
            raise ValueError("`warmup_init=True` requires `relative_step=True`")
        if not relative_step and scale_parameter:
            raise NotImplementedError("relative_step=False with scale_parameter=True compounds group['lr'] in the "
                                      "reference (adafactor.py:214); use scale_parameter=False with an external lr")
        defaults = dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate, beta1=beta1,
                        weight_decay=weight_decay, scale_parameter=scale_parameter, relative_step=relative_step,
                        warmup_init=warmup_init, differentiable=False)
        super().__init__(params, defaults)
        self.refresh_mirrors = refresh_mirrors
        self._plan = None
        self._key = None

    def load_state_dict(self, state_dict) -> None:
        """resume: the loaded moment tensors are new objects, so the device table (which holds raw pointers to the
        state) is rebuilt on the next step."""
        super().load_state_dict(state_dict)
        self._plan = None
        self._key = None

    def zero_grad(self, set_to_none: bool = False) -> None:
        """zero IN PLACE by default (torch's default drops the tensors): stable gradient storage keeps the device
        table valid from step to step — with `BucketedGradReducer` the gradients are views of its buckets anyway."""
        super().zero_grad(set_to_none=set_to_none)

    # ---- reference helpers (host scalars) ------------------------------------------------------
    @staticmethod
    def _rel_step(group: dict, step: int) -> float:
        if group["relative_step"]:
            min_step = 1e-6 * step if group["warmup_init"] else 1e-2
            return min(min_step, 1.0 / math.sqrt(step))
        return group["lr"]

    @staticmethod
    def _get_lr(param_group: dict, param_state: dict) -> float:
        """adafactor.py:130-139 (used by AdafactorScheduler.get_lr)."""
        rel = Adafactor._rel_step(param_group, max(1, param_state["step"]))
        scale = 1.0
        if param_group["scale_parameter"]:
            scale = max(param_group["eps"][1], float(param_state["RMS"]))
        return scale * rel

    # ---- device tables -------------------------------------------------------------------------
    def _build(self) -> None:
        rows, blk_start, live, scratches = [], [], [], []
        nblk = 0
        dev = None
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("neurosis_b200 Adafactor needs contiguous CUDA fp32 parameters (no CPU fallback)")
                g = p.grad
                if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()):
                    raise RuntimeError("neurosis_b200 Adafactor needs contiguous CUDA fp32 gradients")
                dev = p.device
                st = self.state[p]
                factored = p.dim() >= 2
                if len(st) == 0:
                    st["step"] = 0
                    if group["beta1"] is not None:
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    if factored:
                        st["exp_avg_sq_row"] = torch.zeros(p.shape[:-1], dtype=torch.float32, device=dev)
                        st["exp_avg_sq_col"] = torch.zeros(p.shape[:-2] + p.shape[-1:], dtype=torch.float32, device=dev)
                    else:
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["RMS"] = 0
                mirror = ops.registered_mirror(p) if self.refresh_mirrors else None
                ex = st.get("exp_avg")
                recs = plan_tensor(p.shape)
                owner = len(rows)
                for kind, off, Bt, R, C in recs:
                    n = p.numel()
                    b = off // (R * C) if kind == KIND_MAT else 0
                    scratch = 0
                    if kind == KIND_MAT:
                        sc = torch.zeros(R + C, dtype=torch.float32, device=dev)
                        scratches.append(sc)
                        scratch = sc.data_ptr()
                    if kind == KIND_VEC:
                        vr, vc = st["exp_avg_sq"].data_ptr(), 0
                    else:
                        vr = st["exp_avg_sq_row"].data_ptr() + 4 * b * R
                        vc = st["exp_avg_sq_col"].data_ptr() + 4 * b * C
                    rows.append((p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, vr, vc,
                                 ex.data_ptr() + 4 * off if ex is not None else 0,
                                 mirror.data_ptr() + 2 * off if mirror is not None else 0, scratch, n, kind, Bt, R, C,
                                 gi, owner))
                    blk_start.append(nblk)
                    nblk += blocks_for(kind, n, Bt, R, C)
                live.append((p, len(recs), mirror))  # the mirror tensor is held: the table stores its raw pointer
        if not rows:
            self._plan = None
            return
        table = np.array(rows, dtype=_TENSOR_DTYPE)
        T = len(rows)
        self._plan = dict(
            table=torch.from_numpy(table.view(np.uint8).copy()).to(dev),
            blk_start=torch.tensor(blk_start, dtype=torch.int32).to(dev),
            n_tensors=T, n_blocks=nblk, live=live, scratch=scratches,
            hyper_host=torch.zeros((len(self.param_groups), 8), dtype=torch.float32).pin_memory(),
            hyper=torch.zeros((len(self.param_groups), 8), dtype=torch.float32, device=dev),
            scal=torch.zeros((T, 4), dtype=torch.float32, device=dev),
            rms=torch.zeros((T,), dtype=torch.float32, device=dev))
        # state["RMS"] of the reference = RMS(p) before the update: a view of the kernel's output slot
        t = 0
        for p, nrec, _ in live:
            self.state[p]["RMS"] = self._plan["rms"][t]
            t += nrec

    def _mirror_ptr(self, p) -> int:
        m = ops.registered_mirror(p) if self.refresh_mirrors else None
        return m.data_ptr() if m is not None else 0

    def _table_key(self):
        """parameter, gradient AND bf16-mirror addresses: mirrors are re-created by `ops.invalidate_weight_cache()`
        (engine.init_from_ckpt) and by `bf16_weight_group` re-homing a parameter in a stacked buffer — a table built
        before that would write bf16 through a dangling pointer."""
        return tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0, self._mirror_ptr(p))
                     for g in self.param_groups for p in g["params"])

    def _fresh_items(self):
        """[(parameter, mirror_is_fresh)]: a mirror counts as rewritten only if it is still the tensor the table wrote."""
        return [(p, m is not None and ops.registered_mirror(p) is m) for p, _, m in self._plan["live"]]

    def prepare(self) -> None:
        """(re)build the device tables if parameters / gradient buffers moved (host work; never inside a capture)."""
        key = self._table_key()
        if self._plan is None or key != self._key:
            self._build()
            self._key = key

    def _launch(self) -> None:
        plan = self._plan
        check(lib.nk_adafactor_step(plan["table"].data_ptr(), plan["blk_start"].data_ptr(), plan["n_tensors"],
                                    plan["n_blocks"], plan["hyper"].data_ptr(), plan["scal"].data_ptr(),
                                    plan["rms"].data_ptr(), ops._stream()), "adafactor_step")
        ops._count(4)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.prepare()
        plan = self._plan
        if plan is None:
            return loss
        for p, _, _ in plan["live"]:
            self.state[p]["step"] += 1
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            ss = {self.state[p]["step"] for p in ps}
            if len(ss) != 1:
                raise RuntimeError("neurosis_b200 Adafactor: parameters of one group must share the step count")
            step = ss.pop()
            h = plan["hyper_host"][gi]
            h[0] = 1.0 - math.pow(step, group["decay_rate"])
            h[1] = self._rel_step(group, step)
            h[2], h[3] = group["eps"]
            h[4] = group["clip_threshold"]
            h[5] = group["weight_decay"]
            h[6] = group["beta1"] if group["beta1"] is not None else 0.0
            h[7] = 1.0 if group["scale_parameter"] else 0.0
        plan["hyper"].copy_(plan["hyper_host"], non_blocking=True)
        self._launch()
        ops.parameters_updated_in_place(self._fresh_items())
        return loss

    # ---- CUDA-graph form: the step count lives on the device ------------------------------------------------------
    def graph_prepare(self) -> None:
        """call once before capture (after the gradients exist): builds the tables and moves the per-group constants
        and the current step counts to the device."""
        self.prepare()
        plan = self._plan
        if plan is None:
            raise RuntimeError("Adafactor.graph_prepare: no parameter has a gradient yet")
        dev = plan["hyper"].device
        consts, steps = [], []
        for group in self.param_groups:
            flags = (1 if group["relative_step"] else 0) | (2 if group["warmup_init"] else 0) | (
                4 if group["scale_parameter"] else 0)
            consts.append([group["decay_rate"], group["lr"] if group["lr"] is not None else 0.0, group["eps"][0],
                           group["eps"][1], group["clip_threshold"], group["weight_decay"],
                           group["beta1"] if group["beta1"] is not None else 0.0, float(flags)])
            ps = [p for p in group["params"] if p.grad is not None]
            steps.append(self.state[ps[0]]["step"] if ps else 0)
        plan["consts"] = torch.tensor(consts, dtype=torch.float32).to(dev)
        plan["step_dev"] = torch.tensor(steps, dtype=torch.int64).to(dev)

    @torch.no_grad()
    def graph_launch(self) -> None:
        """the capturable step: advance the device-side step counts, derive the hyper-parameters, update.  The host
        `state[p]["step"]` integers are brought up to date by `sync_steps_from_device()`."""
        plan = self._plan
        if plan is None or "consts" not in plan:
            raise RuntimeError("Adafactor.graph_launch: call graph_prepare() first")
        check(lib.nk_adafactor_hyper(plan["consts"].data_ptr(), plan["step_dev"].data_ptr(), plan["hyper"].data_ptr(),
                                     len(self.param_groups), ops._stream()), "adafactor_hyper")
        ops._count()
        self._launch()

    def sync_steps_from_device(self) -> None:
        steps = self._plan["step_dev"].tolist()
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is not None and len(self.state[p]) > 0:
                    self.state[p]["step"] = int(steps[gi])
        ops.parameters_updated_in_place(self._fresh_items())


class AdafactorScheduler(LambdaLR):
    """proxy scheduler of the reference (adafactor.py:249-284): reports the lr Adafactor computed itself."""

    def __init__(self, optimizer: Optimizer, initial_lr: float = 0.0):
        self.initial_lr = initial_lr
        for group in optimizer.param_groups:
            group["initial_lr"] = initial_lr
        super().__init__(optimizer, lambda _: self.initial_lr)
        for group in optimizer.param_groups:
            del group["initial_lr"]

    def get_lr(self):
        opt = self.optimizer
        lrs = [opt._get_lr(group, opt.state[group["params"][0]]) for group in opt.param_groups
               if group["params"][0].grad is not None and len(opt.state[group["params"][0]]) > 0]
        return lrs if lrs else self.base_lrs


class LitEma(nn.Module):
    """`LitEma` (reference modules/ema.py:11-97): shadow copies as buffers named by the parameter name with '.'
    removed, `decay = min(decay, (1 + n) / (10 + n))` warm-up, `shadow -= (1 - decay) * (shadow - p)` — as ONE
    multi-tensor launch over all trainable parameters."""

    SPAN = 1 << 18

    def __init__(self, model: nn.Module, decay: float = 0.9999, use_num_updates: bool = True):
        super().__init__()
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.m_name2s_name = {}
        self.register_buffer("decay", torch.tensor(decay, dtype=torch.float32))
        self.register_buffer("num_updates", torch.tensor(0 if use_num_updates else -1, dtype=torch.int))
        for name, p in model.named_parameters():
            if p.requires_grad:
                s_name = name.replace(".", "_")
                self.m_name2s_name[name] = s_name
                self.register_buffer(s_name, p.clone().detach().data)
        self.collected_params = []
        self._n_host = 0 if use_num_updates else -1
        self._decay_host = float(self.decay)  # the fp32 value the reference compares against
        self._table = None
        self._key = None
        # resume: `num_updates` / `decay` arrive through load_state_dict (engine.init_from_ckpt, Lightning checkpoints);
        # the reference reads the buffers directly (ema.py:44-46), so the host mirrors follow every load
        self.register_load_state_dict_post_hook(lambda module, _incompatible: module.sync_from_device())

    def reset_num_updates(self):
        del self.num_updates
        self.register_buffer("num_updates", torch.tensor(0, dtype=torch.int))
        self._n_host = 0

    def update(self, model: nn.Module):
        self.forward(model)

    def current_decay(self) -> float:
        """host mirror of ema.py:44-46 after the increment of num_updates."""
        if self._n_host >= 0:
            n = self._n_host
            return min(self._decay_host, (1 + n) / (10 + n))
        return self._decay_host

    def _prepare(self, model: nn.Module) -> None:
        shadow = dict(self.named_buffers())
        pairs = []
        for key, p in model.named_parameters():
            if p.requires_grad:
                pairs.append((shadow[self.m_name2s_name[key]], p))
            elif key in self.m_name2s_name:
                raise ValueError(f"Parameter {key} is not trainable, but has a shadow parameter")
        key = tuple((s.data_ptr(), p.data_ptr(), p.numel()) for s, p in pairs)
        if pairs and (self._table is None or key != self._key):
            for s, p in pairs:
                if not (s.is_cuda and p.is_cuda and s.dtype == p.dtype == torch.float32 and s.is_contiguous()
                        and p.is_contiguous()):
                    raise RuntimeError("neurosis_b200 LitEma needs contiguous CUDA fp32 parameters (no CPU fallback)")
            rows = [(sp + 4 * off, pp + 4 * off, min(self.SPAN, n - off)) for sp, pp, n in key
                    for off in range(0, n, self.SPAN)]
            dev = pairs[0][1].device
            self._table = torch.tensor(rows, dtype=torch.int64).to(dev)
            self._omd_host = torch.zeros(1, dtype=torch.float32).pin_memory()
            self._omd = torch.zeros(1, dtype=torch.float32, device=dev)
            self._key = key
        self._have = bool(pairs)

    def _launch(self) -> None:
        check(lib.nk_ema_update_multi(self._table.data_ptr(), self._table.shape[0], self._omd.data_ptr(),
                                      ops._stream()), "ema_update_multi")
        ops._count()

    @torch.no_grad()
    def forward(self, model: nn.Module):
        if self._n_host >= 0:
            self._n_host += 1
            self.num_updates += 1
        omd = 1.0 - float(np.float32(self.current_decay()))
        self._prepare(model)
        if not self._have:
            return
        self._omd_host[0] = omd
        self._omd.copy_(self._omd_host, non_blocking=True)
        self._launch()

    # ---- CUDA-graph form: the update counter is the `num_updates` buffer itself -------------------------------------
    def graph_prepare(self, model: nn.Module) -> None:
        """call once before capture.  The update counter and decay are created as CPU buffers (as in the reference,
        ema.py:20-21) and normally follow the owning module's `.to(device)`; the captured decay kernel dereferences
        `num_updates`, so it must live on the parameters' device."""
        self._prepare(model)
        if not self._have:
            raise RuntimeError("LitEma.graph_prepare: the model has no trainable parameters")
        dev = self._omd.device
        if self.num_updates.device != dev:
            self.num_updates = self.num_updates.to(dev)
            self.decay = self.decay.to(dev)

    @torch.no_grad()
    def graph_launch(self) -> None:
        if not (self.num_updates.is_cuda and self.num_updates.dtype == torch.int32):
            raise RuntimeError("LitEma.graph_launch: call graph_prepare(model) first (num_updates must be a CUDA int32 buffer)")
        check(lib.nk_ema_decay(self._decay_host, self.num_updates.data_ptr(), self._omd.data_ptr(), ops._stream()),
              "ema_decay")
        ops._count()
        self._launch()

    def sync_from_device(self) -> None:
        """host mirrors of the `num_updates` / `decay` buffers (after a checkpoint load or graph replays)."""
        self._n_host = int(self.num_updates)
        self._decay_host = float(self.decay)

    def copy_to(self, model: nn.Module):
        shadow = dict(self.named_buffers())
        changed = []
        for key, p in model.named_parameters():
            if p.requires_grad:
                p.data.copy_(shadow[self.m_name2s_name[key]].data)
                changed.append((p, False))
            elif key in self.m_name2s_name:
                raise ValueError(f"Parameter {key} is not trainable, but has a shadow parameter")
        ops.parameters_updated_in_place(changed)

    def store(self, parameters):
        self.collected_params = [param.clone() for param in parameters]

    def restore(self, parameters):
        changed = []
        for c_param, param in zip(self.collected_params, parameters):
            param.data.copy_(c_param.data)
            changed.append((param, False))
        ops.parameters_updated_in_place(changed)
