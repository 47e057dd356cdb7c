#!/usr/bin/env python
"""N-rank gradient / parameter equality on the REAL kernels over NCCL:

  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_equality.py

Every rank runs the TINY_SDXL diffusion step (VAE-free: latents given) on its own 2-sample batch with FIXED sigmas and
noise through (a) `BucketedGradReducer` (bucketed NCCL all-reduce overlapped with backward, weight gradients written
straight into the buckets on the side stream) and (b) `ShardedOptimizerReducer` + the fused Adafactor (reduce to the
bucket owner -> owner's step -> broadcast).  Rank 0 repeats the step single-process on the CONCATENATED batch with plain
autograd gradients and compares:
  * reduced gradients == single-process gradients (per-sample work is identical — GroupNorm is per sample — so only the
    fp32 reduction order and bf16 rounding flips differ);
  * every rank holds bit-identical reduced gradients;
  * parameters after the sharded optimizer step == parameters after a single-process Adafactor step.
Prints one JSON line and exits non-zero on mismatch."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def rel(a, b) -> float:
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def main() -> int:
    from common import TINY_SDXL
    from neurosis_b200 import ops
    from neurosis_b200.ddp import BucketedGradReducer, ShardedOptimizerReducer
    from neurosis_b200.modules import UNetModel
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import OpenAIWrapper, StandardDiffusionLoss
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    from neurosis_b200.optim import Adafactor
    from oracle.unet import unet_param_shapes
    from oracle.weights import synth_state_dict, synth_tensor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = TINY_SDXL
    sd = synth_state_dict(unet_param_shapes(cfg), seed=1)
    per = 2  # samples per rank

    def make_model():
        m = UNetModel(**cfg)
        m.load_state_dict(sd)
        return m.to(dev)

    den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()).to(dev)
    idx = torch.tensor([100 + 37 * i for i in range(world * per)])
    sig_all = den.sigmas[idx.to(dev)].float().cpu()

    def batch(r: int):
        return (synth_tensor(f"ddp.lat.{r}", (per, 4, 16, 16)), synth_tensor(f"ddp.noise.{r}", (per, 4, 16, 16)),
                {"crossattn": synth_tensor(f"ddp.ctx.{r}", (per, 77, cfg["context_dim"])),
                 "vector": synth_tensor(f"ddp.y.{r}", (per, cfg["adm_in_channels"]))}, sig_all[r * per: (r + 1) * per])

    class Fixed:
        def __init__(self, s):
            self.s = s

        def __call__(self, n, t=None):
            return self.s

    def loss_of(model, lat, noise, cond, sig):
        loss_fn = StandardDiffusionLoss(sigma_generator=Fixed(sig), loss_weighting=EpsWeighting())
        return loss_fn._forward(OpenAIWrapper(model), den, {k: v.to(dev) for k, v in cond.items()}, lat.to(dev), {},
                                noise=noise.to(dev)).mean()

    out = {"world": world}
    ok = True
    # ---- (a) bucketed all-reduce ------------------------------------------------------------------------------
    m = make_model()
    names = [n for n, _ in m.named_parameters()]
    red = BucketedGradReducer(list(m.parameters()), bucket_mb=2.0)
    red.attach_as_grad_sink()
    red.zero_grad()
    loss_of(m, *batch(rank)).backward()
    ops.join_side_streams()
    pending_before_finish = [b["pending"] for b in red.buckets]
    red.finish()
    torch.cuda.synchronize()
    red.detach_grad_sink()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    # every bucket must have been launched exactly when its last gradient arrived (0 left, none negative: a negative
    # count means a parameter was reported twice and the bucket was all-reduced before it was complete)
    out["pending_before_finish"] = pending_before_finish
    ok = ok and all(v == 0 for v in pending_before_finish)
    # identical on every rank: compare a checksum vector
    chk = torch.stack([g.double().sum() for g in grads.values()])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    out["reduced_grads_identical_on_all_ranks"] = bool(same)
    if not same:
        bad = [n for i, n in enumerate(names) if any(float(g[i]) != float(gathered[0][i]) for g in gathered)]
        out["params_that_differ_between_ranks"] = bad[:12]
        out["n_params_that_differ"] = len(bad)
    ok = ok and same
    if rank == 0:
        ref = make_model()
        parts = [batch(r) for r in range(world)]
        lat = torch.cat([p[0] for p in parts])
        noise = torch.cat([p[1] for p in parts])
        cond = {k: torch.cat([p[2][k] for p in parts]) for k in parts[0][2]}
        loss_of(ref, lat, noise, cond, sig_all).backward()
        torch.cuda.synchronize()
        errs = np.array([rel(grads[n], p.grad) for n, p in ref.named_parameters()])
        out["grad_rel_l2_median"] = float(np.median(errs))
        out["grad_rel_l2_max"] = float(errs.max())
        order = np.argsort(-errs)[:6]
        out["worst"] = [(names[i], float(errs[i])) for i in order]
        out["buckets"] = len(red.buckets)
        ok = ok and np.median(errs) < 2e-2 and errs.max() < 1.5e-1
        ref_grads = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
    # ---- (b) optimizer-state sharding: reduce -> owner's fused Adafactor step -> broadcast -------------------
    kw = dict(scale_parameter=True, relative_step=True, warmup_init=True)
    m2 = make_model()
    sh = ShardedOptimizerReducer(list(m2.parameters()), bucket_mb=2.0)
    sh.attach_as_grad_sink()
    opt = Adafactor(sh.owned_params(), **kw)
    for _ in range(2):
        ops.refresh_weight_copies(force=True)
        sh.zero_grad()
        loss_of(m2, *batch(rank)).backward()
        sh.finish()
        opt.step()
        sh.broadcast_params()
    torch.cuda.synchronize()
    sh.detach_grad_sink()
    pchk = torch.stack([p.detach().double().sum() for p in m2.parameters()])
    pg = [torch.zeros_like(pchk) for _ in range(world)]
    dist.all_gather(pg, pchk)
    same_p = all(torch.equal(pg[0], g) for g in pg)
    out["params_identical_on_all_ranks_after_sharded_step"] = bool(same_p)
    ok = ok and same_p
    if rank == 0:
        ref2 = make_model()
        opt2 = Adafactor(list(ref2.parameters()), **kw)
        for _ in range(2):
            ops.refresh_weight_copies(force=True)
            for p in ref2.parameters():
                p.grad = None
            loss_of(ref2, lat, noise, cond, sig_all).backward()
            opt2.step()
        torch.cuda.synchronize()
        w0 = make_model()
        upd = np.array([rel(p2.detach() - p0.detach(), pr.detach() - p0.detach())
                        for p2, pr, p0 in zip(m2.parameters(), ref2.parameters(), w0.parameters())])
        out["update_rel_l2_median"] = float(np.median(upd))
        out["update_rel_l2_max"] = float(upd.max())
        out["owned_params_rank0"] = len(sh.owned_params())
        ok = ok and np.median(upd) < 5e-2
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
