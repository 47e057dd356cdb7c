"""Tensor-level wrappers (and autograd Functions) over the C ABI of libnk_b200.so.

PyTorch is used only for device memory, streams and the autograd graph; every computation on
the hot path is a kernel of this repository.  All wrappers raise if the tensors are not on a
CUDA device — there is no CPU path.

Layout conventions
  * image activations: contiguous bf16 tensors of shape (N, H, W, C)  ("NHWC")
  * token activations: bf16 (..., C) with a contiguous last dim
  * parameters: the fp32 nn.Parameters of the reference layout; bf16 / packed copies for the
    kernels are cached per parameter version (`bf16_weight`, `packed_conv_weight`).
"""
from __future__ import annotations

import collections
import ctypes
import math
import weakref
from typing import Optional

import torch
from torch import Tensor

from ._lib import check, nk_gemm_desc
from ._lib import lib as _real_lib

BF16 = torch.bfloat16
F32 = torch.float32

LAUNCHES = 0  # number of kernel-launching C-ABI calls made (bench.py reports it)
FORCE_MATERIALIZED_ATTN_BWD = False  # tests: exercise the batched-GEMM attention backward for head_dim 64 too
PROFILE_GEMM = None  # bench.py sets this to a list: (start_event, end_event, algorithmic_flops) per tensor-core launch


PROFILE_KERNELS = None  # bench.py --breakdown: list of (entry point, start_event, end_event) for every C-ABI call


class _LibProxy:
    """forwards to the ctypes library; when PROFILE_KERNELS is a list every call is bracketed by CUDA events."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        fn = getattr(self._real, name)

        def call(*args):
            prof = PROFILE_KERNELS
            if prof is None:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            prof.append((name, e0, e1))
            return rc

        self.__dict__[name] = call
        return call


lib = _LibProxy(_real_lib)


GRAD_SINK = None  # set by neurosis_b200.ddp.BucketedGradReducer: weight gradients are written straight into its buckets


def _grad_sink(weight: Tensor):
    """(fp32 buffer shaped like `weight`, base parameter) if a reducer owns this parameter's gradient storage."""
    sink = GRAD_SINK
    if sink is None:
        return None, None
    base = weight._base if weight._base is not None else weight
    buf = sink.buffer_for(base)
    if buf is None:
        return None, None
    return buf.view(weight.shape), base


def _grad_sink_pair(a: Optional[Tensor], b: Optional[Tensor]):
    """sinks of two parameters at once (norm scale + shift): ((buf_a, buf_b), (base_a, base_b)) or (None, None)."""
    if a is None or b is None:
        return None, None
    ba, pa = _grad_sink(a)
    bb, pb = _grad_sink(b)
    if ba is None or bb is None:
        return None, None
    return (ba, bb), (pa, pb)


FLASH_HEAD_DIMS = (64, 128)  # head dims routed to the fused attention forward (tests widen this to exercise 64..512)
# ---- weight gradients on a side stream ---------------------------------------------------------------------------
# dW (and the conv weight-gradient chain) of a layer depends only on tensors that exist when its backward starts and
# nothing downstream in the backward pass reads it, so it runs on a second stream next to the data-gradient chain.
# The persistent GEMM kernels leave part of the SMs idle in their last wave; CTAs of the other stream's kernel take
# those SMs, and memory-bound kernels of the main chain overlap tensor-bound weight-gradient GEMMs.
# Only used when the gradient goes to a sink (ddp.BucketedGradReducer): autograd never sees the tensor, and the
# parameter is reported ready (mark_ready -> bucket all-reduce) only after the main stream has joined the side work.
import os as _os

WGRAD_OVERLAP = not _os.environ.get("NK_NO_WGRAD_OVERLAP")
_side_streams: dict = {}
_inflight: collections.deque = collections.deque()  # (event on the side stream, tensors kept alive, sink base params)


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=device)
    return st


def _retire_one(main) -> None:
    ev, _keep, bases = _inflight.popleft()
    main.wait_event(ev)  # from here on the main stream may reuse the kept tensors' memory
    sink = GRAD_SINK
    if sink is not None:
        for b in bases:
            sink.mark_ready(b)


def _fork_wgrad(fn, keep: tuple, bases: tuple) -> None:
    """run `fn` (kernel launches writing into gradient sinks) on the side stream, ordered after everything queued on
    the current stream so far.  `keep` holds the input tensors until the main stream has waited for the side work
    (the caching allocator must not hand their memory to a later main-stream kernel before that)."""
    main = torch.cuda.current_stream()
    side = _side_stream(main.device)
    side.wait_event(main.record_event())
    with torch.cuda.stream(side):
        fn()
        ev = side.record_event()
    _inflight.append((ev, keep, bases))
    while len(_inflight) > 2:  # at most two weight-gradient jobs trail the main stream
        _retire_one(main)


def join_side_streams() -> None:
    """main stream waits for all outstanding side-stream weight gradients (end of backward / before all-reduce)."""
    if not _inflight:
        return
    main = torch.cuda.current_stream()
    while _inflight:
        _retire_one(main)


NCU_SAMPLE = 0          # > 0: every NCU_SAMPLE-th tensor-core launch runs inside a cudaProfilerStart/Stop range
NCU_SAMPLE_LOG: list = []  # (what, small integer arguments) of the sampled launches, in launch order
_ncu_seen = 0


def _tc(rc_fn, what: str, flops: float, *args) -> None:
    """call a tensor-core GEMM entry point; optionally bracket it with CUDA events for the roofline report, or (for
    `ncu --profile-from-start off`) put a strided sample of the launches inside profiler ranges."""
    global _ncu_seen
    if NCU_SAMPLE > 0:
        _ncu_seen += 1
        if _ncu_seen % NCU_SAMPLE == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            check(rc_fn(*args), what)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            NCU_SAMPLE_LOG.append((what, flops, [a for a in args if isinstance(a, int) and 0 < a < (1 << 24)]))
            return
    if PROFILE_GEMM is None:
        check(rc_fn(*args), what)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(rc_fn(*args), what)
    e1.record()
    PROFILE_GEMM.append((e0, e1, flops, what, tuple(a for a in args if isinstance(a, int) and 0 < a < (1 << 24))))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _req_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("neurosis_b200 ops need CUDA tensors (there is no CPU fallback)")


# --------------------------------------------------------------------------------------------
# parameter copies for the kernels
# --------------------------------------------------------------------------------------------
_wcache: dict[int, tuple] = {}  # id(param) -> (weakref(param), cache dict)


def _cache_for(p: Tensor) -> dict:
    """per-parameter cache of kernel-side copies, invalidated by in-place updates (optimizer steps)."""
    base = p._base if p._base is not None else p
    key = id(base)
    ent = _wcache.get(key)
    if ent is not None and ent[0]() is base:
        d = ent[1]
        if d["version"] == base._version and d["ptr"] == base.data_ptr():
            return d
    d = {"version": base._version, "ptr": base.data_ptr()}
    _wcache[key] = (weakref.ref(base, lambda _r, k=key: _wcache.pop(k, None)), d)
    return d


def invalidate_weight_cache() -> None:
    """drop every cached bf16 / packed weight copy (they are re-derived by kernels on next use).  A training loop
    whose optimizer updates the fp32 parameters in place gets this for free through the version check; callers
    that capture CUDA graphs call it before capture so the casts are part of the graph and replayed every step."""
    _wcache.clear()
    _mirror.clear()
    _wgroups.clear()


class _WeightMirror:
    """Persistent bf16 copies of every parameter that `bf16_weight` has served, refreshed by ONE kernel
    (nk_cast_f32_bf16_multi) instead of one cast launch per parameter: ~750 linear weights per SDXL step."""
    SPAN = 1 << 18  # elements per thread block

    def __init__(self):
        self.clear()

    def clear(self) -> None:
        self.params: dict[int, tuple] = {}  # id(base) -> (weakref(base), bf16 copy)
        self.table = None
        self.table_key = None

    def register(self, base: Tensor, copy: Tensor) -> None:
        self.params[id(base)] = (weakref.ref(base), copy)
        self.table = None

    def refresh(self) -> int:
        """re-derive every registered copy from the current fp32 values; returns the number of parameters."""
        live = [(r(), c) for r, c in self.params.values() if r() is not None]
        if not live:
            return 0
        key = tuple((b.data_ptr(), c.data_ptr(), b.numel()) for b, c in live)
        if self.table is None or self.table_key != key:
            rows = []
            for src, dst, n in key:
                for off in range(0, n, self.SPAN):
                    rows.append((src + 4 * off, dst + 2 * off, min(self.SPAN, n - off)))
            self.table = torch.tensor(rows, dtype=torch.int64).to(live[0][0].device)
            self.table_key = key
        check(lib.nk_cast_f32_bf16_multi(self.table.data_ptr(), self.table.shape[0], _stream()), "cast_f32_bf16_multi")
        _count()
        for b, c in live:
            d = _cache_for(b)  # a fresh dict when the parameter changed since the last copy
            d["bf16"] = c
        return len(live)


_mirror = _WeightMirror()


def refresh_weight_copies(force: bool = False) -> None:
    """Start-of-step hook of a training loop: re-derive the kernels' bf16 weight copies after the optimizer updated
    the fp32 parameters.  Linear weights go through one multi-tensor launch; packed convolution weights are re-packed
    lazily by their first use.  `force` treats every parameter as changed (benchmarks without an optimizer)."""
    if force:
        for ent in list(_wcache.values()):
            ent[1].pop("packed", None)
            ent[1].pop("bf16", None)
            ent[1].pop("patch3x3", None)
    _mirror.refresh()


def registered_mirror(p: Tensor) -> Optional[Tensor]:
    """the persistent bf16 copy of a parameter that `bf16_weight` serves to the GEMMs, if one exists (the fused
    optimizer step rewrites it in its apply pass instead of leaving it to the per-step refresh)."""
    base = p._base if p._base is not None else p
    ent = _mirror.params.get(id(base))
    if ent is None or ent[0]() is not base:
        return None
    c = ent[1]
    return c if c.is_contiguous() and c.numel() == base.numel() else None


def parameters_updated_in_place(items) -> None:
    """`items` = [(parameter, mirror_is_fresh)]: a kernel of this library wrote the fp32 values through raw pointers.
    Bumps the autograd version counters (the per-parameter caches of packed / bf16 copies key on them) and, where the
    same kernel also rewrote the bf16 mirror, re-attaches that copy to the new version so it is not cast again."""
    for p, fresh in items:
        base = p._base if p._base is not None else p
        torch.autograd.graph.increment_version(base)
        if fresh:
            ent = _mirror.params.get(id(base))
            if ent is not None and ent[0]() is base:
                _cache_for(base)["bf16"] = ent[1]


def cast_bf16(x: Tensor) -> Tensor:
    """fp32 -> bf16 copy with our kernel (bf16 input is returned as is)."""
    if x.dtype == BF16:
        return x.contiguous()
    _req_cuda(x)
    x = x.contiguous().float()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(lib.nk_cast_f32_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "cast_f32_bf16")
    _count()
    return y


def bf16_weight(p: Tensor) -> Tensor:
    d = _cache_for(p)
    w = d.get("bf16")
    if w is None:
        base = p._base if p._base is not None else p
        w = cast_bf16(base.detach())
        d["bf16"] = w
        if base.dtype == F32 and base.is_contiguous():
            _mirror.register(base, w)
    return w.view(p.shape) if w.shape != p.shape else w


_wgroups: dict = {}  # tuple(id(base)) -> (weakrefs, stacked bf16 buffer, per-parameter views)


def bf16_weight_group(ps: tuple) -> Tensor:
    """bf16 copies of several [N_i, K] weights stacked in ONE [sum N_i, K] buffer (fused projections: q|k|v).  Every
    parameter's slice is an ordinary cached copy (refreshed by `refresh_weight_copies`), so the stack stays valid."""
    bases = tuple(p._base if p._base is not None else p for p in ps)
    key = tuple(id(b) for b in bases)
    ent = _wgroups.get(key)
    if ent is None or any(r() is not b for r, b in zip(ent[0], bases)):
        k = bases[0].shape[1]
        assert all(b.dim() == 2 and b.shape[1] == k and b.dtype == F32 and b.is_contiguous() for b in bases)
        buf = torch.empty((sum(b.shape[0] for b in bases), k), dtype=BF16, device=bases[0].device)
        views, r0 = [], 0
        for b in bases:
            views.append(buf[r0: r0 + b.shape[0]])
            r0 += b.shape[0]
        ent = (tuple(weakref.ref(b) for b in bases), buf, views)
        _wgroups[key] = ent
    for b, view in zip(bases, ent[2]):
        d = _cache_for(b)
        if d.get("bf16") is not view:  # stale (parameter changed) or cached elsewhere: cast into the stacked slice
            check(lib.nk_cast_f32_bf16(b.detach().data_ptr(), view.data_ptr(), b.numel(), _stream()), "cast_f32_bf16")
            _count()
            d["bf16"] = view
            _mirror.register(b, view)
    return ent[1]


def f32_param(p: Optional[Tensor]) -> Optional[Tensor]:
    if p is None:
        return None
    t = p.detach()
    return t if t.dtype == F32 and t.is_contiguous() else t.float().contiguous()


def _pad64(c: int) -> int:
    return c if c >= 64 else 64


def packed_conv_weight(p: Tensor, need_dgrad: bool = True):
    """(wp_fwd [CoP, taps*CiP], wp_dgrad [CiP, taps*CoP]) bf16 packings of an OIHW fp32 weight."""
    d = _cache_for(p)
    pk = d.get("packed")
    if pk is None:
        co, ci, ks, _ = p.shape
        cop, cip = _pad64(co), _pad64(ci)
        w = f32_param(p)
        wf = torch.empty((cop, ks * ks * cip), dtype=BF16, device=p.device)
        wd = torch.empty((cip, ks * ks * cop), dtype=BF16, device=p.device)
        check(lib.nk_conv_pack_weights(w.data_ptr(), wf.data_ptr(), wd.data_ptr(), co, ci, ks, cop, cip, _stream()),
              "conv_pack_weights")
        _count()
        pk = (wf, wd)
        d["packed"] = pk
    return pk


# --------------------------------------------------------------------------------------------
# raw (non-autograd) kernels
# --------------------------------------------------------------------------------------------
def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
               out_f32: bool = False) -> Tensor:
    """y[M,N] = x[M,K] @ w[N,K]^T + bias + residual;  x, w, residual bf16; bias fp32."""
    _req_cuda(x, w)
    K = x.shape[-1]
    N = w.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty((M, N), dtype=F32 if out_f32 else BF16, device=x.device)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    _tc(lib.nk_linear_fwd, "linear_fwd", 2.0 * M * N * K, x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0),
        _p(bias), _p(r2), r2.stride(0) if r2 is not None else 0, y.data_ptr(), N, int(out_f32), M, N, K, _stream())
    _count()
    return y.view(*x.shape[:-1], N)


def linear_dgrad(dy: Tensor, w: Tensor, residual: Optional[Tensor] = None) -> Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K] (+ residual)."""
    N, K = w.shape
    d2 = dy.reshape(-1, N)
    if d2.stride(-1) != 1:
        d2 = d2.contiguous()
    M = d2.shape[0]
    dx = torch.empty((M, K), dtype=BF16, device=dy.device)
    r2 = residual.reshape(-1, K) if residual is not None else None
    _tc(lib.nk_linear_dgrad, "linear_dgrad", 2.0 * M * N * K, d2.data_ptr(), d2.stride(0), w.data_ptr(), w.stride(0),
        _p(r2), r2.stride(0) if r2 is not None else 0, dx.data_ptr(), K, M, N, K, _stream())
    _count()
    return dx.view(*dy.shape[:-1], K)


def linear_wgrad(dy: Tensor, x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """dw[N,K] fp32 = dy[M,N]^T @ x[M,K]; with `out` given the product is accumulated into it."""
    N = dy.shape[-1]
    K = x.shape[-1]
    d2 = dy.reshape(-1, N)
    x2 = x.reshape(-1, K)
    if d2.stride(-1) != 1:
        d2 = d2.contiguous()
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    M = d2.shape[0]
    acc = out is not None
    dw = out if acc else torch.empty((N, K), dtype=F32, device=dy.device)
    _tc(lib.nk_linear_wgrad, "linear_wgrad", 2.0 * M * N * K, d2.data_ptr(), d2.stride(0), x2.data_ptr(), x2.stride(0),
        dw.data_ptr(), dw.stride(0), int(acc), M, N, K, _stream())
    _count()
    return dw


def linear_geglu_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], keep_h: bool = True):
    """(h [M, 2D] or None, out [M, D]): h = x @ w[2D, K]^T + bias, out = h[:, :D] * gelu(h[:, D:]) — the gate is applied
    in the GEMM epilogue (one N tile holds the value and the gate columns of the same features)."""
    _req_cuda(x, w)
    K = x.shape[-1]
    D = w.shape[0] // 2
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    h = torch.empty((M, 2 * D), dtype=BF16, device=x.device) if keep_h else None
    out = torch.empty((M, D), dtype=BF16, device=x.device)
    _tc(lib.nk_linear_geglu_fwd, "linear_geglu_fwd", 2.0 * M * 2 * D * K, x2.data_ptr(), x2.stride(0), w.data_ptr(),
        w.stride(0), _p(bias), _p(h), 2 * D, out.data_ptr(), D, M, D, K, _stream())
    _count()
    return (h.view(*x.shape[:-1], 2 * D) if h is not None else None), out.view(*x.shape[:-1], D)


def linear_dgrad_geglu(dy: Tensor, w: Tensor, h: Tensor) -> Tensor:
    """dh [M, 2D] of a GEGLU whose output feeds y = out @ w[N, D]^T: d_out = dy @ w goes through the gate's derivative in
    the epilogue of the data-gradient GEMM and is never written."""
    N, D = w.shape
    d2 = dy.reshape(-1, N)
    if d2.stride(-1) != 1:
        d2 = d2.contiguous()
    h2 = h.reshape(-1, 2 * D)
    M = d2.shape[0]
    dh = torch.empty((M, 2 * D), dtype=BF16, device=dy.device)
    _tc(lib.nk_linear_dgrad_geglu, "linear_dgrad_geglu", 2.0 * M * N * D, d2.data_ptr(), d2.stride(0), w.data_ptr(),
        w.stride(0), h2.data_ptr(), h2.stride(0), dh.data_ptr(), 2 * D, M, N, D, _stream())
    _count()
    return dh.view(*dy.shape[:-1], 2 * D)


def colsum(x: Tensor, groups: int = 1, out: Optional[Tensor] = None) -> Tensor:
    """fp32 [groups, C] column sums of a bf16 [groups*rows, C] matrix; with `out` given the sums are added to it."""
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    rows = x2.shape[0] // groups
    acc = out is not None
    if not acc:
        out = torch.empty((groups, C), dtype=F32, device=x.device)
    check(lib.nk_colsum(x2.data_ptr(), x2.stride(0), out.data_ptr(), groups, rows, C, int(acc), _stream()), "colsum")
    _count()
    return out


def conv2d_fwd(x: Tensor, wp: Tensor, cout: int, ksize: int, bias: Optional[Tensor] = None,
               bias_img: Optional[Tensor] = None, residual: Optional[Tensor] = None) -> Tensor:
    """x: (N,H,W,Cin) bf16 with Cin % 64 == 0; wp packed [CoP, taps*Cin]; returns (N,H,W,CoP) with
    the first `cout` channels written (CoP = max(cout, 64); padded channels are zero)."""
    _req_cuda(x, wp)
    n, h, w_, cin = x.shape
    cop = _pad64(cout)
    if cop != cout:
        y = torch.zeros((n, h, w_, cop), dtype=BF16, device=x.device)
    else:
        y = torch.empty((n, h, w_, cop), dtype=BF16, device=x.device)
    _tc(lib.nk_conv2d_fwd, "conv2d_fwd", 2.0 * n * h * w_ * cout * ksize * ksize * cin, x.data_ptr(), x.stride(2),
        wp.data_ptr(), _p(bias), _p(bias_img), _p(residual), residual.stride(2) if residual is not None else 0,
        y.data_ptr(), cop, n, h, w_, cin, cout, ksize, _stream())
    _count()
    return y


def conv3x3_thin_input_fwd(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """First convolution of the VAE encoder (RGB -> 128 channels, 3x3 pad 1), forward only: x (N,C,H,W) fp32 with
    C in {1,3,4}; the 9*C patch values of every pixel become one 64-deep GEMM row.  Returns (N,H,W,Cout) bf16."""
    _req_cuda(x)
    n, c, h, w_ = x.shape
    co = weight.shape[0]
    x = x.float().contiguous()
    col = torch.empty((n * h * w_, 64), dtype=BF16, device=x.device)
    check(lib.nk_image_patches3x3(x.data_ptr(), col.data_ptr(), n, c, h, w_, _stream()), "image_patches3x3")
    _count()
    d = _cache_for(weight)
    wp = d.get("patch3x3")
    if wp is None:  # [Cout, 64] bf16, column = tap*C + c (tiny: host-side torch glue, cached per parameter version)
        w2 = weight.detach().float().permute(0, 2, 3, 1).reshape(co, 9 * c)
        wp = torch.nn.functional.pad(w2, (0, 64 - 9 * c)).to(BF16).contiguous()
        d["patch3x3"] = wp
    return linear_fwd(col, wp, f32_param(bias)).view(n, h, w_, co)


def conv2d_stride2_fwd(x: Tensor, wp: Tensor, cout: int, ksize: int, bias: Optional[Tensor], pad_t: int, pad_l: int,
                       ho: int, wo: int) -> Tensor:
    """3x3 stride-2 convolution of NHWC bf16 x (physical channels % 64 == 0) with packed weights -> (n, ho, wo, cout)."""
    _req_cuda(x, wp)
    n, h, w_, cin = x.shape
    y = torch.empty((n, ho, wo, cout), dtype=BF16, device=x.device)
    _tc(lib.nk_conv2d_stride2_fwd, "conv2d_s2_fwd", 2.0 * n * ho * wo * cout * ksize * ksize * cin, x.data_ptr(),
        x.stride(2), wp.data_ptr(), _p(bias), y.data_ptr(), cout, n, h, w_, cin, cout, ksize, pad_t, pad_l, ho, wo,
        _stream())
    _count()
    return y


def conv2d_wgrad(dy: Tensor, x: Tensor, cout: int, ksize: int) -> Tensor:
    """packed fp32 gradient [cout, taps, Cin_phys]."""
    n, h, w_, cin = x.shape
    dwp = torch.zeros((cout, ksize * ksize, cin), dtype=F32, device=x.device)
    _tc(lib.nk_conv2d_wgrad, "conv2d_wgrad", 2.0 * n * h * w_ * cout * ksize * ksize * cin, dy.data_ptr(), dy.stride(2),
        x.data_ptr(), x.stride(2), dwp.data_ptr(), n, h, w_, cin, cout, ksize, _stream())
    _count()
    return dwp


def conv_unpack_wgrad(dwp: Tensor, co: int, ci: int, ks: int, out: Optional[Tensor] = None) -> Tensor:
    """packed [co, taps, cip] fp32 -> OIHW fp32; with `out` given the values are ADDED into it."""
    cip = dwp.shape[-1] if dwp.dim() == 3 else dwp.shape[-1] // (ks * ks)
    dw = out if out is not None else torch.empty((co, ci, ks, ks), dtype=F32, device=dwp.device)
    check(lib.nk_conv_unpack_wgrad(dwp.data_ptr(), dw.data_ptr(), co, ci, ks, cip, int(out is not None), _stream()),
          "conv_unpack_wgrad")
    _count()
    return dw


def im2col(x: Tensor, ks: int, stride: int, pad_t: int, pad_l: int, ho: int, wo: int) -> Tensor:
    n, h, w_, c = x.shape
    col = torch.empty((n * ho * wo, ks * ks * c), dtype=BF16, device=x.device)
    check(lib.nk_im2col(x.data_ptr(), x.stride(2), col.data_ptr(), n, h, w_, c, ks, stride, pad_t, pad_l, ho, wo,
                        _stream()), "im2col")
    _count()
    return col


def col2im(dcol: Tensor, shape, ks: int, stride: int, pad_t: int, pad_l: int, ho: int, wo: int) -> Tensor:
    n, h, w_, c = shape
    dx = torch.empty((n, h, w_, c), dtype=BF16, device=dcol.device)
    check(lib.nk_col2im(dcol.data_ptr(), dx.data_ptr(), n, h, w_, c, ks, stride, pad_t, pad_l, ho, wo, _stream()),
          "col2im")
    _count()
    return dx


def groupnorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float, silu: bool):
    _req_cuda(x)
    n, h, w_, c = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((n, groups), dtype=F32, device=x.device)
    rstd = torch.empty((n, groups), dtype=F32, device=x.device)
    ws_bytes = lib.nk_groupnorm_workspace_bytes(n, h * w_, c, groups)
    if ws_bytes < 0:
        raise ValueError(f"groupnorm: unsupported channel count {c}")
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    check(lib.nk_groupnorm_fwd(x.data_ptr(), x.stride(2), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), c,
                               mean.data_ptr(), rstd.data_ptr(), ws.data_ptr(), ws_bytes, n, h * w_, c, groups,
                               float(eps), int(silu), _stream()), "groupnorm_fwd")
    _count(3)
    return y, mean, rstd


def groupnorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, beta: Tensor, mean: Tensor, rstd: Tensor, groups: int,
                  silu: bool, out: Optional[tuple] = None):
    """`out` = (dgamma, dbeta) fp32 buffers the parameter gradients are ADDED to (gradient buckets)."""
    n, h, w_, c = x.shape
    dx = torch.empty_like(x)
    if out is not None:
        dgamma, dbeta = out
    else:
        dgamma = torch.zeros((c,), dtype=F32, device=x.device)
        dbeta = torch.zeros((c,), dtype=F32, device=x.device)
    ws_bytes = lib.nk_groupnorm_workspace_bytes(n, h * w_, c, groups)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    check(lib.nk_groupnorm_bwd(dy.data_ptr(), dy.stride(2), x.data_ptr(), x.stride(2), gamma.data_ptr(),
                               beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), c, dgamma.data_ptr(),
                               dbeta.data_ptr(), ws.data_ptr(), ws_bytes, n, h * w_, c, groups, int(silu), _stream()),
          "groupnorm_bwd")
    _count(4)
    return dx, dgamma, dbeta


def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float):
    _req_cuda(x)
    c = x.shape[-1]
    x2 = x.reshape(-1, c)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    rows = x2.shape[0]
    y = torch.empty((rows, c), dtype=BF16, device=x.device)
    mean = torch.empty((rows,), dtype=F32, device=x.device)
    rstd = torch.empty((rows,), dtype=F32, device=x.device)
    check(lib.nk_layernorm_fwd(x2.data_ptr(), x2.stride(0), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), c,
                               mean.data_ptr(), rstd.data_ptr(), rows, c, float(eps), _stream()), "layernorm_fwd")
    _count()
    return y.view(x.shape), mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, out: Optional[tuple] = None,
                  dres: Optional[Tensor] = None):
    """`out` = (dgamma, dbeta) fp32 buffers the parameter gradients are ADDED to (gradient buckets); `dres` (bf16, shape
    of x) is a gradient that bypasses the norm through a residual connection and is added into dx."""
    c = x.shape[-1]
    x2 = x.reshape(-1, c)
    d2 = dy.reshape(-1, c)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if d2.stride(-1) != 1:
        d2 = d2.contiguous()
    rows = x2.shape[0]
    dx = torch.empty((rows, c), dtype=BF16, device=x.device)
    if out is not None:
        dgamma, dbeta = out
    else:
        dgamma = torch.zeros((c,), dtype=F32, device=x.device)
        dbeta = torch.zeros((c,), dtype=F32, device=x.device)
    r2 = None
    if dres is not None:
        r2 = dres.reshape(-1, c)
        if r2.dtype != BF16:
            r2 = cast_bf16(r2)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    check(lib.nk_layernorm_bwd(d2.data_ptr(), d2.stride(0), x2.data_ptr(), x2.stride(0), gamma.data_ptr(),
                               mean.data_ptr(), rstd.data_ptr(), _p(r2), r2.stride(0) if r2 is not None else 0,
                               dx.data_ptr(), c, dgamma.data_ptr(), dbeta.data_ptr(), rows, c, _stream()),
          "layernorm_bwd")
    _count(2)
    return dx.view(x.shape), dgamma, dbeta


def geglu_fwd(h: Tensor) -> Tensor:
    d2 = h.shape[-1]
    d = d2 // 2
    h2 = h.reshape(-1, d2)
    out = torch.empty((h2.shape[0], d), dtype=BF16, device=h.device)
    check(lib.nk_geglu_fwd(h2.data_ptr(), h2.stride(0), out.data_ptr(), d, h2.shape[0], d, _stream()), "geglu_fwd")
    _count()
    return out.view(*h.shape[:-1], d)


def geglu_bwd(h: Tensor, dout: Tensor) -> Tensor:
    d2 = h.shape[-1]
    d = d2 // 2
    h2 = h.reshape(-1, d2)
    do2 = dout.reshape(-1, d)
    if do2.stride(-1) != 1:
        do2 = do2.contiguous()
    dh = torch.empty_like(h2)
    check(lib.nk_geglu_bwd(h2.data_ptr(), h2.stride(0), do2.data_ptr(), do2.stride(0), dh.data_ptr(), d2, h2.shape[0],
                           d, _stream()), "geglu_bwd")
    _count()
    return dh.view(h.shape)


def silu_fwd(x: Tensor) -> Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib.nk_silu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "silu_fwd")
    _count()
    return y


def silu_bwd(x: Tensor, dy: Tensor) -> Tensor:
    x = x.contiguous()
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    check(lib.nk_silu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _stream()), "silu_bwd")
    _count()
    return dx


def add(a: Tensor, b: Tensor) -> Tensor:
    a = a.contiguous()
    b = b.contiguous()
    y = torch.empty_like(a)
    check(lib.nk_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), "add")
    _count()
    return y


def copy_channels(src: Tensor, dst: Tensor, c: int) -> None:
    """copy the first c channels of every pixel of src (.., Cs) into dst (.., Cd) (views allowed)."""
    npix = src.numel() // src.shape[-1] if src.is_contiguous() else math.prod(src.shape[:-1])
    check(lib.nk_copy_channels(src.data_ptr(), src.stride(-2), dst.data_ptr(), dst.stride(-2), npix, c, _stream()),
          "copy_channels")
    _count()


def cat_channels(a: Tensor, b: Tensor) -> Tensor:
    ca, cb = a.shape[-1], b.shape[-1]
    out = torch.empty((*a.shape[:-1], ca + cb), dtype=BF16, device=a.device)
    copy_channels(a, out, ca)
    copy_channels(b, out[..., ca:], cb)
    return out


def upsample2x_fwd(x: Tensor) -> Tensor:
    n, h, w_, c = x.shape
    y = torch.empty((n, 2 * h, 2 * w_, c), dtype=BF16, device=x.device)
    check(lib.nk_upsample2x_fwd(x.data_ptr(), y.data_ptr(), n, h, w_, c, _stream()), "upsample2x_fwd")
    _count()
    return y


def upsample2x_bwd(dy: Tensor) -> Tensor:
    n, h2, w2, c = dy.shape
    dx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=BF16, device=dy.device)
    check(lib.nk_upsample2x_bwd(dy.data_ptr(), dx.data_ptr(), n, h2 // 2, w2 // 2, c, _stream()), "upsample2x_bwd")
    _count()
    return dx


def nchw_to_nhwc(x: Tensor, cpad: Optional[int] = None, scale: Optional[Tensor] = None) -> Tensor:
    """(N,C,H,W) fp32|bf16 contiguous -> (N,H,W,Cpad) bf16, zero padded channels, optional per-image scale."""
    _req_cuda(x)
    if x.dtype not in (F32, BF16):
        x = x.float()
    x = x.contiguous()
    n, c, h, w_ = x.shape
    cpad = cpad or c
    y = torch.empty((n, h, w_, cpad), dtype=BF16, device=x.device)
    check(lib.nk_nchw_to_nhwc(x.data_ptr(), int(x.dtype == F32), y.data_ptr(), _p(scale), n, c, h * w_, cpad,
                              _stream()), "nchw_to_nhwc")
    _count()
    return y


def nhwc_to_nchw(x: Tensor, c: Optional[int] = None, out_f32: bool = True) -> Tensor:
    n, h, w_, cp = x.shape
    c = c or cp
    y = torch.empty((n, c, h, w_), dtype=F32 if out_f32 else BF16, device=x.device)
    check(lib.nk_nhwc_to_nchw(x.data_ptr(), x.stride(2), y.data_ptr(), int(out_f32), n, c, h * w_, _stream()),
          "nhwc_to_nchw")
    _count()
    return y


def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    _req_cuda(t)
    tf = t.float().contiguous()
    out = torch.empty((tf.shape[0], dim), dtype=BF16, device=t.device)
    check(lib.nk_timestep_embedding(tf.data_ptr(), out.data_ptr(), tf.shape[0], dim, float(max_period), _stream()),
          "timestep_embedding")
    _count()
    return out


# ---- generic batched GEMM descriptor helpers (attention) --------------------------------------
def _operand(o, t: Tensor, mn_major: int, rows: int, inner: int, row_stride: int, nb2: int, b2_stride: int, nb1: int,
             b1_stride: int) -> None:
    o.ptr = t.data_ptr()
    o.mn_major = mn_major
    o.conv = 0
    o.inner = inner
    o.rows = rows
    o.row_stride = row_stride
    o.nb2 = nb2
    o.b2_stride = b2_stride
    o.nb1 = nb1
    o.b1_stride = b1_stride


def _bhnd(t: Tensor):
    """strides (batch, row, head) in elements of a (B, N, H, D) view with contiguous D."""
    assert t.stride(3) == 1
    return t.stride(0), t.stride(1), t.stride(2)


def attention_fwd(q: Tensor, k: Tensor, v: Tensor, scale: float):
    """q (B,Nq,H,D), k/v (B,Nk,H,D) bf16 views (last dim contiguous) -> o (B,Nq,H,D), lse (B,H,Nq)."""
    _req_cuda(q, k, v)
    B, Nq, H, D = q.shape
    Nk = k.shape[1]
    o = torch.empty((B, Nq, H, D), dtype=BF16, device=q.device)
    lse = torch.empty((B, H, Nq), dtype=F32, device=q.device)
    packed = all(t.stride(3) == 1 and (H == 1 or t.stride(2) == D) for t in (q, k, v))
    # flash kernel: D = 64, or D = 128 as two 64-column chunks.  The kernel accepts up to D = 512, but every V/O chunk
    # recomputes the full score tile, so (D/64 + 1)/2 x the useful tensor work: measured on the VAE mid block
    # (1 x 512, 16384 tokens) it is 52 ms against 19.5 ms for the materialised path below, which therefore stays.
    if (D in FLASH_HEAD_DIMS or (D < 64 and D % 8 == 0)) and packed:
        check(lib.nk_attention_fwd(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                   v.data_ptr(), v.stride(1), v.stride(0), o.data_ptr(), o.stride(1), o.stride(0),
                                   lse.data_ptr(), B, H, Nq, Nk, D, float(scale), _stream()), "attention_fwd")
        _count()
        return o, lse
    # materialised path: S = Q K^T (fp32) -> row softmax -> O = P V
    S = torch.empty((B, H, Nq, Nk), dtype=F32, device=q.device)
    d = nk_gemm_desc()
    qb, qr, qh = _bhnd(q)
    kb, kr, kh = _bhnd(k)
    _operand(d.A, q, 0, Nq, D, qr, H, qh, B, qb)
    _operand(d.B, k, 0, Nk, D, kr, H, kh, B, kb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nq, Nk, D, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = S.data_ptr(), Nk, Nq * Nk, H * Nq * Nk
    d.out, d.epi, d.alpha = 1, 0, 1.0
    _tc(lib.nk_gemm_ex, "attention S gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), _stream())
    Nkp = (Nk + 7) // 8 * 8
    P = torch.empty((B, H, Nq, Nkp), dtype=BF16, device=q.device) if Nkp == Nk else torch.zeros(
        (B, H, Nq, Nkp), dtype=BF16, device=q.device)
    check(lib.nk_softmax_rows(S.data_ptr(), Nk, P.data_ptr(), Nkp, lse.data_ptr(), B * H * Nq, Nk, float(scale),
                              _stream()), "softmax_rows")
    del S
    _pv(P, v, o, Nk)
    _count(3)
    return o, lse


def _pv(P: Tensor, v: Tensor, o: Tensor, Nk: int) -> None:
    """o[b,:,h,:] = P[b,h] @ v[b,:,h,:]  (B operand = V is MN-major: head dim contiguous)."""
    B, H, Nq, Nkp = P.shape
    D = v.shape[-1]
    d = nk_gemm_desc()
    vb, vr, vh = _bhnd(v)
    _operand(d.A, P, 0, Nq, Nk, Nkp, H, Nq * Nkp, B, H * Nq * Nkp)
    _operand(d.B, v, 1, Nk, D, vr, H, vh, B, vb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nq, D, Nk, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = o.data_ptr(), o.stride(1), o.stride(2), o.stride(0)
    d.out, d.epi, d.alpha = 0, 0, 1.0
    _tc(lib.nk_gemm_ex, "attention PV gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), _stream())


def attention_bwd(do: Tensor, q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, scale: float,
                  out: Optional[tuple] = None):
    """Gradients of attention_fwd.  head_dim 64: the fused flash-style kernel; other head dims: batched tensor-core
       GEMMs with softmax-aware epilogues:
       P = exp(scale*QK^T - lse); dV = P^T dO; dP = dO V^T; dS = P*(dP - delta)*scale; dQ = dS K; dK = dS^T Q.
       `out` = (dq, dk, dv) bf16 destination views (fused path only): (B,N,H,64) views with contiguous heads that
       share one row/batch stride, e.g. the three column blocks of a fused [B, N, 3*H*64] gradient."""
    B, Nq, H, D = q.shape
    Nk = k.shape[1]
    dev = q.device
    do = do.contiguous() if do.stride(3) != 1 else do
    delta = torch.empty((B, H, Nq), dtype=F32, device=dev)
    doc = do.contiguous()
    check(lib.nk_attn_delta(doc.data_ptr(), o.data_ptr(), delta.data_ptr(), B, Nq, H, D, _stream()), "attn_delta")
    if (D <= 64 and D % 8 == 0 and q.stride(2) == D and k.stride(2) == D and v.stride(2) == D
            and not FORCE_MATERIALIZED_ATTN_BWD):
        # fused flash-style backward: one kernel, nothing of size Nq x Nk touches HBM
        dq_acc = torch.zeros((B, Nq, H, D), dtype=F32, device=dev)
        if out is not None:
            dq, dk, dv = out
            assert dk.stride() == dv.stride() and dk.stride(2) == D and dq.stride(2) == D
        else:
            dq = None
            dk = torch.empty((B, Nk, H, D), dtype=BF16, device=dev)
            dv = torch.empty((B, Nk, H, D), dtype=BF16, device=dev)
        check(lib.nk_attention_bwd(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                   v.data_ptr(), v.stride(1), v.stride(0), doc.data_ptr(), doc.stride(1), doc.stride(0),
                                   lse.data_ptr(), delta.data_ptr(), dq_acc.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                   dk.stride(1), dk.stride(0), B, H, Nq, Nk, D, float(scale), _stream()),
              "attention_bwd")
        _count(2)
        if dq is None:
            return cast_bf16(dq_acc), dk, dv
        assert dq.stride(0) == Nq * dq.stride(1)  # rows of all batches are equally spaced: one strided cast
        check(lib.nk_cast_f32_bf16_rows(dq_acc.data_ptr(), H * D, dq.data_ptr(), dq.stride(1), B * Nq, H * D,
                                        _stream()), "cast_f32_bf16_rows")
        _count()
        return dq, dk, dv
    Nkp = (Nk + 7) // 8 * 8
    P = torch.zeros((B, H, Nq, Nkp), dtype=BF16, device=dev) if Nkp != Nk else torch.empty(
        (B, H, Nq, Nkp), dtype=BF16, device=dev)
    LOG2E = 1.4426950408889634
    lse2 = (lse * LOG2E).contiguous()  # tiny (B,H,Nq) helper tensor
    qb, qr, qh = _bhnd(q)
    kb, kr, kh = _bhnd(k)
    vb, vr, vh = _bhnd(v)
    ob, orr, oh = _bhnd(doc)
    st = _stream()
    # P = exp2(scale*log2e * Q K^T - lse*log2e)
    d = nk_gemm_desc()
    _operand(d.A, q, 0, Nq, D, qr, H, qh, B, qb)
    _operand(d.B, k, 0, Nk, D, kr, H, kh, B, kb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nq, Nk, D, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = P.data_ptr(), Nkp, Nq * Nkp, H * Nq * Nkp
    d.out, d.epi, d.alpha = 0, 1, float(scale) * LOG2E
    d.rowvec = lse2.data_ptr()
    _tc(lib.nk_gemm_ex, "attention bwd P gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), st)
    # dV[b,:,h,:] = P^T dO : A = P as MN-major (M = kv), B = dO MN-major (N = d), K = Nq
    dq = torch.empty((B, Nq, H, D), dtype=BF16, device=dev)
    dk = torch.empty((B, Nk, H, D), dtype=BF16, device=dev)
    dv = torch.empty((B, Nk, H, D), dtype=BF16, device=dev)
    d = nk_gemm_desc()
    _operand(d.A, P, 1, Nq, Nk, Nkp, H, Nq * Nkp, B, H * Nq * Nkp)
    _operand(d.B, doc, 1, Nq, D, orr, H, oh, B, ob)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nk, D, Nq, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = dv.data_ptr(), dv.stride(1), dv.stride(2), dv.stride(0)
    d.out, d.epi, d.alpha = 0, 0, 1.0
    _tc(lib.nk_gemm_ex, "attention bwd dV gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), st)
    # dS = P * (dO V^T - delta) * scale
    dS = torch.zeros_like(P) if Nkp != Nk else torch.empty_like(P)
    d = nk_gemm_desc()
    _operand(d.A, doc, 0, Nq, D, orr, H, oh, B, ob)
    _operand(d.B, v, 0, Nk, D, vr, H, vh, B, vb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nq, Nk, D, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = dS.data_ptr(), Nkp, Nq * Nkp, H * Nq * Nkp
    d.out, d.epi, d.alpha = 0, 2, float(scale)
    d.rowvec = delta.data_ptr()
    d.aux = P.data_ptr()
    _tc(lib.nk_gemm_ex, "attention bwd dS gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), st)
    # dQ = dS K : A = dS K-major (K = kv), B = K matrix MN-major (N = d)
    d = nk_gemm_desc()
    _operand(d.A, dS, 0, Nq, Nk, Nkp, H, Nq * Nkp, B, H * Nq * Nkp)
    _operand(d.B, k, 1, Nk, D, kr, H, kh, B, kb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nq, D, Nk, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = dq.data_ptr(), dq.stride(1), dq.stride(2), dq.stride(0)
    d.out, d.epi, d.alpha = 0, 0, 1.0
    _tc(lib.nk_gemm_ex, "attention bwd dQ gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), st)
    # dK = dS^T Q : A = dS MN-major (M = kv), B = Q MN-major (N = d), K = Nq
    d = nk_gemm_desc()
    _operand(d.A, dS, 1, Nq, Nk, Nkp, H, Nq * Nkp, B, H * Nq * Nkp)
    _operand(d.B, q, 1, Nq, D, qr, H, qh, B, qb)
    d.M, d.N, d.K, d.nb2, d.nb1, d.ksize = Nk, D, Nq, H, B, 1
    d.C, d.ldc, d.c_b2_stride, d.c_b1_stride = dk.data_ptr(), dk.stride(1), dk.stride(2), dk.stride(0)
    d.out, d.epi, d.alpha = 0, 0, 1.0
    _tc(lib.nk_gemm_ex, "attention bwd dK gemm", 2.0 * d.M * d.N * d.K * d.nb1 * d.nb2, ctypes.byref(d), st)
    _count(6)
    return dq, dk, dv


# ---- diffusion objective ----------------------------------------------------------------------
def noise_mix(x: Tensor, noise: Tensor, sigma: Tensor, rectified_flow: bool = False) -> Tensor:
    _req_cuda(x, noise, sigma)
    x = x.float().contiguous()
    noise = noise.float().contiguous()
    sigma = sigma.float().contiguous()
    z = torch.empty_like(x)
    B = x.shape[0]
    check(lib.nk_noise_mix(x.data_ptr(), noise.data_ptr(), sigma.data_ptr(), z.data_ptr(), B, x.numel() // B,
                           int(rectified_flow), _stream()), "noise_mix")
    _count()
    return z


# --------------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------------
def _linear_param_grads(dy: Tensor, x: Tensor, weight: Tensor, bias: Optional[Tensor], need_w: bool, need_b: bool):
    """(dW, db) of y = x W^T + b.  With a gradient sink both go straight into the bucket storage — on the side stream
    when WGRAD_OVERLAP is on: the weight-gradient GEMM AND the bias column sum (a bandwidth-bound pass over dy that
    nothing downstream waits for; on the main stream it was 374 launches / 9.7 ms of the SDXL step)."""
    dw = db = None
    wbuf, wbase = _grad_sink(weight) if need_w else (None, None)
    bbuf, bbase = _grad_sink(bias) if (need_b and bias is not None) else (None, None)
    jobs, bases = [], []
    if need_w:
        if wbuf is not None:
            jobs.append(lambda: linear_wgrad(dy, x, out=wbuf))
            bases.append(wbase)
        else:
            dw = linear_wgrad(dy, x)
    if need_b and bias is not None:
        if bbuf is not None:
            jobs.append(lambda: colsum(dy, out=bbuf.view(1, -1)))
            bases.append(bbase)
        else:
            db = colsum(dy)[0]
    if jobs:
        if WGRAD_OVERLAP:
            _fork_wgrad(lambda: [j() for j in jobs], (dy, x), tuple(bases))
        else:
            for j in jobs:
                j()
            for b in bases:
                GRAD_SINK.mark_ready(b)
    return dw, db


class FeedForwardGegluFn(torch.autograd.Function):
    """y = (GEGLU(x W1^T + b1)) W2^T + b2 (+ residual) — `FeedForward` with `glu=True` (reference
    modules/attention.py:50-74) as two GEMMs with the gate and its derivative fused into their epilogues:
      forward   h, a = nk_linear_geglu_fwd(x, W1, b1)      (gate in the epilogue; h kept for the backward)
                y    = nk_linear_fwd(a, W2, b2, residual)
      backward  dh   = nk_linear_dgrad_geglu(dy, W2, h)    (d_a is never written)
                dx   = nk_linear_dgrad(dh, W1)
                dW2 = dy^T a, db2, dW1 = dh^T x, db1        (side stream, into the gradient buckets)"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        h, a = linear_geglu_fwd(x, bf16_weight(w1), f32_param(b1))
        y = linear_fwd(a, bf16_weight(w2), f32_param(b2), residual)
        ctx.save_for_backward(x, h, a, w1, w2)
        ctx.biases = (b1, b2)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h, a, w1, w2 = ctx.saved_tensors
        b1, b2 = ctx.biases
        if dy.dtype != BF16:
            dy = cast_bf16(dy)
        dy = dy.contiguous()
        need = ctx.needs_input_grad
        dh = linear_dgrad_geglu(dy, bf16_weight(w2), h)
        dx = linear_dgrad(dh, bf16_weight(w1)) if need[0] else None
        dw2, db2 = _linear_param_grads(dy, a, w2, b2, need[3], b2 is not None and need[4])
        dw1, db1 = _linear_param_grads(dh, x, w1, b1, need[1], b1 is not None and need[2])
        dres = dy if (ctx.has_res and need[5]) else None
        return dx, dw1, db1, dw2, db2, dres


def feed_forward_geglu(x: Tensor, w1: Tensor, b1: Optional[Tensor], w2: Tensor, b2: Optional[Tensor],
                       residual: Optional[Tensor] = None) -> Tensor:
    return FeedForwardGegluFn.apply(x, w1, b1, w2, b2, residual)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ residual); W, b are the fp32 parameters (reference layout [out, in])."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, out_f32):
        w = bf16_weight(weight)
        y = linear_fwd(x, w, f32_param(bias), residual, out_f32)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias = bias
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if dy.dtype != BF16:
            dy = cast_bf16(dy)
        dy = dy.contiguous()
        w = bf16_weight(weight)
        dx = linear_dgrad(dy, w) if ctx.needs_input_grad[0] else None
        dw = db = None
        dw, db = _linear_param_grads(dy, x, weight, ctx.bias if ctx.has_bias else None, ctx.needs_input_grad[1],
                                     ctx.has_bias and ctx.needs_input_grad[2])
        dres = dy.view(-1, dy.shape[-1]).view(dy.shape) if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
           out_f32: bool = False) -> Tensor:
    return LinearFn.apply(x, weight, bias, residual, out_f32)


class Conv2dFn(torch.autograd.Function):
    """stride-1 3x3 (pad 1) / 1x1 convolution on NHWC bf16 (implicit GEMM).  Input channels must be
    physically padded to >= 64 (zero) — `cin_phys = x.shape[-1]`."""

    @staticmethod
    def forward(ctx, x, weight, bias, bias_img, residual):
        co, ci, ks, _ = weight.shape
        wf, _ = packed_conv_weight(weight)
        y = conv2d_fwd(x, wf, co, ks, f32_param(bias), bias_img, residual)
        ctx.save_for_backward(x, weight)
        ctx.flags = (bias is not None, bias_img is not None, residual is not None)
        ctx.bias_param = bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        co, ci, ks, _ = weight.shape
        has_bias, has_bimg, has_res = ctx.flags
        dy = dy.contiguous()
        _, wd = packed_conv_weight(weight)
        dx = dw = db = dbi = dres = None
        if ctx.needs_input_grad[0]:
            dx = conv2d_fwd(dy, wd, x.shape[-1], ks)
            if dx.shape[-1] != x.shape[-1]:
                dx = dx[..., : x.shape[-1]].contiguous()
        # plain bias (no per-image bias): its column sum joins the weight-gradient job on the side stream
        bias_p = ctx.bias_param if (has_bias and not has_bimg and ctx.needs_input_grad[2]) else None
        bbuf, bbase = _grad_sink(bias_p) if (bias_p is not None and dy.shape[-1] == co) else (None, None)
        if ctx.needs_input_grad[1]:
            buf, base = _grad_sink(weight)
            if buf is not None and WGRAD_OVERLAP:
                def job():
                    conv_unpack_wgrad(conv2d_wgrad(dy, x, co, ks), co, ci, ks, out=buf)
                    if bbuf is not None:
                        colsum(dy, out=bbuf.view(1, -1))
                _fork_wgrad(job, (dy, x), (base,) if bbuf is None else (base, bbase))
            else:
                dwp = conv2d_wgrad(dy, x, co, ks)
                if buf is not None:
                    conv_unpack_wgrad(dwp, co, ci, ks, out=buf)
                    GRAD_SINK.mark_ready(base)
                else:
                    dw = conv_unpack_wgrad(dwp, co, ci, ks)
                if bbuf is not None:
                    colsum(dy, out=bbuf.view(1, -1))
                    GRAD_SINK.mark_ready(bbase)
        elif bbuf is not None:
            colsum(dy, out=bbuf.view(1, -1))
            GRAD_SINK.mark_ready(bbase)
        if (has_bias and ctx.needs_input_grad[2]) or (has_bimg and ctx.needs_input_grad[3]):
            if has_bimg and ctx.needs_input_grad[3]:
                s = colsum(dy, groups=dy.shape[0])[:, :co]  # per-image sums (timestep-embedding gradient)
                dbi = s.contiguous()
                if has_bias and ctx.needs_input_grad[2]:
                    db = s.sum(0)  # O(batch x C) glue
            elif has_bias and ctx.needs_input_grad[2] and bbuf is None:
                db = colsum(dy)[0, :co].contiguous()
        if has_res and ctx.needs_input_grad[4]:
            dres = dy
        return dx, dw, db, dbi, dres


def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, bias_img: Optional[Tensor] = None,
           residual: Optional[Tensor] = None) -> Tensor:
    return Conv2dFn.apply(x, weight, bias, bias_img, residual)


class ConvStridedFn(torch.autograd.Function):
    """3x3 stride-2 convolution (UNet Downsample pad 1; VAE Downsample pad (0,1,0,1)) = im2col + GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad_t, pad_l, ho, wo):
        co, ci, ks, _ = weight.shape
        wf, _ = packed_conv_weight(weight)
        if ks == 3 and x.shape[-1] % 64 == 0 and wf.shape[1] == ks * ks * x.shape[-1]:
            # implicit GEMM: the TMA unit walks the input with element stride 2, nothing is materialised
            y = conv2d_stride2_fwd(x, wf, co, ks, f32_param(bias), pad_t, pad_l, ho, wo)
        else:
            col = im2col(x, ks, 2, pad_t, pad_l, ho, wo)
            y = linear_fwd(col, wf[:co], f32_param(bias))
        ctx.save_for_backward(x, weight)
        ctx.geom = (pad_t, pad_l, ho, wo)
        ctx.has_bias = bias is not None
        return y.view(x.shape[0], ho, wo, co)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        co, ci, ks, _ = weight.shape
        pad_t, pad_l, ho, wo = ctx.geom
        dy2 = dy.contiguous().view(-1, co)
        wf, _ = packed_conv_weight(weight)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dcol = linear_dgrad(dy2, wf[:co])
            dx = col2im(dcol, x.shape, ks, 2, pad_t, pad_l, ho, wo)
        if ctx.needs_input_grad[1]:
            buf, base = _grad_sink(weight)
            if buf is not None and WGRAD_OVERLAP:
                _fork_wgrad(lambda: conv_unpack_wgrad(linear_wgrad(dy2, im2col(x, ks, 2, pad_t, pad_l, ho, wo)), co, ci,
                                                      ks, out=buf), (dy2, x), (base,))
            else:
                col = im2col(x, ks, 2, pad_t, pad_l, ho, wo)
                dwp = linear_wgrad(dy2, col)
                if buf is not None:
                    conv_unpack_wgrad(dwp, co, ci, ks, out=buf)
                    GRAD_SINK.mark_ready(base)
                else:
                    dw = conv_unpack_wgrad(dwp, co, ci, ks)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)[0]
        return dx, dw, db, None, None, None, None


def conv2d_stride2(x: Tensor, weight: Tensor, bias: Optional[Tensor], asymmetric: bool = False) -> Tensor:
    n, h, w_, c = x.shape
    if asymmetric:  # ConstantPad2d((0,1,0,1)) + conv k3 s2 p0
        ho, wo = (h + 1 - 3) // 2 + 1, (w_ + 1 - 3) // 2 + 1
        return ConvStridedFn.apply(x, weight, bias, 0, 0, ho, wo)
    ho, wo = (h + 2 - 3) // 2 + 1, (w_ + 2 - 3) // 2 + 1
    return ConvStridedFn.apply(x, weight, bias, 1, 1, ho, wo)


class GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, silu):
        g, b = f32_param(gamma), f32_param(beta)
        y, mean, rstd = groupnorm_fwd(x, g, b, groups, eps, silu)
        ctx.save_for_backward(x, g, b, mean, rstd)
        ctx.cfg = (groups, silu)
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, b, mean, rstd = ctx.saved_tensors
        groups, silu = ctx.cfg
        bufs, bases = _grad_sink_pair(*ctx.params) if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else (None, None)
        dx, dg, db = groupnorm_bwd(dy.contiguous(), x, g, b, mean, rstd, groups, silu, out=bufs)
        if bufs is not None:
            GRAD_SINK.mark_ready(bases[0])
            GRAD_SINK.mark_ready(bases[1])
            dg = db = None
        return dx, dg, db, None, None, None


def group_norm(x: Tensor, gamma: Tensor, beta: Tensor, groups: int = 32, eps: float = 1e-5, silu: bool = False):
    return GroupNormFn.apply(x, gamma, beta, groups, eps, silu)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        g, b = f32_param(gamma), f32_param(beta)
        y, mean, rstd = layernorm_fwd(x, g, b, eps)
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd = ctx.saved_tensors
        bufs, bases = _grad_sink_pair(*ctx.params) if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else (None, None)
        dx, dg, db = layernorm_bwd(dy, x, g, mean, rstd, out=bufs)
        if bufs is not None:
            GRAD_SINK.mark_ready(bases[0])
            GRAD_SINK.mark_ready(bases[1])
            dg = db = None
        return dx, dg, db, None


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    return LayerNormFn.apply(x, gamma, beta, eps)


class LayerNormResidualFn(torch.autograd.Function):
    """(x, LN(x)) for the pre-norm residual pattern  x -> x + f(LN(x)):  the first output is x itself and carries the
    residual branch.  Backward receives the gradients of BOTH uses of x at once and adds the residual one inside the
    LayerNorm-backward kernel, so the autograd engine never launches a separate gradient add for the fork."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        g, b = f32_param(gamma), f32_param(beta)
        y, mean, rstd = layernorm_fwd(x, g, b, eps)
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.params = (gamma, beta)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dres, dy):
        x, g, mean, rstd = ctx.saved_tensors
        if dy is None:  # LN branch unused: only the residual gradient flows
            return dres, None, None, None
        bufs, bases = _grad_sink_pair(*ctx.params) if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else (None, None)
        dx, dg, db = layernorm_bwd(dy, x, g, mean, rstd, out=bufs, dres=dres)
        if bufs is not None:
            GRAD_SINK.mark_ready(bases[0])
            GRAD_SINK.mark_ready(bases[1])
            dg = db = None
        return dx, dg, db, None


def layer_norm_residual(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5):
    """returns (x, LN(x)); use the returned x for the residual add that follows (see LayerNormResidualFn)."""
    return LayerNormResidualFn.apply(x, gamma, beta, eps)


class GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        return geglu_fwd(h)

    @staticmethod
    def backward(ctx, dout):
        (h,) = ctx.saved_tensors
        return geglu_bwd(h, dout)


def geglu(h: Tensor) -> Tensor:
    return GegluFn.apply(h)


class SiluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return silu_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return silu_bwd(x, dy if dy.dtype == BF16 else cast_bf16(dy))


def silu(x: Tensor) -> Tensor:
    return SiluFn.apply(x)


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return add(a, b)

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add_bf16(a: Tensor, b: Tensor) -> Tensor:
    return AddFn.apply(a, b)


class AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale):
        o, lse = attention_fwd(q, k, v, scale)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = attention_bwd(do, q, k, v, o, lse, ctx.scale)
        return dq, dk, dv, None


class SelfAttentionQKVFn(torch.autograd.Function):
    """o = attention(x Wq^T, x Wk^T, x Wv^T) with the three projections as ONE GEMM (N = 3*inner) in each direction:
    forward x[M,C] @ [Wv;Wk;Wq]^T, data gradient d[v|k|q][M,3*inner] @ [Wv;Wk;Wq] (which also sums the three
    contributions to dx), weight gradient d[v|k|q]^T @ x.  The attention kernels read q/k/v and write dq/dk/dv as
    column slices of the fused buffers.  Stack order v,k,q = the order of the parameters' gradient storage in
    ddp.BucketedGradReducer (reverse registration), so the fused weight gradient lands in the buckets with one GEMM."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, heads, scale):
        B, N, C = x.shape
        inner = wq.shape[0]
        D = inner // heads
        wf = bf16_weight_group((wv, wk, wq))
        vkq = linear_fwd(x, wf).view(B, N, 3, heads, D)
        v, k, q = vkq[:, :, 0], vkq[:, :, 1], vkq[:, :, 2]
        o, lse = attention_fwd(q, k, v, scale)
        ctx.save_for_backward(x, vkq, o, lse, wq, wk, wv)
        ctx.scale = scale
        return o.view(B, N, inner)

    @staticmethod
    def backward(ctx, do):
        x, vkq, o, lse, wq, wk, wv = ctx.saved_tensors
        B, N, _, H, D = vkq.shape
        inner = H * D
        v, k, q = vkq[:, :, 0], vkq[:, :, 1], vkq[:, :, 2]
        if do.dtype != BF16:
            do = cast_bf16(do)
        dvkq = torch.empty_like(vkq)
        outs = (dvkq[:, :, 2], dvkq[:, :, 1], dvkq[:, :, 0])
        res = attention_bwd(do.reshape(B, N, H, D), q, k, v, o, lse, ctx.scale, out=outs)
        for r_, o_ in zip(res, outs):  # the materialised fallback returns fresh tensors
            if r_.data_ptr() != o_.data_ptr():
                o_.copy_(r_)
        d2 = dvkq.view(B * N, 3 * inner)
        wf = bf16_weight_group((wv, wk, wq))
        dx = linear_dgrad(d2, wf).view(x.shape) if ctx.needs_input_grad[0] else None
        grads = [None, None, None]  # for wq, wk, wv
        need = ctx.needs_input_grad[1:4]
        sinks = [_grad_sink(w) for w in (wv, wk, wq)]
        x2 = x.reshape(B * N, -1)
        # adjacent slices of ONE bucket (adjacent addresses alone are not enough: two buckets' flat buffers can be
        # neighbours in the allocator — found by the 2-rank NCCL test with 2 MB buckets)
        stacked = all(need) and all(b is not None for b, _ in sinks) and all(
            sinks[i][0].data_ptr() + sinks[i][0].numel() * 4 == sinks[i + 1][0].data_ptr()
            and sinks[i][0].untyped_storage().data_ptr() == sinks[i + 1][0].untyped_storage().data_ptr() for i in range(2))
        if stacked:  # [dWv; dWk; dWq] is one contiguous [3*inner, C] block of a gradient bucket
            bases = tuple(b for _, b in sinks)
            buf = torch.as_strided(sinks[0][0], (3 * inner, x2.shape[1]), (x2.shape[1], 1))
            if WGRAD_OVERLAP:
                _fork_wgrad(lambda: linear_wgrad(d2, x2, out=buf), (d2, x2), bases)
            else:
                linear_wgrad(d2, x2, out=buf)
                for b in bases:
                    GRAD_SINK.mark_ready(b)
        else:
            for slot, (w, col) in enumerate(((wq, 2), (wk, 1), (wv, 0))):
                if not need[slot]:
                    continue
                dy = d2[:, col * inner: (col + 1) * inner]
                buf, base = _grad_sink(w)
                if buf is not None:
                    linear_wgrad(dy, x2, out=buf)
                    GRAD_SINK.mark_ready(base)
                else:
                    grads[slot] = linear_wgrad(dy, x2)
        return dx, grads[0], grads[1], grads[2], None, None


# Cross-attention with the k and v projections of the context as ONE GEMM (module switch; off until neurosis_b200.tune's step
# guard has compared the training step with and without it on the device — written after the last GPU run)
FUSE_CROSS_KV = _os.environ.get("NK_FUSED_CROSS_KV", "0") not in ("", "0")


class CrossAttentionKVFn(torch.autograd.Function):
    """o = attention(q, ctx Wk^T, ctx Wv^T) with the two context projections as ONE GEMM (N = 2*inner) in each direction:
    forward ctx[M,C] @ [Wv;Wk]^T, weight gradient d[v|k]^T @ ctx (one launch into adjacent bucket slices, like the fused
    q|k|v of self-attention), data gradient d[v|k] @ [Wv;Wk] only if the context needs one (it is an input of the step).
    The 77-token context gives M = 77*B rows: 25 output tiles per separate projection on 74 CTA pairs — half as many,
    twice as wide launches (reference modules/attention.py:283-290, 346-352: to_k / to_v of CrossAttention)."""

    @staticmethod
    def forward(ctx, q, context, wk, wv, heads, scale):
        B, Nk, _ = context.shape
        inner = wk.shape[0]
        D = inner // heads
        wf = bf16_weight_group((wv, wk))
        vk = linear_fwd(context, wf).view(B, Nk, 2, heads, D)
        v, k = vk[:, :, 0], vk[:, :, 1]
        o, lse = attention_fwd(q, k, v, scale)
        ctx.save_for_backward(q, context, vk, o, lse, wk, wv)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, context, vk, o, lse, wk, wv = ctx.saved_tensors
        B, Nk, _, H, D = vk.shape
        inner = H * D
        v, k = vk[:, :, 0], vk[:, :, 1]
        if do.dtype != BF16:
            do = cast_bf16(do)
        dvk = torch.empty_like(vk)
        dq = torch.empty(q.shape, dtype=BF16, device=q.device)
        outs = (dq, dvk[:, :, 1], dvk[:, :, 0])
        res = attention_bwd(do.reshape(q.shape), q, k, v, o, lse, ctx.scale, out=outs)
        for r_, o_ in zip(res, outs):  # the materialised fallback returns fresh tensors
            if r_.data_ptr() != o_.data_ptr():
                o_.copy_(r_)
        d2 = dvk.view(B * Nk, 2 * inner)
        wf = bf16_weight_group((wv, wk))
        dctx = linear_dgrad(d2, wf).view(context.shape) if ctx.needs_input_grad[1] else None
        grads = [None, None]  # for wk, wv
        need = ctx.needs_input_grad[2:4]
        sinks = [_grad_sink(w) for w in (wv, wk)]
        c2 = context.reshape(B * Nk, -1)
        stacked = all(need) and all(b is not None for b, _ in sinks) and (
            sinks[0][0].data_ptr() + sinks[0][0].numel() * 4 == sinks[1][0].data_ptr()
            and sinks[0][0].untyped_storage().data_ptr() == sinks[1][0].untyped_storage().data_ptr())
        if stacked:  # [dWv; dWk] is one contiguous [2*inner, C] block of a gradient bucket
            bases = tuple(b for _, b in sinks)
            buf = torch.as_strided(sinks[0][0], (2 * inner, c2.shape[1]), (c2.shape[1], 1))
            if WGRAD_OVERLAP:
                _fork_wgrad(lambda: linear_wgrad(d2, c2, out=buf), (d2, c2), bases)
            else:
                linear_wgrad(d2, c2, out=buf)
                for b in bases:
                    GRAD_SINK.mark_ready(b)
        else:
            for slot, (w, col) in enumerate(((wk, 1), (wv, 0))):
                if not need[slot]:
                    continue
                dy = d2[:, col * inner: (col + 1) * inner]
                buf, base = _grad_sink(w)
                if buf is not None:
                    linear_wgrad(dy, c2, out=buf)
                    GRAD_SINK.mark_ready(base)
                else:
                    grads[slot] = linear_wgrad(dy, c2)
        return dq, dctx, grads[0], grads[1], None, None


def cross_attention_kv(q: Tensor, context: Tensor, wk: Tensor, wv: Tensor, heads: int, scale: float) -> Tensor:
    """q (B, Nq, H, D) bf16, context (B, Nk, C) bf16 -> o (B, Nq, H, D): fused k|v projection of the context + attention."""
    return CrossAttentionKVFn.apply(q, context, wk, wv, heads, float(scale))


def self_attention_qkv(x: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, heads: int, scale: float) -> Tensor:
    """fused q/k/v projection + attention for self-attention with head_dim <= 64 (multiple of 8); returns (B, N, inner)."""
    return SelfAttentionQKVFn.apply(x, wq, wk, wv, heads, float(scale))


def attention(q: Tensor, k: Tensor, v: Tensor, scale: Optional[float] = None) -> Tensor:
    """softmax(scale * q k^T) v for (B, N, H, D) bf16 tensors."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return AttentionFn.apply(q, k, v, float(scale))


class CatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca, ctx.cb = a.shape[-1], b.shape[-1]
        return cat_channels(a, b)

    @staticmethod
    def backward(ctx, dy):
        ca, cb = ctx.ca, ctx.cb
        da = torch.empty((*dy.shape[:-1], ca), dtype=BF16, device=dy.device)
        db = torch.empty((*dy.shape[:-1], cb), dtype=BF16, device=dy.device)
        copy_channels(dy, da, ca)
        copy_channels(dy[..., ca:], db, cb)
        return da, db


def cat(a: Tensor, b: Tensor) -> Tensor:
    return CatFn.apply(a, b)


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return upsample2x_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        return upsample2x_bwd(dy.contiguous())


def upsample2x(x: Tensor) -> Tensor:
    return Upsample2xFn.apply(x)


class ToNhwcFn(torch.autograd.Function):
    """(N,C,H,W) float -> (N,H,W,Cpad) bf16 with optional per-image scale (c_in of the denoiser)."""

    @staticmethod
    def forward(ctx, x, cpad, scale):
        ctx.c = x.shape[1]
        ctx.save_for_backward(scale if scale is not None else torch.empty(0, device=x.device))
        ctx.in_dtype = x.dtype
        return nchw_to_nhwc(x, cpad, scale)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        dx = nhwc_to_nchw(dy.contiguous(), ctx.c, out_f32=True)
        if scale.numel():
            dx = dx * scale.view(-1, 1, 1, 1)
        return dx.to(ctx.in_dtype), None, None


def to_nhwc(x: Tensor, cpad: Optional[int] = None, scale: Optional[Tensor] = None) -> Tensor:
    return ToNhwcFn.apply(x, cpad, scale)


def lincomb_per_sample(x: Tensor, a: Tensor, y: Optional[Tensor] = None, c: Optional[Tensor] = None) -> Tensor:
    """out[b] = a[b]*x[b] + c[b]*y[b] on fp32 tensors (no autograd)."""
    _req_cuda(x)
    x = x.float().contiguous()
    a = a.float().contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    if y is not None:
        y = y.float().contiguous()
        c = c.float().contiguous()
    check(lib.nk_lincomb_per_sample(x.data_ptr(), a.data_ptr(), _p(y), _p(c), out.data_ptr(), B, x.numel() // B,
                                    _stream()), "lincomb_per_sample")
    _count()
    return out


class DenoiseCombineFn(torch.autograd.Function):
    """D = net * c_out[b] + z * c_skip[b] on (N,C,H,W) fp32 tensors (denoiser.py:53); grad flows to net."""

    @staticmethod
    def forward(ctx, net, z, c_out, c_skip):
        c_out = c_out.float().reshape(-1).contiguous()
        ctx.save_for_backward(c_out)
        return lincomb_per_sample(net, c_out, z, None if z is None else c_skip.float().reshape(-1))

    @staticmethod
    def backward(ctx, dD):
        (c_out,) = ctx.saved_tensors
        return lincomb_per_sample(dD, c_out), None, None, None


def denoise_combine(net: Tensor, z: Optional[Tensor], c_out: Tensor, c_skip: Tensor) -> Tensor:
    return DenoiseCombineFn.apply(net, z, c_out, c_skip)


class FromNhwcFn(torch.autograd.Function):
    """(N,H,W,Cp) bf16 -> (N,C,H,W) fp32 taking the first C channels; backward re-pads with zeros."""

    @staticmethod
    def forward(ctx, y, c):
        ctx.cp = y.shape[-1]
        return nhwc_to_nchw(y, c, out_f32=True)

    @staticmethod
    def backward(ctx, d):
        return nchw_to_nhwc(d.float().contiguous(), ctx.cp), None


def from_nhwc_f32(y: Tensor, channels: int) -> Tensor:
    return FromNhwcFn.apply(y, channels)


class WeightedMseFn(torch.autograd.Function):
    """loss[b] = w[b] * mean((D[b]-T[b])^2) in fp32 (loss.py:153-155, losses/functions.py:81-94)."""

    @staticmethod
    def forward(ctx, D, T, w):
        D = D.float().contiguous()
        T = T.float().contiguous()
        w = w.float().contiguous()
        B = D.shape[0]
        loss = torch.empty((B,), dtype=F32, device=D.device)
        check(lib.nk_weighted_mse_fwd(D.data_ptr(), T.data_ptr(), w.data_ptr(), loss.data_ptr(), B, D.numel() // B,
                                      _stream()), "weighted_mse_fwd")
        _count()
        ctx.save_for_backward(D, T, w)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        D, T, w = ctx.saved_tensors
        B = D.shape[0]
        dD = torch.empty_like(D)
        dl = dloss.float().contiguous()
        check(lib.nk_weighted_mse_bwd(D.data_ptr(), T.data_ptr(), w.data_ptr(), dl.data_ptr(), dD.data_ptr(), B,
                                      D.numel() // B, _stream()), "weighted_mse_bwd")
        _count()
        return dD, None, None


def weighted_mse(D: Tensor, T: Tensor, w: Tensor) -> Tensor:
    return WeightedMseFn.apply(D, T, w)


class WeightedL1Fn(torch.autograd.Function):
    """loss[b] = w[b] * mean(|D[b]-T[b]|) in fp32 (loss.py:153-155 with loss_type "l1", losses/functions.py:65-78)."""

    @staticmethod
    def forward(ctx, D, T, w):
        D = D.float().contiguous()
        T = T.float().contiguous()
        w = w.float().contiguous()
        B = D.shape[0]
        loss = torch.empty((B,), dtype=F32, device=D.device)
        check(lib.nk_weighted_l1_fwd(D.data_ptr(), T.data_ptr(), w.data_ptr(), loss.data_ptr(), B, D.numel() // B,
                                     _stream()), "weighted_l1_fwd")
        _count()
        ctx.save_for_backward(D, T, w)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        D, T, w = ctx.saved_tensors
        B = D.shape[0]
        dD = torch.empty_like(D)
        dl = dloss.float().contiguous()
        check(lib.nk_weighted_l1_bwd(D.data_ptr(), T.data_ptr(), w.data_ptr(), dl.data_ptr(), dD.data_ptr(), B,
                                     D.numel() // B, _stream()), "weighted_l1_bwd")
        _count()
        return dD, None, None


def weighted_l1(D: Tensor, T: Tensor, w: Tensor) -> Tensor:
    return WeightedL1Fn.apply(D, T, w)


# --------------------------------------------------------------------------------------------
# VAE training step: posterior, thin 1x1 convolutions, trainable RGB stem
# --------------------------------------------------------------------------------------------
class DiagGaussianFn(torch.autograd.Function):
    """(z, kl[B]) of the diagonal Gaussian posterior (reference modules/distributions.py:29-51): moments (B,2C,H,W) fp32
    NCHW; eps (B,C,H,W) fp32 standard normal draws, or None for the mode."""

    @staticmethod
    def forward(ctx, moments, eps):
        _req_cuda(moments, eps)
        moments = moments.float().contiguous()
        B, C2 = moments.shape[0], moments.shape[1]
        half = moments.numel() // B // 2
        z = torch.empty((B, C2 // 2, *moments.shape[2:]), dtype=F32, device=moments.device)
        kl = torch.empty((B,), dtype=F32, device=moments.device)
        eps = eps.float().contiguous() if eps is not None else None
        check(lib.nk_diag_gaussian_fwd(moments.data_ptr(), _p(eps), z.data_ptr(), kl.data_ptr(), B, half, _stream()),
              "diag_gaussian_fwd")
        _count()
        ctx.save_for_backward(moments, eps if eps is not None else torch.empty(0, device=moments.device))
        return z, kl

    @staticmethod
    def backward(ctx, dz, dkl):
        moments, eps = ctx.saved_tensors
        eps = eps if eps.numel() else None
        B = moments.shape[0]
        dz = dz.float().contiguous() if dz is not None else None
        dkl = dkl.float().contiguous() if dkl is not None else None
        dm = torch.empty_like(moments)
        check(lib.nk_diag_gaussian_bwd(moments.data_ptr(), _p(eps), _p(dz), _p(dkl), dm.data_ptr(), B,
                                       moments.numel() // B // 2, _stream()), "diag_gaussian_bwd")
        _count()
        return dm, None


def diag_gaussian(moments: Tensor, eps: Optional[Tensor] = None):
    return DiagGaussianFn.apply(moments, eps)


class ThinConv1x1Fn(torch.autograd.Function):
    """1x1 convolution between thin channel counts (quant_conv 8 -> 8, post_quant_conv 4 -> 4; reference
    models/autoencoder.py:452-453) on NHWC bf16 tensors whose channels are physically zero-padded to 64: one
    [pixels, 64] x [64, 64] GEMM with the weight zero-padded (a 64 x 64 host-side pad of a <= 8 x 8 matrix)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        co, ci = weight.shape[0], weight.shape[1]
        cp = x.shape[-1]
        assert cp >= ci and cp % 64 == 0 and co <= 64
        wq = torch.zeros((64, cp), dtype=F32, device=weight.device)
        wq[:co, :ci] = weight.detach().reshape(co, ci)
        wq = cast_bf16(wq)
        bq = None
        if bias is not None:
            bq = torch.zeros((64,), dtype=F32, device=weight.device)
            bq[:co] = bias.detach()
        y = linear_fwd(x.reshape(-1, cp), wq, bq)
        ctx.save_for_backward(x, wq)
        ctx.dims = (co, ci, bias is not None)
        ctx.wshape = weight.shape
        return y.view(*x.shape[:-1], 64)

    @staticmethod
    def backward(ctx, dy):
        x, wq = ctx.saved_tensors
        co, ci, has_bias = ctx.dims
        cp = x.shape[-1]
        dy2 = dy.contiguous().view(-1, 64)
        dx = linear_dgrad(dy2, wq).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = linear_wgrad(dy2, x.reshape(-1, cp))[:co, :ci].reshape(ctx.wshape).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)[0, :co].contiguous()
        return dx, dw, db


def conv1x1_thin(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    return ThinConv1x1Fn.apply(x, weight, bias)


class Conv3x3ThinInputFn(torch.autograd.Function):
    """Trainable form of `conv3x3_thin_input_fwd` (Encoder.conv_in, RGB -> 128): the weight gradient is the
    [Cout, 64] x [pixels, 64] GEMM over the re-generated patch matrix; the image itself takes no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return conv3x3_thin_input_fwd(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        n, c, h, w_ = x.shape
        co = weight.shape[0]
        dy2 = dy.contiguous().view(-1, co)
        dw = db = None
        if ctx.needs_input_grad[1]:
            xf = x.float().contiguous()
            col = torch.empty((n * h * w_, 64), dtype=BF16, device=x.device)
            check(lib.nk_image_patches3x3(xf.data_ptr(), col.data_ptr(), n, c, h, w_, _stream()), "image_patches3x3")
            _count()
            dwp = linear_wgrad(dy2, col)  # [co, 64], column = tap*C + c
            dw = dwp[:, : 9 * c].reshape(co, 3, 3, c).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)[0].contiguous()
        return None, dw, db


def conv3x3_thin_input(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    return Conv3x3ThinInputFn.apply(x, weight, bias)


def mse_loss(pred: Tensor, target: Tensor) -> Tensor:
    """F.mse_loss(pred, target) (mean over all elements) on fp32 tensors of equal per-sample size, through the
    per-sample reduction kernel of the diffusion loss."""
    ones = torch.ones((pred.shape[0],), dtype=F32, device=pred.device)
    return weighted_mse(pred, target, ones).mean()
