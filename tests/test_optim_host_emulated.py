"""Host side of the fused optimizer / EMA, end to end WITHOUT a GPU: the C-ABI entry points are replaced by numpy
emulations that read the very device tables the host code builds (raw pointers, record layout, block table, hyper
array — all valid in-process for CPU tensors), so the table construction, record splitting, owner indices, per-step
scalars, state handling and the EMA span table are checked against the REFERENCE optimizer's goldens.  What this
cannot cover is the CUDA kernels themselves — tests/test_gpu_next.py does that on the B200.

The emulation follows the kernels' arithmetic (csrc/optim.cu) record by record; it is test infrastructure only."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle.weights import synth_tensor
from test_oracle_golden_next import EMA_SHAPES, G, OPT_CASES, OPT_SHAPES, check_optimizer_state


def _f32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(ptr)))


class FakeLib:
    """numpy stand-ins for nk_adafactor_step / nk_adafactor_hyper / nk_ema_update_multi / nk_ema_decay."""

    def __init__(self, optim_mod):
        self.O = optim_mod
        self.calls = []

    def nk_adafactor_step(self, table_ptr, blk_ptr, n_tensors, n_blocks, hyper_ptr, scal_ptr, rms_ptr, stream):
        O = self.O
        raw = np.ctypeslib.as_array((ctypes.c_uint8 * (88 * n_tensors)).from_address(int(table_ptr)))
        T = raw.view(O._TENSOR_DTYPE)
        blk = np.ctypeslib.as_array((ctypes.c_int32 * n_tensors).from_address(int(blk_ptr)))
        # the block table must be the running sum of the per-record block counts the kernels assume
        run = 0
        for t in range(n_tensors):
            assert blk[t] == run, "blk_start"
            run += O.blocks_for(int(T["kind"][t]), int(T["n"][t]), int(T["Bt"][t]), int(T["R"][t]), int(T["C"][t]))
        assert run == n_blocks
        groups = int(T["group"].max()) + 1
        H = _f32(hyper_ptr, 8 * groups).reshape(groups, 8)
        rms = _f32(rms_ptr, n_tensors)
        recs = []
        p_sq = np.zeros(n_tensors)
        u_sq = np.zeros(n_tensors)
        for t in range(n_tensors):
            r = T[t]
            kind, Bt, R, C, n = int(r["kind"]), int(r["Bt"]), int(r["R"]), int(r["C"]), int(r["n"])
            b2, _, eps1 = (float(x) for x in H[int(r["group"])][:3])
            cnt = n if kind != O.KIND_MAT else R * C
            p, g = _f32(r["p"], cnt), _f32(r["g"], cnt)
            owner = int(r["owner"])
            assert T["n"][owner] == n and (kind == O.KIND_MAT or owner == t)
            p_sq[owner] += float((p.astype(np.float64) ** 2).sum())
            if kind == O.KIND_VEC:
                v = _f32(r["vr"], n)
                v[:] = v * b2 + (g * g + eps1) * (1.0 - b2)
                upd = g / np.sqrt(v)
            else:
                g3 = g.reshape(Bt, R, C)
                u = g3 * g3 + eps1
                vr, vc = _f32(r["vr"], Bt * R).reshape(Bt, R), _f32(r["vc"], Bt * C).reshape(Bt, C)
                if kind == O.KIND_MAT:
                    sc = _f32(r["scratch"], R + C)
                    assert not sc.any(), "scratch must be zero between steps"
                vr[:] = vr * b2 + u.mean(-1) * (1.0 - b2)
                vc[:] = vc * b2 + u.mean(-2) * (1.0 - b2)
                rf = 1.0 / np.sqrt(vr / vr.mean(-1, keepdims=True))
                upd = (rf[:, :, None] * (1.0 / np.sqrt(vc))[:, None, :] * g3).reshape(-1)
            u_sq[owner] += float((upd.astype(np.float64) ** 2).sum())
            recs.append((r, p, upd.astype(np.float32), cnt))
        for t, (r, p, upd, cnt) in enumerate(recs):
            owner, n = int(r["owner"]), int(r["n"])
            _, rel, _, eps2, clip, wd, beta1, scale_param = (float(x) for x in H[int(r["group"])])
            rms_p = math.sqrt(p_sq[owner] / n)
            lr = (max(eps2, rms_p) if scale_param else 1.0) * rel
            d = upd / max(1.0, math.sqrt(u_sq[owner] / n) / clip) * lr
            if int(r["exp_avg"]):
                m = _f32(r["exp_avg"], cnt)
                m[:] = m * beta1 + d * (1.0 - beta1)
                d = m
            if wd:
                p += p * (-wd * lr)
            p -= d.astype(np.float32)
            rms[t] = rms_p
        self.calls.append("adafactor_step")
        return 0

    def nk_adafactor_hyper(self, consts_ptr, step_ptr, hyper_ptr, n_groups, stream):
        c = _f32(consts_ptr, 8 * n_groups).reshape(n_groups, 8)
        step = np.ctypeslib.as_array((ctypes.c_int64 * n_groups).from_address(int(step_ptr)))
        H = _f32(hyper_ptr, 8 * n_groups).reshape(n_groups, 8)
        for g in range(n_groups):
            step[g] += 1
            s, flags = float(step[g]), int(c[g][7])
            rel = min(1e-6 * s if flags & 2 else 1e-2, s ** -0.5) if flags & 1 else c[g][1]
            H[g] = [1.0 - s ** float(c[g][0]), rel, c[g][2], c[g][3], c[g][4], c[g][5], c[g][6], 1.0 if flags & 4 else 0.0]
        return 0

    def nk_ema_update_multi(self, spans_ptr, n_spans, omd_ptr, stream):
        spans = np.ctypeslib.as_array((ctypes.c_int64 * (3 * n_spans)).from_address(int(spans_ptr))).reshape(n_spans, 3)
        w = np.float32(_f32(omd_ptr, 1)[0])
        for sp, pp, n in spans:
            s, p = _f32(sp, n), _f32(pp, n)
            s -= w * (s - p)
        return 0

    def nk_ema_decay(self, decay, nu_ptr, omd_ptr, stream):
        nu = np.ctypeslib.as_array((ctypes.c_int32 * 1).from_address(int(nu_ptr)))
        d = np.float32(decay)
        if nu[0] >= 0:
            nu[0] += 1
            d = min(d, np.float32(1 + nu[0]) / np.float32(10 + nu[0]))
        _f32(omd_ptr, 1)[0] = np.float32(1.0) - d
        return 0


@pytest.fixture
def emulated(monkeypatch):
    from neurosis_b200 import ops, optim
    fake = FakeLib(optim)
    monkeypatch.setattr(optim, "lib", fake)
    monkeypatch.setattr(optim, "check", lambda rc, what="": None if rc == 0 else (_ for _ in ()).throw(RuntimeError(what)))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)  # "device" = host memory
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self, raising=False)
    return fake


@pytest.mark.parametrize("graph_form", [False, True])
@pytest.mark.parametrize("case", ["yaml", "ext"])
def test_adafactor_host_tables_against_reference_goldens(emulated, case, graph_form):
    from neurosis_b200.optim import Adafactor
    kw = OPT_CASES[case]
    params = {k: torch.nn.Parameter(synth_tensor(f"opt.p.{k}", s, scale=0.05)) for k, s in OPT_SHAPES.items()}
    p0 = {k: p.detach().numpy().copy() for k, p in params.items()}
    opt = Adafactor(list(params.values()), **kw)
    grads = {k: torch.zeros(s) for k, s in OPT_SHAPES.items()}  # stable gradient storage, as with the reducer's buckets
    for k, p in params.items():
        p.grad = grads[k]
    for step in range(3):
        for k in params:
            grads[k].copy_(synth_tensor(f"opt.g.{k}.{step}", OPT_SHAPES[k], scale=0.02 * (step + 1)))
        v0 = {k: p._version for k, p in params.items()}
        if graph_form:
            if step == 0:
                opt.graph_prepare()
            opt.graph_launch()
        else:
            opt.step()
            assert all(p._version > v0[k] for k, p in params.items())
    if graph_form:
        opt.sync_steps_from_device()
    assert emulated.calls == ["adafactor_step"] * 3
    for k, p in params.items():
        st = {sk: (v.detach().numpy() if torch.is_tensor(v) else v) for sk, v in opt.state[p].items()}
        check_optimizer_state(case, k, p.detach().numpy(), st, p0[k])
        assert opt.state[p]["step"] == 3
    # resume: state dict round trip rebuilds the table around the loaded tensors
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    assert opt._plan is None


def test_lit_ema_host_path_against_reference_goldens(emulated):
    from neurosis_b200.optim import LitEma

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(40, 30)
            self.b = torch.nn.Conv2d(8, 8, 3)

    for graph_form in (False, True):
        m = M()
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.copy_(synth_tensor(f"ema.p.{n}", tuple(p.shape)))
        ema = LitEma(m, decay=0.9999)
        if graph_form:
            ema.graph_prepare(m)
        for it in range(12):
            with torch.no_grad():
                for n, p in m.named_parameters():
                    p.add_(synth_tensor(f"ema.d.{n}.{it}", tuple(p.shape), scale=0.1))
            ema.graph_launch() if graph_form else ema(m)
        if graph_form:
            ema.sync_from_device()
        sh = dict(ema.named_buffers())
        for n in EMA_SHAPES:
            np.testing.assert_allclose(sh[n.replace(".", "_")].numpy(), G[f"ema.{n}"], rtol=1e-6, atol=1e-7)
        assert int(ema.num_updates) == 12 and ema._n_host == 12


def test_lit_ema_resume_rederives_the_decay_from_the_loaded_counter(emulated):
    """after `load_state_dict` the warm-up decay must come from the LOADED `num_updates` (the
    reference reads the buffer directly, modules/ema.py:44-46) — resumed at 50 000 updates the decay is 0.9999, not the
    2/11 of a fresh counter — and an interrupted-and-resumed run equals an uninterrupted one."""
    from neurosis_b200.optim import LitEma

    def model():
        m = torch.nn.Linear(40, 30)
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.copy_(synth_tensor(f"ema.p.a.{n}", tuple(p.shape)))
        return m

    def drift(m, it):
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.add_(synth_tensor(f"ema.d.a.{n}.{it}", tuple(p.shape), scale=0.1))

    # (1) decay after loading a long-trained counter
    m = model()
    ema = LitEma(m, decay=0.9999)
    sd = ema.state_dict()
    sd["num_updates"] = torch.tensor(50000, dtype=torch.int)
    ema2 = LitEma(model(), decay=0.5)
    ema2.load_state_dict(sd)
    assert ema2._n_host == 50000 and abs(ema2._decay_host - 0.9999) < 1e-7
    ema2._n_host += 1
    assert abs(ema2.current_decay() - min(0.9999, 50002 / 50011)) < 1e-7  # 0.99982, not the 2/11 of a fresh counter
    # (2) 4 updates + save/load + 4 updates == 8 updates
    ma, mb = model(), model()
    ea, eb = LitEma(ma, decay=0.9999), LitEma(mb, decay=0.9999)
    for it in range(8):
        drift(ma, it)
        ea(ma)
    for it in range(4):
        drift(mb, it)
        eb(mb)
    ec = LitEma(mb, decay=0.9999)
    ec.load_state_dict(eb.state_dict())
    for it in range(4, 8):
        drift(mb, it)
        ec(mb)
    assert int(ec.num_updates) == 8 == ec._n_host
    for (na, ba), (nc, bc) in zip(ea.named_buffers(), ec.named_buffers()):
        assert na == nc
        np.testing.assert_allclose(bc.numpy(), ba.numpy(), rtol=1e-6, atol=1e-7)
