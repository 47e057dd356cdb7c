"""Host side of the diffusion objective WITHOUT a GPU: the real `StandardDiffusionLoss`, `Denoiser` / `DiscreteDenoiser`
and `DiagonalGaussianRegularizer` code of neurosis_b200 runs with the fp32 objective kernels (nk_noise_mix,
nk_lincomb_per_sample, nk_weighted_{mse,l1}_{fwd,bwd}, nk_diag_gaussian_{fwd,bwd}) replaced by numpy emulations that
read the raw pointers the wrappers pass, and with the CPU oracle UNet as the network.  Loss values and the gradient of
every UNet parameter are compared with goldens produced by the REFERENCE's StandardDiffusionLoss for the edm / L2,
edm / L1 and rectified-flow configurations.  The CUDA kernels themselves are covered by tests/test_gpu_*.py."""
import ctypes

import numpy as np
import pytest
import torch

from common import ROOT, TINY_SDXL
from oracle.unet import unet_forward, unet_param_shapes
from oracle.weights import synth_state_dict, synth_tensor

G0 = np.load(str(ROOT / "tests/golden/reference_golden.npz"))
G1 = np.load(str(ROOT / "tests/golden/reference_golden_next.npz"))


def _f32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(ptr)))


class FakeObjectiveLib:
    """numpy stand-ins following csrc/diffusion.cu"""

    def nk_noise_mix(self, x, noise, sigma, z, B, per, rf, stream):
        X, N, Z, s = _f32(x, B * per).reshape(B, per), _f32(noise, B * per).reshape(B, per), \
            _f32(z, B * per).reshape(B, per), _f32(sigma, B)[:, None]
        Z[:] = ((1.0 - s) if rf else 1.0) * X + s * N
        return 0

    def nk_lincomb_per_sample(self, x, a, y, c, out, B, per, stream):
        O = _f32(out, B * per).reshape(B, per)
        v = _f32(a, B)[:, None] * _f32(x, B * per).reshape(B, per)
        if y:
            v = v + _f32(c, B)[:, None] * _f32(y, B * per).reshape(B, per)
        O[:] = v
        return 0

    def _pair(self, D, T, B, per):
        return _f32(D, B * per).reshape(B, per), _f32(T, B * per).reshape(B, per)

    def nk_weighted_mse_fwd(self, D, T, w, loss, B, per, stream):
        d, t = self._pair(D, T, B, per)
        _f32(loss, B)[:] = ((d - t) ** 2).mean(1) * _f32(w, B)
        return 0

    def nk_weighted_mse_bwd(self, D, T, w, dloss, dD, B, per, stream):
        d, t = self._pair(D, T, B, per)
        _f32(dD, B * per).reshape(B, per)[:] = (_f32(dloss, B) * _f32(w, B))[:, None] * (2.0 / per) * (d - t)
        return 0

    def nk_weighted_l1_fwd(self, D, T, w, loss, B, per, stream):
        d, t = self._pair(D, T, B, per)
        _f32(loss, B)[:] = np.abs(d - t).mean(1) * _f32(w, B)
        return 0

    def nk_weighted_l1_bwd(self, D, T, w, dloss, dD, B, per, stream):
        d, t = self._pair(D, T, B, per)
        _f32(dD, B * per).reshape(B, per)[:] = (_f32(dloss, B) * _f32(w, B))[:, None] * (1.0 / per) * np.sign(d - t)
        return 0

    def nk_diag_gaussian_fwd(self, moments, eps, z, kl, B, half, stream):
        m = _f32(moments, 2 * B * half).reshape(B, 2, half)
        mean, lv = m[:, 0], np.clip(m[:, 1], -30.0, 20.0)
        Z = _f32(z, B * half).reshape(B, half)
        Z[:] = mean + np.exp(0.5 * lv) * _f32(eps, B * half).reshape(B, half) if eps else mean
        if kl:
            _f32(kl, B)[:] = 0.5 * (mean ** 2 + np.exp(lv) - 1.0 - lv).sum(1)
        return 0

    def nk_diag_gaussian_bwd(self, moments, eps, dz, dkl, dmoments, B, half, stream):
        m = _f32(moments, 2 * B * half).reshape(B, 2, half)
        mean, lvr = m[:, 0], m[:, 1]
        lv = np.clip(lvr, -30.0, 20.0)
        g = _f32(dz, B * half).reshape(B, half) if dz else np.zeros((B, half), np.float32)
        gk = _f32(dkl, B)[:, None] if dkl else np.zeros((B, 1), np.float32)
        dm = _f32(dmoments, 2 * B * half).reshape(B, 2, half)
        dm[:, 0] = g + gk * mean
        inside = (lvr >= -30.0) & (lvr <= 20.0)
        e = _f32(eps, B * half).reshape(B, half) if eps else 0.0
        dm[:, 1] = inside * (g * e * 0.5 * np.exp(0.5 * lv) + gk * 0.5 * (np.exp(lv) - 1.0))
        return 0


@pytest.fixture
def emulated(monkeypatch):
    from neurosis_b200 import ops
    monkeypatch.setattr(ops, "lib", FakeObjectiveLib())
    monkeypatch.setattr(ops, "check", lambda rc, what="": None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)


def _setup():
    cfg = TINY_SDXL
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
    lat, noise = synth_tensor("step.latent", (2, 4, 16, 16)), synth_tensor("step.noise", (2, 4, 16, 16))
    cond = {"crossattn": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])),
            "vector": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"]))}

    def network(x, t, c, **kw):  # the wrapped-UNet call convention of denoiser.py:49
        return unet_forward(sd, cfg, x, t, c["crossattn"], c["vector"])

    return sd, lat, noise, cond, network


def _fixed(sig):
    class Fixed:
        def __call__(self, n, t=None):
            return sig
    return Fixed()


def _grad_l2(sd):
    return np.array([sd[n].grad.norm().item() for n in sorted(sd)])


def test_edm_l2_objective_host_code_vs_reference(emulated):
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    sd, lat, noise, cond, network = _setup()
    den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization())
    loss_fn = StandardDiffusionLoss(_fixed(torch.from_numpy(G0["step.sigmas"])), EpsWeighting())
    loss, info = loss_fn._forward(network, den, cond, lat, {}, return_dict=True, noise=noise)
    assert loss.shape == (2,) and loss.dtype == torch.float32 and "sigmas" in info
    np.testing.assert_allclose(loss.detach().numpy(), G0["step.loss"], rtol=1e-4)
    loss.mean().backward()
    np.testing.assert_allclose(_grad_l2(sd), G0["step.grad_l2"], rtol=5e-4, atol=1e-7)


def test_edm_l1_objective_host_code_vs_reference(emulated):
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    sd, lat, noise, cond, network = _setup()
    den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization())
    loss_fn = StandardDiffusionLoss(_fixed(torch.from_numpy(G1["l1.sigmas"])), EpsWeighting(), loss_type="l1")
    loss = loss_fn._forward(network, den, cond, lat, {}, noise=noise)
    np.testing.assert_allclose(loss.detach().numpy(), G1["l1.loss"], rtol=1e-4)
    loss.mean().backward()
    np.testing.assert_allclose(_grad_l2(sd), G1["l1.grad_l2"], rtol=5e-4, atol=1e-7)


def test_rectified_flow_objective_host_code_vs_reference(emulated):
    from neurosis_b200.modules.denoiser import Denoiser, RectifiedFlowComfyPreconditioning, RectifiedFlowComfyWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    sd, lat, noise, cond, network = _setup()
    loss_fn = StandardDiffusionLoss(_fixed(torch.from_numpy(G1["rf.sigmas"])), RectifiedFlowComfyWeighting(),
                                    objective_type="rf")
    loss = loss_fn._forward(network, Denoiser(RectifiedFlowComfyPreconditioning()), cond, lat, {}, noise=noise)
    np.testing.assert_allclose(loss.detach().numpy(), G1["rf.loss"], rtol=1e-4)
    loss.mean().backward()
    np.testing.assert_allclose(_grad_l2(sd), G1["rf.grad_l2"], rtol=5e-4, atol=1e-7)


def test_sampling_posterior_host_code_vs_oracle(emulated):
    from neurosis_b200.modules.vae import DiagonalGaussianRegularizer
    from oracle.vae import diag_gaussian
    m = synth_tensor("dg.m", (3, 8, 6, 5))
    m[0, 4:, :2] = 25.0
    m[1, 4:, :2] = -40.0
    eps = synth_tensor("dg.eps", (3, 4, 6, 5))
    gz = synth_tensor("dg.gz", (3, 4, 6, 5))
    a = m.clone().requires_grad_(True)
    z, log = DiagonalGaussianRegularizer(sample=True)(a, eps)
    ((z * gz).sum() + 0.3 * log["kl_loss"]).backward()
    b = m.clone().requires_grad_(True)
    zr, klr = diag_gaussian(b, eps)
    ((zr * gz).sum() + 0.3 * klr.sum() / 3).backward()
    np.testing.assert_allclose(z.detach().numpy(), zr.detach().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(float(log["kl_loss"]), float(klr.sum() / 3), rtol=1e-6)
    np.testing.assert_allclose(a.grad.numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-6)
    zm, log0 = DiagonalGaussianRegularizer(sample=False)(m)
    assert torch.equal(zm, m[:, :4]) and log0 == {}


def test_hook_sample_weights_enter_the_reduction_kernel(emulated):
    """the tag-frequency multipliers are folded into the weight vector of the weighted-MSE kernel.
    (1) `_forward(sample_weights=w)` == w * `_forward()` in value and in every gradient; (2) the hook protocol of
    `DiffusionEngine.training_step` — pre_hook leaves the weights in the batch, the loss consumes them, the post-loss hook
    call does not multiply a second time — gives the same numbers as the multiply-afterwards form."""
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import (APPLIED_KEY, SAMPLE_WEIGHT_KEY, StandardDiffusionLoss, TagFreqScale,
                                            TagFrequencyHook)
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization())
    sig = torch.from_numpy(G0["step.sigmas"])
    w = torch.tensor([0.8, 1.3])
    sd, lat, noise, cond, network = _setup()
    loss_fn = StandardDiffusionLoss(_fixed(sig), EpsWeighting())
    fused = loss_fn._forward(network, den, cond, lat, {}, noise=noise, sample_weights=w)
    np.testing.assert_allclose(fused.detach().numpy(), G0["step.loss"] * w.numpy(), rtol=1e-4)
    fused.mean().backward()
    g_fused = {n: sd[n].grad.clone() for n in sd}
    sd2, lat, noise, cond, network2 = _setup()
    plain = loss_fn._forward(network2, den, cond, lat, {}, noise=noise)
    (plain * w).mean().backward()
    for n in sd2:
        np.testing.assert_allclose(g_fused[n].numpy(), sd2[n].grad.numpy(), rtol=2e-4, atol=1e-8)
    # hook protocol: same captions through two identical hooks, fused vs multiply-afterwards
    caps = ["1girl solo smile", "landscape scenery sky cloud"]
    mk = lambda: TagFrequencyHook(alpha=0.5, beta=0.9, freq_scale=TagFreqScale([[-1, 1.4], [0.5, 0.7]]))  # noqa: E731
    h1, h2 = mk(), mk()
    batch = {"caption": list(caps)}
    batch = h1.pre_hook(None, None, batch, 0)
    assert SAMPLE_WEIGHT_KEY in batch
    sd3, lat, noise, cond, network3 = _setup()
    l1 = loss_fn._forward(network3, den, cond, lat, batch, noise=noise)
    assert batch.get(APPLIED_KEY)
    l1b, log1 = h1(None, batch, l1, {})
    assert l1b is l1 and SAMPLE_WEIGHT_KEY not in batch and APPLIED_KEY not in batch
    l2, log2 = h2(None, {"caption": list(caps)}, plain.detach(), {})  # no pre_hook: the hook multiplies afterwards
    np.testing.assert_allclose(l1.detach().numpy(), l2.numpy(), rtol=1e-4)
    assert abs(float(log1["TagFrequencyHook/scale_mean"]) - float(log2["TagFrequencyHook/scale_mean"])) < 1e-6
    assert h1.counts == h2.counts  # the running tag counts advanced exactly once per step in both forms
