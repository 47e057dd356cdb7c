"""The oracle (CPU restatement) against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Runs anywhere — no GPU, no /root/reference."""
import numpy as np
import pytest
import torch

from common import TINY_SD15, TINY_SDXL, TINY_VAE
from oracle import objective as O
from oracle.unet import unet_forward, unet_param_shapes
from oracle.vae import vae_encode, vae_param_shapes
from oracle.weights import synth_state_dict, synth_tensor

G = np.load(__file__.rsplit("/", 1)[0] + "/golden/reference_golden.npz")


def test_sigma_tables_bit_exact():
    assert np.array_equal(O.ddpm_sigma_table(1000, flip=False).numpy(), G["table_desc"])
    assert np.array_equal(O.ddpm_sigma_table(1000, flip=True).numpy(), G["table_asc"])
    assert G["table_desc"].shape == (1001,) and G["table_desc"][-1] == 0.0


def test_sigma_index_bit_exact():
    table = torch.from_numpy(G["table_desc"])
    probe = (synth_tensor("sigma_probe", (64,), uniform=True).abs() * 15.0).float()
    assert np.array_equal(O.sigma_to_idx(table, probe).numpy(), G["sigma_probe_idx_f32"])
    assert np.array_equal(O.sigma_to_idx(table, probe.to(torch.bfloat16)).numpy(), G["sigma_probe_idx_bf16"])
    assert np.array_equal(table[O.sigma_to_idx(table, probe)].numpy(), G["sigma_probe_quant"])


def test_sigma_generator_draws():
    asc = torch.from_numpy(G["table_asc"])
    t = torch.from_numpy(G["gen_t"])
    assert np.array_equal(O.discrete_sigma_draw(asc, 1000, 8, t).numpy(), G["gen_sigma_from_t"])
    assert (G["gen_sigma_from_t"] == 0.0).all()  # the reference's t in [0,1) -> idx 0 quirk
    torch.manual_seed(42)
    assert np.array_equal(O.discrete_sigma_draw(asc, 1000, 8, None).numpy(), G["gen_sigma_randint"])


@pytest.mark.parametrize("tag,cfg", [("sdxl", TINY_SDXL), ("sd15", TINY_SD15)])
def test_unet_forward_backward(tag, cfg):
    shapes = unet_param_shapes(cfg)
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, seed=1).items()}
    x = synth_tensor(f"{tag}.x", (2, 4, 16, 16))
    ctx = synth_tensor(f"{tag}.ctx", (2, 77, cfg["context_dim"]))
    y = synth_tensor(f"{tag}.y", (2, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    out = unet_forward(sd, cfg, x, torch.tensor([17, 803]), ctx, y)
    np.testing.assert_allclose(out.detach().numpy(), G[f"{tag}.out"], rtol=1e-4, atol=2e-5)
    (out * synth_tensor(f"{tag}.gout", tuple(out.shape), scale=0.1)).sum().backward()
    names = sorted(shapes)
    l2 = np.array([sd[n].grad.norm().item() for n in names])
    np.testing.assert_allclose(l2, G[f"{tag}.grad_l2"], rtol=2e-4, atol=1e-6)
    for n in ("out.2.weight", "input_blocks.0.0.bias", "time_embed.0.bias"):
        np.testing.assert_allclose(sd[n].grad.numpy(), G[f"{tag}.grad.{n}"], rtol=1e-3, atol=1e-5)


def test_vae_encode():
    sd = synth_state_dict(vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True), seed=2)
    img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
    with torch.no_grad():
        z = vae_encode(sd, TINY_VAE, img)
    np.testing.assert_allclose(z.numpy(), G["vae.z"], rtol=1e-4, atol=2e-5)


def test_step_loss_and_grads():
    cfg = TINY_SDXL
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
    table = torch.from_numpy(G["table_desc"])
    lat, noise = synth_tensor("step.latent", (2, 4, 16, 16)), synth_tensor("step.noise", (2, 4, 16, 16))
    cond = {"crossattn": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])),
            "vector": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"]))}
    net = lambda x, t, c: unet_forward(sd, cfg, x, t, c["crossattn"], c["vector"])  # noqa: E731
    loss = O.diffusion_loss(net, table, lat, cond, torch.from_numpy(G["step.sigmas"]), noise)
    np.testing.assert_allclose(loss.detach().numpy(), G["step.loss"], rtol=1e-4)
    loss.mean().backward()
    np.testing.assert_allclose(sd["out.2.weight"].grad.numpy(), G["step.grad.out.2.weight"], rtol=1e-3, atol=1e-6)
    l2 = np.array([sd[n].grad.norm().item() for n in sorted(sd)])
    np.testing.assert_allclose(l2, G["step.grad_l2"], rtol=5e-4, atol=1e-7)
