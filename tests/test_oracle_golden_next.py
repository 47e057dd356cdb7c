"""The oracle restatements of the SURVEY.md §8(f) rows against golden vectors produced by the reference itself
(tests/golden/make_golden_next.py).  Runs anywhere — no GPU, no /root/reference."""
import numpy as np
import torch

from common import TINY_VAE
from oracle.vae import vae_decode, vae_decoder_param_shapes, vae_param_shapes, vae_train_loss
from oracle.weights import synth_state_dict, synth_tensor

G = np.load(__file__.rsplit("/", 1)[0] + "/golden/reference_golden_next.npz")


def test_vae_decoder_forward_backward():
    shapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, seed=5).items()}
    z = synth_tensor("vaedec.z", (2, 4, 16, 16))
    xr = vae_decode(sd, TINY_VAE, z)
    np.testing.assert_allclose(xr.detach().numpy(), G["vaedec.out"], rtol=1e-4, atol=2e-5)
    (xr * synth_tensor("vaedec.g", tuple(xr.shape), scale=0.1)).sum().backward()
    l2 = np.array([sd[n].grad.norm().item() for n in sorted(shapes)])
    np.testing.assert_allclose(l2, G["vaedec.grad_l2"], rtol=2e-4, atol=1e-6)
    for n in ("conv_out.weight", "post_quant_conv.weight"):
        np.testing.assert_allclose(sd[n].grad.numpy(), G[f"vaedec.grad.{n}"], rtol=1e-3, atol=1e-5)


def test_vae_training_step():
    eshapes = vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    dshapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    esd = {k: v.requires_grad_(True) for k, v in synth_state_dict(eshapes, seed=2).items()}
    dsd = {k: v.requires_grad_(True) for k, v in synth_state_dict(dshapes, seed=5).items()}
    img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
    eps = synth_tensor("vaetrain.eps", (2, 4, 16, 16))
    loss, z, xrec, kl = vae_train_loss(esd, dsd, TINY_VAE, img, eps)
    np.testing.assert_allclose(z.detach().numpy(), G["vaetrain.z"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(xrec.detach().numpy(), G["vaetrain.xrec"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(loss.item(), G["vaetrain.loss"], rtol=1e-5)
    np.testing.assert_allclose(kl.item(), G["vaetrain.kl_loss"], rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(np.array([esd[n].grad.norm().item() for n in sorted(eshapes)]),
                               G["vaetrain.enc_grad_l2"], rtol=5e-4, atol=1e-8)
    np.testing.assert_allclose(np.array([dsd[n].grad.norm().item() for n in sorted(dshapes)]),
                               G["vaetrain.dec_grad_l2"], rtol=5e-4, atol=1e-8)
    np.testing.assert_allclose(esd["quant_conv.weight"].grad.numpy(), G["vaetrain.grad.quant_conv.weight"], rtol=1e-3,
                               atol=1e-7)
    np.testing.assert_allclose(esd["conv_in.weight"].grad.numpy(), G["vaetrain.grad.conv_in.weight"], rtol=1e-3,
                               atol=1e-7)
