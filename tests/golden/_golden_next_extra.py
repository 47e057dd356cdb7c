# Rows 2-4 of make_golden_next.py (executed by it with `out`, `np`, `torch`, `synth_tensor` in scope).
from neurosis.modules.ema import LitEma
from neurosis.modules.encoders.metadata import ConcatTimestepEmbedderND
from neurosis.optimizers import Adafactor

# ---- row 2: Adafactor, three steps over parameters of every layout class -------------------------------------
OPT_SHAPES = {"lin": (96, 200), "lin_ragged": (130, 70), "conv3": (24, 16, 3, 3), "conv1": (8, 12, 1, 1),
              "bias": (300,), "stack": (2, 70, 40)}
OPT_CASES = {
    "yaml": dict(scale_parameter=True, relative_step=True, warmup_init=True),   # configs/sdxl/sdxl.example.yaml:160-164
    "ext": dict(lr=1e-3, scale_parameter=False, relative_step=False, warmup_init=False, beta1=0.9, weight_decay=0.01,
                clip_threshold=0.5),
}
for case, kw in OPT_CASES.items():
    params = {k: torch.nn.Parameter(synth_tensor(f"opt.p.{k}", s, scale=0.05)) for k, s in OPT_SHAPES.items()}
    opt = Adafactor(list(params.values()), **kw)
    for step in range(3):
        for k, p in params.items():
            p.grad = synth_tensor(f"opt.g.{k}.{step}", tuple(p.shape), scale=0.02 * (step + 1))
        opt.step()
    for k, p in params.items():
        out[f"opt.{case}.{k}.p"] = p.detach().numpy().copy()
        st = opt.state[p]
        out[f"opt.{case}.{k}.rms"] = np.array(float(st["RMS"]))
        for sk in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq", "exp_avg"):
            if sk in st:
                out[f"opt.{case}.{k}.{sk}"] = st[sk].numpy().copy()

# ---- row 4: LitEma, 12 updates (crosses the (1+n)/(10+n) warm-up) ------------------------------------------------
class _M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(40, 30)
        self.b = torch.nn.Conv2d(8, 8, 3)

m = _M()
with torch.no_grad():
    for n_, p in m.named_parameters():
        p.copy_(synth_tensor(f"ema.p.{n_}", tuple(p.shape)))
ema = LitEma(m, decay=0.9999)
for it in range(12):
    with torch.no_grad():
        for n_, p in m.named_parameters():
            p.add_(synth_tensor(f"ema.d.{n_}.{it}", tuple(p.shape), scale=0.1))
    ema(m)
for n_, _ in m.named_parameters():
    out[f"ema.{n_}"] = dict(ema.named_buffers())[n_.replace(".", "_")].numpy().copy()
out["ema.num_updates"] = np.array(int(ema.num_updates))

# ---- row 3: ConcatTimestepEmbedderND (SDXL size / crop conditioning) ---------------------------------------------
emb = ConcatTimestepEmbedderND(256)
sizes = torch.tensor([[1024.0, 1024.0], [1152.0, 896.0], [832.0, 1216.0], [0.0, 64.0]])
out["cond.sizes"] = sizes.numpy()
out["cond.fourier"] = emb(sizes).numpy()

# ---- §8(a) rows 4-6 with the rectified-flow objective (loss.py:130-139: z_t = (1-s) x + s n, target = noise, "F" output)
from neurosis.modules.diffusion import OpenAIWrapper, StandardDiffusionLoss, UNetModel
from neurosis.modules.diffusion.denoiser import Denoiser as _Denoiser
from neurosis.modules.diffusion.denoiser_preconditioning import RectifiedFlowComfyPreconditioning
from neurosis.modules.diffusion.denoiser_weighting import RectifiedFlowComfyWeighting

from common import TINY_SDXL
from oracle.unet import unet_param_shapes as _ushapes

_cfg = TINY_SDXL
_sd = synth_state_dict(_ushapes(_cfg), seed=1)
_ref = UNetModel(**_cfg)
_ref.load_state_dict(_sd)
_sig = torch.tensor([0.3, 0.8])


class _Fixed:
    def __call__(self, n, t=None):
        return _sig


class _Cond(torch.nn.Module):
    def forward(self, batch):
        return {"crossattn": batch["ctx"], "vector": batch["vec"]}


_lat, _noise = synth_tensor("step.latent", (2, 4, 16, 16)), synth_tensor("step.noise", (2, 4, 16, 16))
_loss_fn = StandardDiffusionLoss(sigma_generator=_Fixed(), loss_weighting=RectifiedFlowComfyWeighting(),
                                 objective_type="rf")
_orig_rl = torch.randn_like
torch.randn_like = lambda t_, **kw: _noise.to(t_)
try:
    _loss = _loss_fn(OpenAIWrapper(_ref), _Denoiser(RectifiedFlowComfyPreconditioning()), _Cond(), _lat,
                     {"ctx": synth_tensor("sdxl.ctx", (2, 77, _cfg["context_dim"])),
                      "vec": synth_tensor("sdxl.y", (2, _cfg["adm_in_channels"]))})
finally:
    torch.randn_like = _orig_rl
_loss.mean().backward()
out["rf.sigmas"] = _sig.numpy()
out["rf.loss"] = _loss.detach().double().numpy()
out["rf.grad.out.2.weight"] = _ref.get_parameter("out.2.weight").grad.numpy()
out["rf.grad_l2"] = np.array([_ref.get_parameter(n).grad.norm().item() for n in sorted(_sd)], dtype=np.float64)

# ---- aspect-bucket (non-square) latents through the reference UNet: pins the oracle on the shapes the GPU test
# tests/test_gpu_modules.py::test_unet_aspect_bucket_shapes_vs_oracle uses -----------------------------------------
for _h, _w in ((24, 16), (12, 20)):
    _ref.zero_grad()
    _o = _ref(synth_tensor("bucket.x", (2, 4, _h, _w)), torch.tensor([3, 977]),
              synth_tensor("bucket.ctx", (2, 77, _cfg["context_dim"])), synth_tensor("bucket.y", (2, _cfg["adm_in_channels"])))
    (_o * synth_tensor("bucket.gout", (2, 4, _h, _w), scale=0.1)).sum().backward()
    out[f"bucket.{_h}x{_w}.out"] = _o.detach().numpy()
    out[f"bucket.{_h}x{_w}.grad_l2"] = np.array([_ref.get_parameter(n).grad.norm().item() for n in sorted(_sd)],
                                                 dtype=np.float64)

# ---- row 3: GeneralConditioner assembly with per-sample UCG masks (embedding.py:90-149), fixed torch seed ---------
from neurosis.modules.encoders.embedding import GeneralConditioner as _GC
from neurosis.modules.encoders.misc import IdentityEncoder as _IdE

_gc = _GC([_IdE(input_key="ctx"), _IdE(input_key="pooled", ucg_rate=0.3),
           ConcatTimestepEmbedderND(256, input_key="original_size_as_tuple"),
           ConcatTimestepEmbedderND(256, input_key="crop_coords_top_left", ucg_rate=0.5),
           ConcatTimestepEmbedderND(256, input_key="target_size_as_tuple")])
_gb = {"image": torch.zeros(6, 3, 8, 8), "ctx": synth_tensor("gc.ctx", (6, 77, 32)),
       "pooled": synth_tensor("gc.pooled", (6, 48)),
       "original_size_as_tuple": [(1024, 1024), (1152, 896), (832, 1216), (1024, 1024), (640, 1536), (512, 512)],
       "crop_coords_top_left": [(0, 0), (16, 0), (0, 32), (8, 8), (0, 0), (64, 64)],
       "target_size_as_tuple": [(1024, 1024), (896, 1152), (1216, 832), (1024, 1024), (1536, 640), (512, 512)]}
torch.manual_seed(1234)
_go = _gc(_gb)
out["gc.vector"], out["gc.crossattn"] = _go["vector"].numpy(), _go["crossattn"].numpy()
torch.manual_seed(1234)
_go = _gc(_gb, force_zero_embeddings=["pooled", "target_size_as_tuple"])
out["gc.vector_force_zero"] = _go["vector"].numpy()

# ---- StandardDiffusionLoss(loss_type="l1") (loss.py:89-93, losses/functions.py:65-78), edm objective, Eps weighting ----
from neurosis.modules.diffusion import DiscreteDenoiser as _DDen
from neurosis.modules.diffusion import EpsPreconditioning as _EpsP
from neurosis.modules.diffusion import EpsWeighting as _EpsW
from neurosis.modules.diffusion import LegacyDDPMDiscretization as _LD

_den = _DDen(_EpsP(), 1000, _LD())
_den.sigmas, _den.log_sigmas = _den.sigmas.detach(), _den.log_sigmas.detach()
_sig1 = _den.sigmas[torch.tensor([800, 300])].clone()  # descending table: [800] small sigma, [300] large


class _Fixed1:
    def __call__(self, n, t=None):
        return _sig1


_ref.zero_grad()
_l1 = StandardDiffusionLoss(sigma_generator=_Fixed1(), loss_weighting=_EpsW(), loss_type="l1")
torch.randn_like = lambda t_, **kw: _noise.to(t_)
try:
    _loss1 = _l1(OpenAIWrapper(_ref), _den, _Cond(), _lat,
                 {"ctx": synth_tensor("sdxl.ctx", (2, 77, _cfg["context_dim"])),
                  "vec": synth_tensor("sdxl.y", (2, _cfg["adm_in_channels"]))})
finally:
    torch.randn_like = _orig_rl
_loss1.mean().backward()
out["l1.sigmas"] = _sig1.numpy()
out["l1.loss"] = _loss1.detach().double().numpy()
out["l1.grad_l2"] = np.array([_ref.get_parameter(n).grad.norm().item() for n in sorted(_sd)], dtype=np.float64)
