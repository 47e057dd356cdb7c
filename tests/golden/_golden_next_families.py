# §8(a) rows 6-8, all shipped variants (executed by make_golden_next.py with `out`, `np`, `torch`, `synth_tensor` in
# scope): every Discretization, SigmaGenerator, DenoiserPreconditioning and DenoiserWeighting class of the reference with
# its default constructor arguments, evaluated on fixed inputs.  Keys: fam.<kind>.<ClassName>[.<what>].
import inspect

from neurosis.modules.diffusion import denoiser_preconditioning as _P
from neurosis.modules.diffusion import denoiser_weighting as _W
from neurosis.modules.diffusion import discretization as _D
from neurosis.modules.diffusion.sampling import sigma_generators as _S


def _classes(mod, base):
    return sorted((n, c) for n, c in vars(mod).items()
                  if inspect.isclass(c) and issubclass(c, base) and c is not base and c.__module__ == mod.__name__)


FAM_ERRORS = {}
for name, cls in _classes(_D, _D.Discretization):
    for n in (1000, 40):
        for flip in (False, True):
            try:
                out[f"fam.disc.{name}.{n}.{int(flip)}"] = cls()(n, flip=flip).detach().double().numpy()
            except Exception as e:  # reference crashes are recorded, not reproduced
                FAM_ERRORS[f"fam.disc.{name}.{n}.{int(flip)}"] = type(e).__name__

sig = (synth_tensor("fam.sigma", (16,), uniform=True).abs() * 14.0 + 0.02).float()
sig01 = (synth_tensor("fam.sigma01", (16,), uniform=True).abs() * 0.98 + 0.01).float()
out["fam.sigma"], out["fam.sigma01"] = sig.numpy(), sig01.numpy()
for name, cls in _classes(_P, _P.DenoiserPreconditioning):
    s = sig01 if "RectifiedFlow" in name else sig
    try:
        res = cls()(s)
        for i, r in enumerate(res):
            out[f"fam.precond.{name}.{i}"] = torch.as_tensor(r).detach().double().numpy()
    except Exception as e:
        FAM_ERRORS[f"fam.precond.{name}"] = type(e).__name__
for name, cls in _classes(_W, _W.DenoiserWeighting):
    s = sig01 if "RectifiedFlow" in name else sig
    try:
        obj = cls(_W.EpsWeighting()) if name == "MinSNRGammaModifier" else cls()
        out[f"fam.weight.{name}"] = torch.as_tensor(obj(s)).detach().double().numpy()
    except Exception as e:
        FAM_ERRORS[f"fam.weight.{name}"] = type(e).__name__

t = torch.linspace(0.02, 0.97, 16, dtype=torch.float64)
out["fam.t"] = t.numpy()
for name, cls in _classes(_S, _S.SigmaGenerator):
    try:
        gen = cls(_D.LegacyDDPMDiscretization(), 1000) if name == "DiscreteSigmaGenerator" else cls()
        if hasattr(gen, "sigmas") and torch.is_tensor(gen.sigmas):
            gen.sigmas = gen.sigmas.detach()
        out[f"fam.gen.{name}.t"] = torch.as_tensor(gen(16, t.clone())).detach().double().numpy()
        torch.manual_seed(7)
        out[f"fam.gen.{name}.rand"] = torch.as_tensor(gen(16, None)).detach().double().numpy()
    except Exception as e:
        FAM_ERRORS[f"fam.gen.{name}"] = type(e).__name__
out["fam.errors"] = np.array(sorted(f"{k}={v}" for k, v in FAM_ERRORS.items()))
print("family goldens:", sum(1 for k in out if k.startswith("fam.")), "arrays; reference errors:", FAM_ERRORS)

# ---- DiscreteDenoiser option combinations (denoiser.py:60-97): quantised sigma and the c_noise the UNet receives ----
from neurosis.modules.diffusion.denoiser import DiscreteDenoiser as _DD

for _q in (True, False):
    for _flip in (False, True):
        for _pname in ("EpsPreconditioning", "VPreconditioning"):
            _den = _DD(getattr(_P, _pname)(), 1000, _D.LegacyDDPMDiscretization(), quantize_c_noise=_q, flip=_flip)
            _den.sigmas, _den.log_sigmas = _den.sigmas.detach(), _den.log_sigmas.detach()
            _s = _den.possibly_quantize_sigma(sig)
            _cn = _den.possibly_quantize_c_noise(_den.preconditioning(_s)[3])
            out[f"fam.dd.{_pname}.{int(_q)}.{int(_flip)}.sigma"] = _s.double().numpy()
            out[f"fam.dd.{_pname}.{int(_q)}.{int(_flip)}.c_noise"] = _cn.double().numpy()
