"""Dry run of bench.py's main() on the CPU: the control flow of the default invocation — variant verdicts, eager warm-up,
profiling passes, capture + timed regions, the step-level A/B, tear-down, secondary configurations, the ONE JSON line —
executed end to end with the GPU-only pieces replaced by stand-ins (a toy engine, a fake captured step, fake CUDA events).
What it pins is the glue: names, ordering, the shape of the printed line; no kernel runs here."""
import json
import sys
import types
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


class _Event:
    _clock = [0.0]

    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        _Event._clock[0] += 7.0
        self.t = _Event._clock[0]

    def elapsed_time(self, other):
        return other.t - self.t


class _Engine(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = torch.nn.Linear(8, 8)
        self.first_stage_model = None

    def training_step(self, batch):
        x = batch["image"].float().mean(dim=(2, 3))[:, :1].expand(-1, 8)
        return self.model(x).pow(2).mean()


class _Graphed:
    launches_per_replay = 321
    made = 0
    variants_faster = False

    def __init__(self, eng, reducer, image, ctx, vec, warmup=1, optimizer=None, ema=None, pool=None):
        _Graphed.made += 1
        self.loss = torch.tensor(0.25)
        self.pool = pool or (0, _Graphed.made)

    def step(self, image=None, crossattn=None, vector=None, weights=None):
        from neurosis_b200._lib import lib
        _Event._clock[0] += 1.0 if (self.variants_faster and lib.nk_gemm_set_dual(-1)) else 2.0
        return self.loss


@pytest.fixture
def dry(monkeypatch):
    from neurosis_b200 import graph, ops
    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu") if a and a[0] == "cuda" else real_device(*a, **k))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "memory_reserved", lambda *a, **k: 0)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(ops, "refresh_weight_copies", lambda force=False: None)
    monkeypatch.setattr(ops, "invalidate_weight_cache", lambda: None)
    monkeypatch.setattr(graph, "GraphedTrainStep", _Graphed)
    monkeypatch.setattr(bench, "build_engine", lambda dev, seed=42, family="sdxl": _Engine())
    monkeypatch.setattr(bench, "cpu_reference_sample", lambda *a, **k: {"img_per_s": 0.01, "sec_per_sample_step": 1.0, "cores": 4,
                                                                        "kind": "reference", "sample": "stand-in"})
    child = ("import json; print(json.dumps({'metric': 'm', 'value': 3.0, 'unit': 'images/s', 'n_gpus': 1, 'steps': 2, 'warmup': 3,"
             " 'ms_per_step': 10.0, 'config': {'batch_per_gpu': 1, 'workload': 'w'}}))")
    monkeypatch.setattr(bench, "_child_cmd", lambda name, args, world: [sys.executable, "-c", child])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "NK_GEMM_DUAL", "NK_NORM_VARIANT", "NK_GEMM_EPI_PREFETCH", "NK_FUSED_CROSS_KV",
              "NK_BENCH_EXTRAS", "NK_BENCH_NO_STEP_GUARD"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("NK_B200_TUNE_CACHE", "0")
    monkeypatch.setattr(bench, "T_START", bench.time.monotonic())
    _Graphed.made = 0
    _Graphed.variants_faster = False
    yield
    from neurosis_b200._lib import lib
    for fn, v in (("nk_gemm_set_dual", 0), ("nk_gemm_set_dual_min_k", 0), ("nk_gemm_set_dual_skew", 0), ("nk_gemm_set_dual_classes", 7),
                  ("nk_norm_set_variant", 0), ("nk_gemm_set_epi_prefetch", 0)):
        getattr(lib, fn)(v)
    ops.FUSE_CROSS_KV = False
    ops.GRAD_SINK = None
    for k in ("NK_GEMM_DUAL", "NK_GEMM_DUAL_MIN_K", "NK_GEMM_DUAL_SKEW", "NK_GEMM_DUAL_CLASSES", "NK_NORM_VARIANT", "NK_GEMM_EPI_PREFETCH",
              "NK_FUSED_CROSS_KV"):
        bench.os.environ.pop(k, None)


def _run(monkeypatch, capsys, argv):
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines  # exactly ONE JSON line
    return json.loads(lines[0])


def test_default_invocation_without_accepted_variants(dry, monkeypatch, capsys):
    """no GPU -> the real probe child fails -> every variant stays off; the line carries the contract's keys, the verdicts and
    the three secondary configurations."""
    line = _run(monkeypatch, capsys, ["--steps", "2", "--warmup", "1"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "other_configs"):
        assert key in line, key
    assert line["metric"] == bench.METRIC and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    assert line["gpu_launches"] == 2 * _Graphed.launches_per_replay and _Graphed.made == 1
    tv = line["config"]["tuned_variants"]
    assert tv["enabled"] is False and "step_ab" not in tv
    assert all(tv[k]["enabled"] is False for k in bench.SUB_VARIANTS)
    assert {k for k in line["other_configs"] if not k.startswith("_")} == set(bench.OTHER_CONFIGS)
    assert all(line["other_configs"][k]["value"] == 3.0 for k in bench.OTHER_CONFIGS)
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 4


def test_accepted_variants_trigger_the_step_level_ab(dry, monkeypatch, capsys):
    """with a verdict that enables variants: the step is captured twice (with / without), both times are reported, the library
    ends in the state of the headline's kernels and the children inherit it through the environment."""
    from neurosis_b200._lib import lib
    verdict = {"variant": "gemm_row_tile_pairing", "ok": True, "enabled": True, "mode": 1, "min_k_iters": 20, "skew": 0, "classes": 5,
               "speedup": 1.1, "layernorm_column_owner": {"ok": True, "enabled": True, "mask": 4, "speedup": 1.3},
               "groupnorm_reverse_apply": {"enabled": False}, "epilogue_l2_prefetch": {"ok": True, "enabled": True, "mask": 1},
               "fused_cross_kv": {"enabled": False}}

    def fake_autotune(world, local, dev):
        bench._apply_tuned(verdict)
        return verdict

    monkeypatch.setattr(bench, "_autotune", fake_autotune)
    monkeypatch.setattr(bench, "_step_guard", lambda args, tuned, world, local, dev: tuned)
    seen = {}
    real = bench.run_other_configs

    def spy(args, world, rank, budget_s, per_config_s):
        seen.update({k: bench.os.environ.get(k) for k in ("NK_GEMM_DUAL", "NK_NORM_VARIANT", "NK_GEMM_EPI_PREFETCH", "NK_GEMM_DUAL_CLASSES")})
        seen["lib"] = (lib.nk_gemm_set_dual(-1), lib.nk_norm_set_variant(-1), lib.nk_gemm_set_epi_prefetch(-1))
        return real(args, world, rank, budget_s, per_config_s)

    monkeypatch.setattr(bench, "run_other_configs", spy)
    line = _run(monkeypatch, capsys, ["--steps", "2", "--warmup", "1"])
    ab = line["config"]["tuned_variants"]["step_ab"]
    assert _Graphed.made == 2 and {"ms_per_step_variants_on", "ms_per_step_variants_off", "headline_uses_variants"} <= set(ab)
    # the fake clock makes both measurements equally long: ties go to the measured kernels
    assert ab["headline_uses_variants"] is False
    assert seen["lib"] == (0, 0, 0) and seen["NK_GEMM_DUAL"] == "0" and seen["NK_NORM_VARIANT"] == "0"
    assert abs(line["ms_per_step"] - ab["ms_per_step_variants_off"]) < 1e-9


def test_variants_that_win_the_step_are_kept_for_the_headline_and_the_children(dry, monkeypatch, capsys):
    from neurosis_b200._lib import lib
    verdict = {"variant": "gemm_row_tile_pairing", "ok": True, "enabled": True, "mode": 1, "min_k_iters": 20, "skew": 3, "classes": 7,
               "speedup": 1.1, "layernorm_column_owner": {"enabled": False}, "groupnorm_reverse_apply": {"ok": True, "enabled": True},
               "epilogue_l2_prefetch": {"enabled": False}, "fused_cross_kv": {"ok": True, "enabled": True}}
    monkeypatch.setattr(bench, "_autotune", lambda world, local, dev: (bench._apply_tuned(verdict), verdict)[1])
    monkeypatch.setattr(bench, "_step_guard", lambda args, tuned, world, local, dev: tuned)
    _Graphed.variants_faster = True
    seen = {}
    real = bench.run_other_configs

    def spy(args, world, rank, budget_s, per_config_s):
        from neurosis_b200 import ops
        seen.update(env=(bench.os.environ.get("NK_GEMM_DUAL"), bench.os.environ.get("NK_GEMM_DUAL_SKEW"), bench.os.environ.get("NK_NORM_VARIANT"),
                         bench.os.environ.get("NK_FUSED_CROSS_KV")), lib=(lib.nk_gemm_set_dual(-1), lib.nk_norm_set_variant(-1)), xkv=ops.FUSE_CROSS_KV)
        return real(args, world, rank, budget_s, per_config_s)

    monkeypatch.setattr(bench, "run_other_configs", spy)
    line = _run(monkeypatch, capsys, ["--steps", "3", "--warmup", "1", "--no-cpu-baseline"])
    ab = line["config"]["tuned_variants"]["step_ab"]
    assert ab["headline_uses_variants"] is True and ab["ms_per_step_variants_on"] < ab["ms_per_step_variants_off"]
    assert abs(line["ms_per_step"] - ab["ms_per_step_variants_on"]) < 1e-9
    assert seen["env"] == ("1", "3", "2", "1") and seen["lib"] == (1, 2) and seen["xkv"] is True
    assert all(line["other_configs"][k]["tuned_variants_on"] is True for k in bench.OTHER_CONFIGS)


def test_secondary_configurations_can_be_switched_off(dry, monkeypatch, capsys):
    line = _run(monkeypatch, capsys, ["--steps", "1", "--warmup", "1", "--other-configs", "none", "--no-cpu-baseline"])
    assert "other_configs" not in line and "cpu_baseline" not in line


def test_guard_child_reports_every_stage(dry, monkeypatch, capsys):
    """`bench.py --guard-child` (the sacrificial process of the step guard) with every variant pinned on: one verdict per stage,
    flushed progressively, the last line carrying all of them plus the baseline repeat and the tolerances in use."""
    from neurosis_b200 import ops
    for k, v in (("NK_GEMM_DUAL", "1"), ("NK_GEMM_DUAL_MIN_K", "20"), ("NK_NORM_VARIANT", "7"), ("NK_GEMM_EPI_PREFETCH", "3"),
                 ("NK_FUSED_CROSS_KV", "1")):
        monkeypatch.setenv(k, v)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--guard-child", "--config", "sdxl", "--batch", "2"])
    bench.main()
    lines = [json.loads(ln) for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    last = lines[-1]
    assert {"baseline_repeat", "tolerance", "prefetch", "gemm", "cross_kv", "groupnorm", "layernorm"} <= set(last)
    assert last["prefetch"]["equal"] and last["gemm"]["equal"] and last["cross_kv"]["equal"] and last["groupnorm"]["equal"]
    assert last["layernorm"]["agree"] and last["tolerance"]["loss"] >= 5e-4
    assert len(lines) >= 5 and "gemm" not in lines[0]          # progressive: the first flushed line only has the first stage
    assert ops.FUSE_CROSS_KV is True                            # (state of the child process after its last stage)


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass


class _CUDAGraph:
    def replay(self):
        pass

    def pool(self):
        return (0, 0)


@pytest.fixture
def dry_streams(dry, monkeypatch):
    import contextlib
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _CUDAGraph)
    monkeypatch.setattr(torch.cuda, "graph", lambda g, pool=None: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "max_memory_allocated", lambda *a, **k: 0)
    yield


def test_bucket_configuration_dry_run(dry_streams, monkeypatch, capsys):
    """--config buckets (BASELINE.json configs[3]): one captured step per aspect bucket, per-step tag-frequency weights,
    the square-only comparison that EVERY rank must run — control flow and the shape of its line."""
    line = _run(monkeypatch, capsys, ["--config", "buckets", "--steps", "2", "--warmup", "1", "--batch", "1"])
    assert line["metric"] == bench.CONFIGS["buckets"]["metric"] and line["n_gpus"] == 1 and line["steps"] == 2
    assert _Graphed.made == len(line["config"]["buckets_wh"]) == 3
    assert line["config"]["square_only_ms_per_step"] is not None and "tuned_variants" in line["config"]
    assert len(line["config"]["buckets_drawn_rank0"]) == 2 and line["e2e"]["d2h_bytes_per_step"] == 4


def test_vae_configuration_dry_run(dry_streams, monkeypatch, capsys):
    """--config vae (BASELINE.json configs[4]) with a toy autoencoder: eager warm-up on a side stream, capture, the three
    timed regions (resident, host batches, encode only)."""
    from neurosis_b200.modules import vae

    class ToyAE(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(3))

        def training_step(self, batch):
            return (batch["image"].mean(dim=(2, 3)) * self.w).pow(2).mean()

        def encode(self, x):
            return x[:, :, ::8, ::8]

    monkeypatch.setattr(vae, "AutoencoderKL", ToyAE)
    monkeypatch.setitem(bench.CONFIGS["vae"], "px", 64)
    line = _run(monkeypatch, capsys, ["--config", "vae", "--steps", "2", "--warmup", "1", "--batch", "2"])
    assert line["metric"] == bench.CONFIGS["vae"]["metric"] and line["config"]["cuda_graph"] is True
    assert line["config"]["encode_images_per_s"] > 0 and "tuned_variants" in line["config"]
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 64 * 64 * 4


@pytest.mark.parametrize("mode", ["plain", "variants"])
def test_two_rank_dry_run_under_torchrun(mode):
    """the default invocation as the driver launches it for N = 2 (torch.distributed.run, one process per rank), on the CPU
    with gloo: the agreement collectives, the A/B decision (identical on both ranks), the children, the final barrier and
    the exit — exactly one JSON line, printed by rank 0.  'variants': rank 1 rejected one variant, so both ranks drop it."""
    import subprocess
    helper = Path(__file__).resolve().parent / "helpers" / "bench_dry_rank.py"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29677" if mode == "plain" else "29679", str(helper), mode],
                       capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{"metric"')]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["global_batch"] == 2 * line["config"]["batch_per_gpu"]
    assert "cpu_baseline" not in line  # N = 1 only
    assert {k for k in line["other_configs"] if not k.startswith("_")} == set(bench.OTHER_CONFIGS)
    assert all(line["other_configs"][k]["n_gpus"] == 2 for k in bench.OTHER_CONFIGS)
    tv = line["config"]["tuned_variants"]
    if mode == "plain":
        assert tv["enabled"] is False
    else:
        assert tv["enabled"] is True and tv["layernorm_column_owner"]["enabled"] is False
        assert tv["step_ab"]["headline_uses_variants"] is True
        assert all(line["other_configs"][k]["workload"] == "1" for k in bench.OTHER_CONFIGS)  # children inherit NK_GEMM_DUAL=1
