"""bench.py's default run also measures BASELINE.json configs[1] / [3] / [4] as child processes after the headline
measurement (`run_other_configs`).  These CPU tests drive that orchestration with stand-in children: a child that
answers, one that crashes, one that hangs (killed with its process group at the limit) — none may cost the parent
its own line."""
import json
import sys
import time
import types
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def _args(steps=2, warmup=1):
    return types.SimpleNamespace(steps=steps, warmup=warmup)


def test_children_answer_crash_and_hang(monkeypatch):
    ok_line = {"metric": "m", "value": 12.5, "unit": "images/s", "n_gpus": 1, "steps": 2, "warmup": 3, "ms_per_step": 80.0,
               "e2e": {"value": 12.0, "unit": "images/s", "h2d_bytes_per_step": 10, "d2h_bytes_per_step": 4},
               "config": {"batch_per_gpu": 32, "workload": "w", "encode_images_per_s": 99.0}, "gpu_launches": 7,
               "clocks": {"sm_mhz": 1900.0, "reasons": ["sw_power_cap"]}, "roofline": {"achieved": 500.0}}

    def cmd(name, args, world):
        if name == "sd15":  # noise on stdout before the line, as a warning would be
            return [sys.executable, "-c", f"print('warming up'); print({json.dumps(ok_line)!r})"]
        if name == "buckets":
            return [sys.executable, "-c", "import sys; sys.stderr.write('boom: device-side assert\\n'); sys.exit(3)"]
        return [sys.executable, "-c", "import time; time.sleep(60)"]

    monkeypatch.setattr(bench, "_child_cmd", cmd)
    t0 = time.monotonic()
    out = bench.run_other_configs(_args(), world=1, rank=0, budget_s=100.0, per_config_s=2.0)
    assert time.monotonic() - t0 < 20.0
    assert out["sd15"]["value"] == 12.5 and out["sd15"]["batch_per_gpu"] == 32 and out["sd15"]["encode_images_per_s"] == 99.0
    assert out["sd15"]["clock_reasons"] == ["sw_power_cap"] and out["sd15"]["e2e"]["value"] == 12.0
    assert "exit 3" in out["buckets"]["error"] and "boom" in out["buckets"]["stderr_tail"]
    assert "timeout" in out["vae"]["error"]


def test_budget_exhaustion_skips_remaining_and_other_ranks_stay_silent(monkeypatch):
    monkeypatch.setattr(bench, "_child_cmd", lambda name, args, world: [sys.executable, "-c", "print('{\"value\": 1}')"])
    out = bench.run_other_configs(_args(), world=2, rank=1, budget_s=100.0, per_config_s=5.0)
    assert out == {}  # only rank 0 collects
    out = bench.run_other_configs(_args(), world=1, rank=0, budget_s=10.0, per_config_s=5.0)
    assert all("skipped" in v["error"] for v in out.values())


def test_children_get_their_own_rendezvous(monkeypatch):
    """the children of the N ranks must not talk to the launcher's agent store: TORCHELASTIC_* is stripped and every
    configuration gets its own port, identical on all ranks."""
    seen = []

    def cmd(name, args, world):
        return [sys.executable, "-c", "import os, json; print(json.dumps({'value': 1, 'metric': 'm', 'unit': 'u', 'n_gpus': 2, "
                "'steps': 1, 'warmup': 1, 'ms_per_step': 1.0, 'config': {'workload': os.environ['MASTER_PORT'] + ':' + "
                "os.environ['MASTER_ADDR'] + ':' + str('TORCHELASTIC_USE_AGENT_STORE' in os.environ) + ':' + "
                "os.environ['NK_BENCH_EXTRAS']}}))"]

    monkeypatch.setattr(bench, "_child_cmd", cmd)
    monkeypatch.setenv("MASTER_PORT", "29613")
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("TORCHELASTIC_USE_AGENT_STORE", "True")
    out = bench.run_other_configs(_args(), world=2, rank=0, budget_s=100.0, per_config_s=10.0)
    ports = set()
    for name in bench.OTHER_CONFIGS:
        port, addr, agent, extras = out[name]["workload"].split(":")
        assert addr == "127.0.0.1" and agent == "False" and extras == "0"
        assert 20000 <= int(port) < 40000 and int(port) != 29613
        ports.add(port)
    assert len(ports) == len(bench.OTHER_CONFIGS)


def test_two_ranks_children_rendezvous_under_torchrun():
    """the real launch shape: torchrun starts 2 parents (own process group, agent store), each spawns one child per
    configuration; the children find each other on the fresh port and the parents' group is left alone."""
    import subprocess
    helpers = Path(__file__).resolve().parent / "helpers"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29641", str(helpers / "orch_parent.py")],
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert len(lines) == 1
    out = json.loads(lines[0][len("RESULT "):])
    for name in bench.OTHER_CONFIGS:
        assert out[name]["value"] == 3.0 and out[name]["n_gpus"] == 2, out[name]
    assert len({out[name]["workload"] for name in bench.OTHER_CONFIGS}) == 3  # one port per configuration


def test_configuration_that_fails_with_tuned_variants_is_retried_on_the_measured_kernels(monkeypatch):
    """the kernel variants are validated on the SDXL step only; a secondary configuration whose child dies with them
    pinned on is run again with every variant off, and both outcomes are reported."""
    code = ("import os, sys, json\n"
            "if os.environ.get('NK_GEMM_DUAL', '0') != '0': sys.stderr.write('trap\\n'); sys.exit(134)\n"
            "print(json.dumps({'metric': 'm', 'value': 7.0, 'unit': 'u', 'n_gpus': 1, 'steps': 1, 'warmup': 1, 'ms_per_step': 1.0,"
            " 'config': {'workload': os.environ['MASTER_PORT']}}))")
    monkeypatch.setattr(bench, "_child_cmd", lambda name, args, world: [sys.executable, "-c", code])
    monkeypatch.setenv("NK_GEMM_DUAL", "1")
    monkeypatch.setenv("MASTER_PORT", "29500")
    out = bench.run_other_configs(_args(), world=1, rank=0, budget_s=200.0, per_config_s=20.0)
    for name in bench.OTHER_CONFIGS:
        assert out[name]["value"] == 7.0 and out[name]["tuned_variants_on"] is False
        assert "exit 134" in out[name + "_with_tuned_variants"]["error"]
    # retry ports differ from first-attempt ports and from each other
    assert len({out[n]["workload"] for n in bench.OTHER_CONFIGS}) == 3
    monkeypatch.setenv("NK_GEMM_DUAL", "0")  # nothing tuned: a failure is final, no second attempt
    monkeypatch.setattr(bench, "_child_cmd", lambda name, args, world: [sys.executable, "-c", "import sys; sys.exit(3)"])
    out = bench.run_other_configs(_args(), world=1, rank=0, budget_s=200.0, per_config_s=20.0)
    assert all("error" in out[n] for n in bench.OTHER_CONFIGS) and not any(k.endswith("_with_tuned_variants") for k in out)


def test_step_guard_keeps_or_drops_variants_and_sets_the_library(monkeypatch):
    monkeypatch.setenv("NK_B200_TUNE_CACHE", "0")
    """parent side of the child-process step guard: verdicts of the child -> `tuned`, library modes and the exported
    environment; a child that dies (no verdict) drops every variant."""
    from neurosis_b200._lib import lib

    class Fake:
        def __init__(self, out, rc=0):
            self.out, self.returncode, self.pid = out, rc, 0

        def communicate(self, timeout=None):
            return self.out, "boom"

    def tuned():
        return {"enabled": True, "mode": 1, "min_k_iters": 20, "skew": 3,
                "layernorm_column_owner": {"enabled": True, "speedup": 1.3, "mask": 5}, "epilogue_l2_prefetch": {"enabled": True, "speedup": 1.1, "mask": 1},
                "groupnorm_reverse_apply": {"enabled": True, "speedup": 1.05}, "fused_cross_kv": {"enabled": True, "speedup": 1.2}}

    args = types.SimpleNamespace(config="sdxl", batch=16)
    both = json.dumps({"prefetch": {"equal": True}, "groupnorm": {"equal": True}, "cross_kv": {"equal": True}, "gemm": {"equal": True, "loss_unpaired": 1.0, "loss_paired": 1.0},
                       "layernorm": {"agree": True, "loss_old": 1.0, "loss_new": 1.0005}})
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: Fake("warm\n" + both + "\n"))
    t = bench._step_guard(args, tuned(), 1, 0, None)
    assert t["enabled"] and t["layernorm_column_owner"]["enabled"] and t["step_guard"]["equal"]
    assert t["epilogue_l2_prefetch"]["enabled"] and lib.nk_gemm_set_epi_prefetch(-1) == 1 and bench.os.environ["NK_GEMM_EPI_PREFETCH"] == "1"
    from neurosis_b200 import ops
    assert t["fused_cross_kv"]["enabled"] and ops.FUSE_CROSS_KV is True and bench.os.environ["NK_FUSED_CROSS_KV"] == "1"
    assert (lib.nk_gemm_set_dual(-1), lib.nk_gemm_set_dual_min_k(-1), lib.nk_gemm_set_dual_skew(-1), lib.nk_norm_set_variant(-1)) == (1, 20, 3, 7)
    assert (bench.os.environ["NK_GEMM_DUAL"], bench.os.environ["NK_GEMM_DUAL_MIN_K"], bench.os.environ["NK_NORM_VARIANT"]) == ("1", "20", "7")
    # the GEMM stage passed and was flushed, then the LayerNorm stage took the child down
    first = json.dumps({"prefetch": {"equal": False}, "gemm": {"equal": True}})
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: Fake(first + "\n", rc=-6))
    t = bench._step_guard(args, tuned(), 1, 0, None)
    assert t["enabled"] and not t["layernorm_column_owner"]["enabled"] and not t["epilogue_l2_prefetch"]["enabled"]
    assert lib.nk_gemm_set_epi_prefetch(-1) == 0 and bench.os.environ["NK_GEMM_EPI_PREFETCH"] == "0"
    assert not t["fused_cross_kv"]["enabled"] and ops.FUSE_CROSS_KV is False and "exit -6" in t["fused_cross_kv"]["step_guard"]["error"]
    assert (lib.nk_gemm_set_dual(-1), lib.nk_norm_set_variant(-1), bench.os.environ["NK_NORM_VARIANT"]) == (1, 0, "0")
    # the step disagrees with pairing on
    bad = json.dumps({"gemm": {"equal": False, "loss_unpaired": 1.0, "loss_paired": 1.7}, "layernorm": {"agree": True}})
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: Fake(bad + "\n"))
    t = bench._step_guard(args, tuned(), 1, 0, None)
    assert not t["enabled"] and t["layernorm_column_owner"]["enabled"] and lib.nk_gemm_set_dual(-1) == 0
    assert not t["groupnorm_reverse_apply"]["enabled"] and lib.nk_norm_set_variant(-1) == 5  # no groupnorm verdict -> dropped
    # no verdict at all
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: Fake("", rc=-9))
    t = bench._step_guard(args, tuned(), 1, 0, None)
    assert not t["enabled"] and not t["layernorm_column_owner"]["enabled"] and "error" in t["step_guard"]
    assert (lib.nk_gemm_set_dual(-1), lib.nk_norm_set_variant(-1), bench.os.environ["NK_GEMM_DUAL"]) == (0, 0, "0")
    # nothing accepted by the probe: no child is started
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: (_ for _ in ()).throw(AssertionError("no child expected")))
    t = bench._step_guard(args, {"enabled": False, "layernorm_column_owner": {"enabled": False}}, 1, 0, None)
    assert not t["enabled"]
    ops.FUSE_CROSS_KV = False
    for k in ("NK_GEMM_DUAL", "NK_GEMM_DUAL_MIN_K", "NK_GEMM_DUAL_SKEW", "NK_GEMM_DUAL_CLASSES", "NK_NORM_VARIANT", "NK_GEMM_EPI_PREFETCH",
              "NK_FUSED_CROSS_KV"):
        bench.os.environ.pop(k, None)
    lib.nk_gemm_set_dual_min_k(0)
    lib.nk_gemm_set_dual_skew(0)
