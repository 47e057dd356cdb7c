"""Host logic of neurosis_b200.tune (no GPU): the mode switch of the C ABI, pinned / disabled environments, and — the
property that matters — a probe that cannot run (no device, crash, timeout) leaves the library on its default kernels."""
import json
import sys

import pytest

from neurosis_b200 import tune
from neurosis_b200._lib import lib


def test_mode_switch_roundtrip_through_the_c_abi():
    start = lib.nk_gemm_set_dual(-1)  # out of range: query only
    try:
        assert lib.nk_gemm_set_dual(1) == start
        assert lib.nk_gemm_set_dual(-1) == 1
        assert lib.nk_gemm_set_dual(7) == 1 and lib.nk_gemm_set_dual(-1) == 1  # ignored
        assert tune.apply(2) == 1 and tune.apply(0) == 2
    finally:
        lib.nk_gemm_set_dual(start if 0 <= start <= 2 else 0)


def test_pinned_and_disabled_environments_skip_the_probe(monkeypatch):
    monkeypatch.setenv("NK_GEMM_DUAL", "1")
    rep = tune.autotune()
    assert rep["enabled"] and rep["mode"] == 1 and "pinned" in rep["source"]
    monkeypatch.setenv("NK_GEMM_DUAL", "0")
    assert not tune.autotune()["enabled"]
    monkeypatch.delenv("NK_GEMM_DUAL")
    monkeypatch.setenv("NK_B200_TUNE", "0")
    rep = tune.autotune()
    assert not rep["enabled"] and rep["mode"] == 0


def test_failed_probe_leaves_the_default_kernels(monkeypatch):
    """no CUDA device here: the child exits with an error -> not enabled, mode 0, the reason is reported."""
    monkeypatch.delenv("NK_GEMM_DUAL", raising=False)
    monkeypatch.delenv("NK_B200_TUNE", raising=False)
    lib.nk_gemm_set_dual(2)
    rep = tune.autotune(timeout_s=240)
    assert rep["enabled"] is False and rep["mode"] == 0 and "error" in rep
    assert lib.nk_gemm_set_dual(-1) == 0
    json.dumps(tune._summary(rep))  # what bench.py prints must serialise


def test_verdict_rules(monkeypatch):
    """enabled only if every comparison passed AND the weighted time went down by >= 1 %."""
    import subprocess

    class FakeProc:
        def __init__(self, rep):
            self.rep, self.returncode, self.pid = rep, 0, 0

        def communicate(self, timeout=None):
            return "noise\n" + json.dumps(self.rep) + "\n", ""

    monkeypatch.delenv("NK_GEMM_DUAL", raising=False)
    monkeypatch.delenv("NK_B200_TUNE", raising=False)
    cases = [({"ok": True, "speedup": 1.08, "checks": [], "timings": []}, True),
             ({"ok": True, "speedup": 1.001, "checks": [], "timings": []}, False),
             ({"ok": False, "speedup": 1.5, "checks": [{"kind": "conv", "dims": [1], "ok": False, "err": 3.0}], "timings": []}, False)]
    for rep, want in cases:
        monkeypatch.setattr(subprocess, "Popen", lambda *a, _r=rep, **k: FakeProc(_r))
        got = tune.autotune()
        assert got["enabled"] is want and lib.nk_gemm_set_dual(-1) == (1 if want else 0)
    lib.nk_gemm_set_dual(0)


def test_probe_shapes_cover_every_paired_mode():
    kinds = {k for k, _ in tune.CHECK_SHAPES} | {k for k, _, _ in tune.TIMED_SHAPES}
    assert {"linear_fwd", "linear_fwd_f32", "linear_dgrad", "linear_wgrad", "linear_wgrad_acc", "conv", "conv_s2"} <= kinds
    # odd pair-tile counts (the half-empty last pair) are present for a matrix and for an image operand
    assert any(k == "linear_fwd" and ((d[0] + 127) // 128 + 1) // 2 % 2 == 1 for k, d in tune.CHECK_SHAPES)
    assert any(k == "conv" and d[0] * d[1] * d[2] < 128 * 2 for k, d in tune.CHECK_SHAPES)
