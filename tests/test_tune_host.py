"""Host logic of neurosis_b200.tune (no GPU): the mode switch of the C ABI, pinned / disabled environments, and — the
property that matters — a probe that cannot run (no device, crash, timeout) leaves the library on its default kernels."""
import json
import sys

import pytest

from neurosis_b200 import tune
from neurosis_b200._lib import lib


@pytest.fixture(autouse=True)
def _no_verdict_cache(monkeypatch):
    """the verdict cache (tune.cache_*) is exercised by its own test; everywhere else every call must probe."""
    monkeypatch.setenv("NK_B200_TUNE_CACHE", "0")


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
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine WITHOUT a CUDA device (the probe must fail)")
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
    V = {"variant": "gemm_row_tile_pairing"}
    cases = [({**V, "ok": True, "speedup": 1.08, "min_k_iters": 20, "checks": [], "timings": []}, True),
             ({**V, "ok": True, "speedup": 1.001, "min_k_iters": 10, "checks": [], "timings": []}, False),
             ({**V, "ok": True, "speedup": 0.0, "min_k_iters": None, "checks": [], "timings": []}, False),
             ({**V, "ok": False, "speedup": 1.5, "min_k_iters": 10,
               "checks": [{"kind": "conv", "dims": [1], "ok": False, "err": 3.0}], "timings": []}, False)]
    for rep, want in cases:
        monkeypatch.setattr(subprocess, "Popen", lambda *a, _r=rep, **k: FakeProc(_r))
        got = tune.autotune()
        assert got["enabled"] is want and lib.nk_gemm_set_dual(-1) == (1 if want else 0)
        assert lib.nk_gemm_set_dual_min_k(-1) == (rep["min_k_iters"] if want else 0)
    # several candidates: the fastest one that passed wins, and its skew is applied
    class Multi(FakeProc):
        def communicate(self, timeout=None):
            return "\n".join(json.dumps(r) for r in self.rep) + "\n", ""
    L = {"variant": "layernorm_column_owner", "checks": [], "timings": []}
    reps = [{**V, "ok": True, "skew": 0, "speedup": 1.05, "min_k_iters": 20, "checks": [], "timings": []},
            {**L, "ok": True, "speedup": 1.4, "mask": 5},
            {**V, "ok": True, "skew": 3, "speedup": 1.09, "min_k_iters": 10, "checks": [], "timings": []}]
    monkeypatch.setattr(subprocess, "Popen", lambda *a, **k: Multi(reps))
    got = tune.autotune()
    assert got["enabled"] and got["skew"] == 3 and lib.nk_gemm_set_dual_skew(-1) == 3 and lib.nk_gemm_set_dual_min_k(-1) == 10
    # the LayerNorm verdict is independent of the GEMM one
    assert got["layernorm_column_owner"]["enabled"] and lib.nk_norm_set_variant(-1) == 5
    json.dumps(tune._summary(got))
    reps[1] = {**L, "ok": True, "speedup": 1.2, "mask": 4}   # only the backward form gained: only its bit is set
    assert tune.autotune()["layernorm_column_owner"]["enabled"] and lib.nk_norm_set_variant(-1) == 4
    reps[1] = {**L, "ok": True, "speedup": 1.0, "mask": 0}   # correct but neither kernel gained: stays off
    assert not tune.autotune()["layernorm_column_owner"]["enabled"] and lib.nk_norm_set_variant(-1) == 0
    reps[1] = {**L, "ok": False, "speedup": 1.6, "mask": 5}   # faster but wrong: stays off
    assert not tune.autotune()["layernorm_column_owner"]["enabled"] and lib.nk_norm_set_variant(-1) == 0
    reps[1] = {**L, "ok": True, "speedup": 1.4, "mask": 5}
    # GroupNorm reverse order: its bit is OR-ed into the norm mask
    reps.append({"variant": "groupnorm_reverse_apply", "ok": True, "speedup": 1.04, "checks": [], "timings": []})
    got = tune.autotune()
    assert got["groupnorm_reverse_apply"]["enabled"] and lib.nk_norm_set_variant(-1) == 7
    reps.pop()
    got = tune.autotune()
    assert not got["groupnorm_reverse_apply"]["enabled"] and lib.nk_norm_set_variant(-1) == 5
    # third variant: the epilogue prefetch hint
    assert not got["epilogue_l2_prefetch"]["enabled"] and "no verdict" in got["epilogue_l2_prefetch"]["error"]
    reps.insert(2, {"variant": "epilogue_l2_prefetch", "ok": True, "speedup": 1.06, "mask": 1, "checks": [], "timings": []})
    got = tune.autotune()
    assert got["epilogue_l2_prefetch"]["enabled"] and lib.nk_gemm_set_epi_prefetch(-1) == 1
    reps[2]["mask"] = 3
    assert tune.autotune()["epilogue_l2_prefetch"]["enabled"] and lib.nk_gemm_set_epi_prefetch(-1) == 3
    reps[2]["mask"] = 0   # neither side input gained: nothing to enable
    assert not tune.autotune()["epilogue_l2_prefetch"]["enabled"] and lib.nk_gemm_set_epi_prefetch(-1) == 0
    reps[2]["mask"] = 1
    json.dumps(tune._summary(got))
    reps[2]["speedup"] = 1.0
    assert not tune.autotune()["epilogue_l2_prefetch"]["enabled"] and lib.nk_gemm_set_epi_prefetch(-1) == 0
    reps[3]["ok"] = False  # the skewed order failed its equality checks: the plain order is used
    got = tune.autotune()
    assert got["enabled"] and got["skew"] == 0 and lib.nk_gemm_set_dual_skew(-1) == 0 and len(got["candidates"]) == 2
    lib.nk_gemm_set_dual(0)
    lib.nk_gemm_set_dual_min_k(0)
    lib.nk_gemm_set_dual_skew(0)
    lib.nk_norm_set_variant(0)
    lib.nk_gemm_set_epi_prefetch(0)


def test_probe_shapes_cover_every_paired_mode():
    kinds = {k for k, _ in tune.CHECK_SHAPES} | {k for k, _, _ in tune.TIMED_SHAPES}
    assert {"linear_fwd", "linear_fwd_f32", "linear_dgrad", "linear_wgrad", "linear_wgrad_acc", "conv", "conv_s2"} <= kinds
    # odd pair-tile counts (the half-empty last pair) are present for a matrix and for an image operand
    assert any(k == "linear_fwd" and ((d[0] + 127) // 128 + 1) // 2 % 2 == 1 for k, d in tune.CHECK_SHAPES)
    assert any(k == "conv" and d[0] * d[1] * d[2] < 128 * 2 for k, d in tune.CHECK_SHAPES)


def test_class_mask_selection():
    """a class of launches that loses at every depth is dropped without taking the others with it."""
    def row(kind, k, off, on):
        return {"kind": kind, "k_iters": k, "ms_unpaired": off, "ms_mode1_no_limit": on}
    rows = [row("linear_fwd", 10, 1.0, 1.1), row("linear_fwd", 20, 1.0, 0.9), row("linear_dgrad", 160, 1.0, 0.85),
            row("linear_wgrad", 256, 1.0, 1.2), row("conv", 45, 1.0, 0.95), row("conv", 180, 1.0, 0.8)]
    assert tune.pick_classes_and_min_k(rows) == (1 | 4, 45)
    rows[3]["ms_mode1_no_limit"] = 0.9
    assert tune.pick_classes_and_min_k(rows) == (7, 45)   # the weight-gradient class does not raise the depth limit
    rows.append(row("linear_wgrad", 1024, 1.0, 0.8))
    rows[3]["ms_mode1_no_limit"] = 1.3                    # loses at one of its depths: the class goes, the others stay
    assert tune.pick_classes_and_min_k(rows) == (1 | 4, 45)
    assert tune.pick_classes_and_min_k([row("linear_fwd", 20, 1.0, 1.3), row("conv", 45, 1.0, 1.2)]) == (0, None)
    prev = lib.nk_gemm_set_dual_classes(5)
    assert lib.nk_gemm_set_dual_classes(-1) == 5 and lib.nk_gemm_set_dual_classes(prev) == 5


def test_depth_threshold_selection():
    """pick_min_k: the smallest reduction depth from which mode 1 never loses (2 % timing noise allowed)."""
    def row(k, off, on):
        return {"k_iters": k, "ms_unpaired": off, "ms_mode1_no_limit": on}
    assert tune.pick_min_k([row(10, 1.0, 1.2), row(20, 1.0, 0.9), row(80, 1.0, 0.8), row(256, 1.0, 1.01)]) == 20
    assert tune.pick_min_k([row(10, 1.0, 0.9), row(20, 1.0, 0.9)]) == 10
    assert tune.pick_min_k([row(10, 1.0, 0.9), row(256, 1.0, 1.3)]) is None  # loses at the deepest reductions: never pays
    assert tune._k_iters("linear_fwd", (16384, 1280, 5120)) == 80 and tune._k_iters("linear_dgrad", (16384, 10240, 1280)) == 160
    assert tune._k_iters("linear_wgrad", (16384, 1280, 1280)) == 256 and tune._k_iters("conv", (16, 32, 32, 1280, 1280, 3)) == 180
    assert lib.nk_gemm_set_dual_min_k(-1) >= 0
    prev = lib.nk_gemm_set_dual_min_k(20)
    assert lib.nk_gemm_set_dual_min_k(prev) == 20


def test_verdict_cache_is_keyed_and_reused(monkeypatch, tmp_path):
    """a complete probe verdict is remembered per (library build, GPU model): the next autotune() on the same machine
    applies it without starting a child; an incomplete probe (a candidate took the child down) is not remembered."""
    import subprocess
    import tempfile
    monkeypatch.setenv("NK_B200_TUNE_CACHE", "1")
    monkeypatch.setattr(tempfile, "gettempdir", lambda: str(tmp_path))
    monkeypatch.delenv("NK_GEMM_DUAL", raising=False)
    monkeypatch.delenv("NK_B200_TUNE", raising=False)
    V = {"variant": "gemm_row_tile_pairing", "checks": [], "timings": []}
    reps = [{**V, "ok": True, "skew": 0, "speedup": 1.07, "min_k_iters": 20}, {**V, "ok": True, "skew": 3, "speedup": 1.03, "min_k_iters": 20},
            {"variant": "layernorm_column_owner", "ok": True, "speedup": 1.5, "mask": 5, "checks": [], "timings": []}]

    class Proc:
        def __init__(self, rows, rc=0):
            self.rows, self.returncode, self.pid = rows, rc, 0

        def communicate(self, timeout=None):
            return "\n".join(json.dumps(r) for r in self.rows) + "\n", ""

    calls = []
    monkeypatch.setattr(subprocess, "Popen", lambda *a, **k: (calls.append(1), Proc(reps))[1])
    first = tune.autotune()
    assert first["enabled"] and first["skew"] == 0 and len(calls) == 1 and "child process" in first["source"]
    assert len(list(tmp_path.glob("nk_b200_tune_*.json"))) == 1
    lib.nk_gemm_set_dual(0)
    lib.nk_norm_set_variant(0)
    second = tune.autotune()
    assert len(calls) == 1 and "cached" in second["source"] and second["enabled"] and second["min_k_iters"] == 20
    assert lib.nk_gemm_set_dual(-1) == 1 and lib.nk_norm_set_variant(-1) == 5 and second["layernorm_column_owner"]["enabled"]
    assert tune.cache_path("probe") != tune.cache_path("guard:sdxl:16")
    # an incomplete probe is not stored
    for f in tmp_path.glob("nk_b200_tune_*.json"):
        f.unlink()
    monkeypatch.setattr(subprocess, "Popen", lambda *a, **k: (calls.append(1), Proc(reps[:1], rc=-6))[1])
    tune.autotune()
    assert not list(tmp_path.glob("nk_b200_tune_*.json"))
    lib.nk_gemm_set_dual(0)
    lib.nk_gemm_set_dual_min_k(0)
    lib.nk_gemm_set_dual_skew(0)
    lib.nk_norm_set_variant(0)
    lib.nk_gemm_set_epi_prefetch(0)
