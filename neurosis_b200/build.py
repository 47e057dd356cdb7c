"""In-tree build of the sm_100a kernel library (libnk_b200.so) and the kernel self-test binary.

`python -m neurosis_b200.build [--force] [--ktest]` — nvcc cross-compiles without a GPU.
The .so is written next to the sources (neurosis_b200/csrc/libnk_b200.so) so it travels with the
repo snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
ROOT = CSRC.parent.parent
LIB = CSRC / "libnk_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math"]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers() -> list[Path]:
    return sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)])
    if jobs:
        if verbose:
            print(f"[nk build] compiling {len(jobs)} file(s)", file=sys.stderr)
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_run, jobs))
    if force or jobs or _stale(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"])
    return LIB


def build_ktest(force: bool = False) -> Path:
    lib = build(force=force)
    out = ROOT / "tools" / "ktest"
    src = ROOT / "tools" / "ktest.cu"
    if force or _stale(out, [src, lib]):
        _run([NVCC, *ARCH, "-O2", "-std=c++17", str(src), "-o", str(out), "-L" + str(CSRC), "-lnk_b200",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../neurosis_b200/csrc"])
    return out


def install_reference(src: Path = Path("/root/reference"), force: bool = False) -> Path | None:
    """The reference arm of bench.py (`--impl reference`, `cpu_baseline`) and tools/stock_torch_bench.py run the UNMODIFIED
    reference from baseline/_ref (git-ignored, shipped to the GPU box with the snapshot).  (Re)install it with the
    contract's offline pip recipe when the source tree is present (the authoring container; the GPU box only uses the
    prebuilt copy).  /root/reference is read-only and the build writes egg-info, so pip works on a copy under /tmp."""
    import shutil
    import tempfile
    dst = ROOT / "baseline" / "_ref"
    marker = dst / "neurosis" / "modules" / "diffusion" / "openaimodel.py"
    if marker.exists() and not force:
        return dst
    if not (src / "src" / "neurosis").exists():
        return None
    with tempfile.TemporaryDirectory() as tmp:
        copy = Path(tmp) / "refcopy"
        shutil.copytree(src, copy, symlinks=True, ignore=shutil.ignore_patterns(".git"))
        dst.parent.mkdir(exist_ok=True)
        _run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
              "/opt/wheelhouse", "--target", str(dst), "--upgrade", str(copy)])
    return dst if marker.exists() else None


if __name__ == "__main__":
    force = "--force" in sys.argv
    p = build(force=force, verbose=True)
    print(p)
    if "--ktest" in sys.argv:
        print(build_ktest(force=force))
