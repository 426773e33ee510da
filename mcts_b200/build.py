"""Builds libmcts_b200.so (the C-ABI rollout engine) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU container; the built .so travels to the GPU
box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_build"
LIB_PATH = OUT_DIR / "libmcts_b200.so"
OBJ_DIR = OUT_DIR / "obj"  # listed in .gpurunignore: only the linked library travels

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-ffp-contract=off",  # the tree driver's f64 selection formulas are compared bit for bit with the oracle's
    "--compress-mode=size",  # the fatbins shrink ~6x; the library travels to the GPU box with every snapshot
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libmcts_b200.so")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compiles every csrc/*.cu to an object IN PARALLEL (one translation unit per game's kernels, plus the host
    side), then links libmcts_b200.so.  Objects go to mcts_b200/_build/obj/."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()
    units = sorted(CSRC.glob("*.cu"))

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        extra = os.environ.get("MR_NVCC_EXTRA", "").split()  # experiment switches (-DNAME=1) for A/B library variants
        cmd = [nvcc, *NVCC_FLAGS, *extra, f"-I{ROOT / 'include'}", f"-I{CSRC}", "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, units))
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH), *[str(o) for o in objs], "-ldl"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
