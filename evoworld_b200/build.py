"""Build recipe for libevoworld_b200.so (sm_100a only) and the C oracle.

The library is built IN-TREE (evoworld_b200/_lib/) with plain nvcc so that it travels with the repo
snapshot to the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "evoworld_b200" / "csrc"
OUT_DIR = ROOT / "evoworld_b200" / "_lib"
LIB_PATH = OUT_DIR / "libevoworld_b200.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "_build" / "libreproj_oracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stamp(srcs: list[Path]) -> str:
    h = hashlib.sha256()
    for p in srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "evoworld_b200.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("EVW_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link them into one shared library."""
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    stamp_file = OUT_DIR / "build.stamp"
    stamp = _stamp(srcs)
    if not force and LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    extra = os.environ.get("EVW_NVCC_EXTRA", "").split()  # experiments only (e.g. -DEVW_POLY_EVERY=2)
    objs = []
    procs = []
    for src in srcs:
        obj = OUT_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            failed = True
    (OUT_DIR / "build.log").write_text("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see evoworld_b200/_lib/build.log")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    subprocess.run(cmd, check=True)
    stamp_file.write_text(stamp)
    return LIB_PATH


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/reproj_oracle.c (CPU restatement; test infrastructure only)."""
    src = ORACLE_DIR / "reproj_oracle.c"
    ORACLE_LIB.parent.mkdir(parents=True, exist_ok=True)
    if not force and ORACLE_LIB.exists() and ORACLE_LIB.stat().st_mtime >= src.stat().st_mtime:
        return ORACLE_LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-mfma", "-fopenmp",
           str(src), "-o", str(ORACLE_LIB), "-lm"]
    subprocess.run(cmd, check=True)
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
    print(build_oracle(force="--force" in sys.argv))
