"""Build libb200ldm.so (in-tree) from csrc/*.cu with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot; nothing is JIT-built at
run time.  `python -m audioldm_with_lora_b200.build` rebuilds unconditionally.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb200ldm.so"
SOURCES = ["conv_gemm.cu", "attention.cu", "attention_bwd.cu", "norm.cu", "sampler.cu", "train.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh"))
    return any(p.stat().st_mtime > t for p in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        sp = CSRC / src
        if not sp.exists():
            continue
        obj = build_dir / (src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(sp), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"[b200 build] {src} FAILED\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[b200 build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed building libb200ldm.so")
    link = [_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-lcudart", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force=True, verbose="-v" in sys.argv)
    print(p)
