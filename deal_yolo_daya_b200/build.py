"""Build libdyd.so (the CUDA hot path + its C ABI) for sm_100a, in-tree.

    python -m deal_yolo_daya_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  ``-fmad=false`` keeps every fp64 multiply and add
separately rounded (the reference's arithmetic is CPython float arithmetic: no FMA).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libdyd.so"
SYNTH_OUT = PKG / "libdyd_synth.so"          # synthetic-table generator (bench / tests), kept out of the product library
SOURCES = ["api.cu", "bbox_iou.cu", "bbox_tma.cu", "hash_dedup.cu", "labels.cu", "host_pipeline.cu", "ingest.cpp", "csv_read.cpp", "np_perm.cpp"]
SYNTH_SOURCES = ["api.cu", "synth.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (looked at $NVCC, /usr/local/cuda/bin/nvcc, PATH)")


def sources():
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def needs_build() -> bool:
    if not OUT.exists() or not SYNTH_OUT.exists():
        return True
    t = min(OUT.stat().st_mtime, SYNTH_OUT.stat().st_mtime)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) + list((PKG.parent / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "--shared",
           "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("DYD_NVCC_FLAGS", "").split()          # tuning experiments: -DDYD_NW=.. -DDYD_TILE_CAP_V=..
    jobs = [(OUT, sources()), (SYNTH_OUT, [CSRC / s for s in SYNTH_SOURCES])]
    procs = [(out, subprocess.Popen(cmd + ["-o", str(out)] + [str(s) for s in srcs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
             for out, srcs in jobs]
    for out, pr in procs:
        log, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(log)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed building {out.name}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
