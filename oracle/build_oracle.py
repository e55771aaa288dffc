"""Compile oracle/dyd_oracle.c -> oracle/libdyd_oracle.so (test infrastructure).

Tries OpenMP first and falls back to a scalar build when libgomp is missing.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "dyd_oracle.c"
OUT = HERE / "libdyd_oracle.so"


def build(force: bool = False) -> Path:
    if OUT.exists() and not force and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    base = [cc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(OUT), str(SRC)]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base + extra, capture_output=True, text=True)
        if r.returncode == 0:
            return OUT
    raise RuntimeError(f"could not compile the C oracle:\n{r.stderr}")


if __name__ == "__main__":
    print(build(force=True))
