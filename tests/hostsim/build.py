"""TEST INFRASTRUCTURE: compiles the kernel sources with g++ against the SIMT emulator (simt.h)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "level-s2fm_official_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libls2fm_hostsim.so")


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "simt.h"),
                                                                os.path.join(ROOT, "include", "ls2fm.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-DLS_HOSTSIM", "-shared", "-fPIC",
           "-Wno-unused-value", "-o", OUT, os.path.join(CSRC, "ls2fm_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("hostsim build failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
