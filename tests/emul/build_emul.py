"""TEST INFRASTRUCTURE ONLY: compile matchering_b200/csrc/*.cu for the HOST through the
CUDA-on-CPU emulator (tests/emul/cuda_emul.h) into tests/emul/_build/libmatchering_b200_emul.so.
Same sources, same C ABI; device pointers are host pointers."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC_DIR = os.path.join(ROOT, "matchering_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libmatchering_b200_emul.so")
SOURCES = ["api.cu", "analyze.cu", "design.cu", "convolve.cu", "correct.cu", "limiter.cu", "pipeline.cu", "hostio.cu", "resample.cu"]
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-DMGB_EMULATE", "-include", os.path.join(HERE, "cuda_emul.h"),
         "-Wno-unused-function", "-Wno-unknown-pragmas", "-fno-strict-aliasing"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    deps_common = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith(".cuh")]
    deps_common += [os.path.join(HERE, "cuda_emul.h"), os.path.join(ROOT, "include", "matchering_b200.h")]
    jobs, objs = [], []
    for src in SOURCES:
        s = os.path.join(SRC_DIR, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + deps_common):
            jobs.append(["g++", *FLAGS, "-x", "c++", "-c", s, "-o", o])
    s = os.path.join(HERE, "cuda_emul.cpp")
    o = os.path.join(OUT_DIR, "cuda_emul.o")
    objs.append(o)
    if force or _stale(o, [s, os.path.join(HERE, "cuda_emul.h")]):
        jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-DMGB_EMULATE", "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulator build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or _stale(LIB_PATH, objs):
        run(["g++", "-shared", "-o", LIB_PATH, *objs])
    return LIB_PATH


if __name__ == "__main__":
    print(build())
