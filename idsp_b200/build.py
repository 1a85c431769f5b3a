"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

    python -m idsp_b200.build [--force] [--verbose]

Output: idsp_b200/libidsp_b200.so  (git-ignored, travels to the GPU box with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libidsp_b200.so")
SOURCES = ["ctx.cu", "biquad.cu", "hbf.cu", "trig_lockin.cu", "cic.cu", "comm.cu", "coeff.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact float parity with the reference: no FMA contraction, no FTZ, IEEE div/sqrt
    "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
    "-cudart", "static",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "idsp_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def _deps(path: str, seen=None) -> set:
    """the source file plus, transitively, every `#include "..."` found next to it or under include/"""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for line in open(path, errors="replace"):
        line = line.strip()
        if line.startswith("#include \""):
            name = line.split('"')[1]
            for base in (os.path.dirname(path), CSRC, os.path.join(HERE, "..", "include")):
                cand = os.path.normpath(os.path.join(base, name))
                if os.path.exists(cand):
                    _deps(cand, seen)
                    break
    seen.add(os.path.abspath(__file__))
    return seen


def build(force: bool = False, verbose: bool = False, out: str = OUT) -> str:
    if out == OUT and not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    for var, macro in (("IDSP_HF_NT", "HF_NT"), ("IDSP_HFS_NT", "HFS_NT"), ("IDSP_HFS_R0", "HFS_R0"),
                       ("IDSP_HFS_NL", "HFS_NL"), ("IDSP_HFS_MINB", "HFS_MINB")):
        if os.environ.get(var):  # experiments on the tiled HBF kernels
            cmd += [f"-D{macro}={int(os.environ[var])}"]
    for d in filter(None, os.environ.get("IDSP_DEFS", "").split(",")):  # e.g. IDSP_DEFS=HFS_BAL=1,HFS_NT=64
        cmd += [f"-D{d}"]
    if os.environ.get("IDSP_TUNE"):  # tile-shape sweep builds (tools/sweep_biquad.py)
        cmd += ["-DIDSP_TUNE"]
    cmd += ["-ccbin", "g++"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    # one nvcc per translation unit, in parallel, then one link step
    objdir = os.path.join(HERE, "build", os.path.splitext(os.path.basename(out))[0])
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in cmd if f not in ("-shared", "-cudart", "static")]
    procs = []
    sig = " ".join(compile_flags)
    sigfile = os.path.join(objdir, "flags.txt")
    same_flags = os.path.exists(sigfile) and open(sigfile).read() == sig
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if same_flags and not force and os.path.exists(obj) and all(
                os.path.getmtime(d) <= os.path.getmtime(obj) for d in _deps(os.path.join(CSRC, src))):
            procs.append((src, obj, None))  # object is newer than the source and every header it includes
            continue
        procs.append((src, obj, subprocess.Popen(compile_flags + ["-c", "-o", obj, os.path.join(CSRC, src)], cwd=CSRC, env=env,
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = "", False
    for src, obj, pr in procs:
        if pr is None:
            continue
        o, _ = pr.communicate()
        log += o
        failed |= pr.returncode != 0
        if pr.returncode != 0 and os.path.exists(obj):
            os.remove(obj)
    if not failed:
        open(sigfile, "w").write(sig)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libidsp_b200.so")
    res = subprocess.run([nvcc, "-shared", "-cudart", "static", "-ccbin", "g++", "-o", out] + [o for _, o, _ in procs] + ["-ldl"],
                         cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libidsp_b200.so")
    if verbose:
        sys.stderr.write(log + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    o = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=o[0] if o else OUT))
