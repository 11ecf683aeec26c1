"""Builds libmedmamba_b200.so in-tree with nvcc for sm_100a (no torch headers: seconds to build).

The shared object lands next to this file; it is git-ignored (``*.so``) but travels to the GPU
box with the repo snapshot.  ``python -m medmamba_b200.build [--force] [--verbose]``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libmedmamba_b200.so")
STAMP = LIB_PATH + ".stamp"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(INCLUDE, "medmamba_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())      # location-independent: the GPU box mounts the repo elsewhere
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def up_to_date() -> bool:
    return (os.path.exists(LIB_PATH) and os.path.exists(STAMP)
            and open(STAMP).read().strip() == _digest())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB_PATH
    # one builder at a time (torchrun starts one process per GPU); the others wait and re-check
    import fcntl
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    with open(os.path.join(HERE, "_build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    objs = []
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    procs = []
    digest = _digest()
    for src in sources():
        obj = os.path.join(HERE, "_build", os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-I", INCLUDE, "-c", src, "-o", obj]
        if os.path.basename(src) == "api.cu":
            # the digest of the sources this library was built from, queryable as mmb_source_digest():
            # the loader refuses a library whose digest differs from the sources next to it
            cmd.insert(1, f'-DMMB_SOURCE_DIGEST="{digest}"')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    tmp = LIB_PATH + ".tmp"
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs + ["-lcudart"])
    os.replace(tmp, LIB_PATH)                        # never expose a half-written library
    with open(os.path.join(HERE, "_build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
