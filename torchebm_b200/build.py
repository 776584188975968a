"""Build libebm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m torchebm_b200.build [--force] [--verbose]

Objects go to torchebm_b200/csrc/_build/, the library to torchebm_b200/lib/libebm_b200.so.  Both are
git-ignored but travel to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libebm_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["ebm_core.cu", "ebm_ess.cu", "ebm_hmc.cu", "ebm_mlp.cu", "ebm_mlp_tc.cu", "ebm_mlp_tc2.cu", "ebm_mlp_wide.cu", "ebm_mlp_deep.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(INCLUDE, "ebm_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(LIBDIR, "build.stamp")
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest()


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B build of the same ABI with extra -D flags -> lib/libebm_b200_<name>.so; load it with EBM_B200_LIB=<path>
    (kernel tuning only: the default library is what ships and what the tests run)."""
    return build(force=True, verbose=verbose, extra=[f"-D{d}" for d in defines], tag=name)


def build(force: bool = False, verbose: bool = False, extra=(), tag: str = "") -> str:
    if not tag and not force and is_current():
        return LIB
    nvcc = _nvcc()
    obj_dir = os.path.join(OBJ, tag) if tag else OBJ
    lib_path = os.path.join(LIBDIR, f"libebm_b200_{tag}.so") if tag else LIB
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if not tag:
        with open(os.path.join(LIBDIR, "build.stamp"), "w") as f:
            f.write(_digest())
    return lib_path


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m torchebm_b200.build --variant NAME DEFINE[=VALUE] ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if not a.startswith("--")], verbose="--verbose" in sys.argv))
    else:
        path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
        print(path)
