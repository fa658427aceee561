"""In-tree build of libdistdiff_sm100.so (sm_100a only) -- ``python -m distdiff_b200.build``.

One ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` compile per ``csrc/*.cu`` (in parallel),
then one link against the NCCL that ships with torch.  The ``.so`` lands next to this file so that it
travels with the repo snapshot to the GPU box; nothing is JIT-built at run time.
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
# DD_LIB_OUT / DD_BUILD_DIR: development only -- an experimental variant next to the product build (tools/build_variant.sh)
BUILD = os.environ.get("DD_BUILD_DIR") or os.path.join(HERE, "_build")
LIB = os.environ.get("DD_LIB_OUT") or os.path.join(HERE, "libdistdiff_sm100.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _extra_flags():
    """DD_EXTRA_NVCC_FLAGS: extra compile flags (e.g. -DDD_ALL_LANES_ARRIVE for the racecheck build of tools/gpu_sanitize.sh)."""
    return os.environ.get("DD_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def nccl_dirs():
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    for base in (spec.submodule_search_locations if spec else []):
        inc = os.path.join(base, "nccl", "include")
        lib = os.path.join(base, "nccl", "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
            return inc, lib
    raise RuntimeError("torch-bundled NCCL (nvidia/nccl) not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(os.path.dirname(HERE), "include", "distdiff_sm100.h")]
    stamp = os.path.join(BUILD, "stamp")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    inc, libdir = nccl_dirs()

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *_extra_flags(), "-I", inc, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        open(obj[:-2] + ".ptxas.log", "w").write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-L", libdir, "-l:libnccl.so.2", "-Xlinker", f"-rpath={libdir}",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
