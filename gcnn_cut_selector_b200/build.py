"""Build libgcnn_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgcnn_b200.so")
SOURCES = ["api.cu", "csr_build.cu", "edge.cu", "node.cu", "node_tc.cu", "node_bwd.cu", "node_fwd.cu", "edge_block.cu", "records.cu", "select.cu", "dp.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), variant: str = "") -> str:
    """``defines`` / ``variant``: an instrumented or A/B build, e.g. ``build(defines=("GCNN_ACT_PIECES=2",), variant="act2")``
    writes ``build/act2/libgcnn_b200.so`` (load it with ``GCNN_LIB=...``); the default build is the product."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "gcnn_b200.h"))
    objdir = os.path.join(HERE, "build", variant) if variant else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(objdir, "libgcnn_b200.so") if variant else LIB
    objs, procs = [], []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + [f"-D{d}" for d in defines]
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, *flags, "-c", path, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0 or verbose:
            print(f"--- nvcc {src} ---\n{out}", file=sys.stderr)
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, variant=var))
