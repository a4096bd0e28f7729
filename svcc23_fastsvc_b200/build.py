"""Build libfsvc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m svcc23_fastsvc_b200.build [--force] [--verbose]

One object per translation unit (compiled in parallel, rebuilt only when it or a header changed), then one link.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libfsvc.so")
SOURCES = ["fsvc_abi.cu", "tc_forward.cu", "train.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "fsvc.h"))
    return deps


def _newer(path, deps):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))


def build(force=False, verbose=False, timeline=False):
    """timeline=True builds libfsvc_tl.so with -DFSVC_TIMELINE (in-kernel event stamps, tools/timeline.py; select it
    with FSVC_LIB=.../libfsvc_tl.so)."""
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = _headers()
    out = OUT.replace("libfsvc.so", "libfsvc_tl.so") if timeline else OUT
    objdir = OBJ + "_tl" if timeline else OBJ
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in sources]
    if not force and os.path.exists(out) and not _newer(out, hdrs + [os.path.join(CSRC, s) for s in sources]):
        return out   # a shipped .so (the GPU box gets no build/ directory it could compare against)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    flags = NVCC_FLAGS + (["-DFSVC_TIMELINE"] if timeline else [])
    todo = []
    for s, o in zip(sources, objs):
        src = os.path.join(CSRC, s)
        if force or _newer(o, hdrs + [src]):
            todo.append([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", o, src])
    with ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        list(ex.map(lambda c: _run(c, verbose), todo))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, verbose)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timeline="--timeline" in sys.argv))
