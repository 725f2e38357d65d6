"""In-tree build of the CUDA engine library (`libletf_b200.so`) for sm_100a.

One translation unit per kernel instantiation (csrc/inst_*.cu), compiled in parallel with
nvcc and linked into a single shared library next to this file, so that the `.so` travels
with the source tree.  Nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libletf_b200.so")
INCLUDE = os.path.join(HERE, "..", "include", "letf_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the engine is CUDA-only and cannot be built without it")
    return p


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [INCLUDE]


def _compile(src: str, force: bool) -> tuple[str, bool]:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    deps = [os.path.join(CSRC, src)] + _headers()
    if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj, False
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    jobs = jobs or min(len(srcs), max(2, (os.cpu_count() or 4)))
    with ThreadPoolExecutor(jobs) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(c for _, c in results)
    manifest = os.path.join(OBJ, "link.manifest")
    want = "\n".join(sorted(os.path.basename(o) for o in objs))
    have = open(manifest).read() if os.path.exists(manifest) else ""
    if rebuilt or force or want != have or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(manifest, "w") as f:
            f.write(want)
    if verbose:
        print(f"built {LIB} from {len(objs)} objects ({sum(c for _, c in results)} recompiled)", file=sys.stderr)
    return LIB


def resource_usage() -> dict:
    """registers / stack / spills per kernel from the ptxas logs of the last build."""
    import re
    out = {}
    for f in sorted(os.listdir(OBJ)) if os.path.isdir(OBJ) else []:
        if not f.endswith(".ptxas.log"):
            continue
        txt = open(os.path.join(OBJ, f)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
            out[f[:-10] + ":" + m.group(1)[:48]] = {"stack": int(m.group(2)), "spill_st": int(m.group(3)),
                                                    "spill_ld": int(m.group(4)), "regs": int(m.group(5))}
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    for k, v in resource_usage().items():
        print(k, v)
