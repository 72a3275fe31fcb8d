"""Build the in-tree CUDA libraries (nvcc, sm_100a only).  No torch dependency: plain CUDA runtime C-ABI libraries.

  libmudg_sm100.so        the product: csrc/*.cu                                  (include/mudg.h)
  libmudg_sm100_test.so   tests only:  the same objects + csrc/test/*.cu          (include/mudg_test.h)
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
CTEST = os.path.join(CSRC, "test")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmudg_sm100.so")
TEST_LIB = os.path.join(HERE, "libmudg_sm100_test.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(os.path.dirname(HERE), "include")]


def _sources():
    prod = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    test = sorted(os.path.join(CTEST, f) for f in os.listdir(CTEST) if f.endswith(".cu")) if os.path.isdir(CTEST) else []
    return prod, test


def _stamp(path: str) -> str:
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include")
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    deps += [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]
    for p in deps + [path]:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(path: str, verbose: bool) -> str:
    tag = "test_" if os.path.dirname(path) == CTEST else ""
    obj = os.path.join(OBJ, tag + os.path.basename(path)[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(path)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {path}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr.strip():
        print(r.stderr, file=sys.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj


def _link(lib: str, objs, force: bool) -> None:
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest:
        cmd = [NVCC, "-shared", "-o", lib] + list(objs) + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    prod, test = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(prod) + len(test))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), prod + test))
    pobjs, tobjs = objs[:len(prod)], objs[len(prod):]
    _link(LIB, pobjs, force)
    if tobjs:
        _link(TEST_LIB, pobjs + tobjs, force)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
