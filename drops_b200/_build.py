"""In-tree build of libdrops_b200.so (CUDA kernels + C ABI) for sm_100a.

`python -m drops_b200._build` or drops_b200._build.build().  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdrops_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-fvisibility=hidden,-O3",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdrops_b200 cannot be built (there is no CPU fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(h) for h in hs)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return src, r.stderr


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libdrops_b200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    hm = _headers_mtime()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or \
            os.path.getmtime(obj) < max(os.path.getmtime(os.path.join(CSRC, src)), hm)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs):
                if verbose and log:
                    print(f"--- {src}\n{log}", file=sys.stderr)
    relink = bool(jobs) or not os.path.exists(LIB_PATH) or \
        any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    if relink:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs,
               "-cudart", "static", "-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
