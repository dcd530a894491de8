"""Build libe4t_b200.so (sm_100a) in-tree with nvcc.  Called by __graft_entry__.build().

    python e4t-diffusion_b200/csrc/build.py [--force] [--verbose]

The shared library exposes only the extern "C" entry points declared in include/e4t_b200.h.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "e4t_b200", "libe4t_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def needs_build(force=False):
    if force or not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in os.listdir(HERE):
        if f.endswith((".cu", ".cuh", ".h")) and os.path.getmtime(os.path.join(HERE, f)) > t:
            return True
    inc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "e4t_b200.h")
    return os.path.exists(inc) and os.path.getmtime(inc) > t


def build(force=False, verbose=False):
    if not needs_build(force):
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    srcs = sources()

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(HERE, f))
                                                for f in os.listdir(HERE) if f.endswith((".cuh", ".h")) or f == src)):
            return obj, ""
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for (o, log), s in zip(results, srcs):
            print(f"== {s}\n{log}")
    link = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
            "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
