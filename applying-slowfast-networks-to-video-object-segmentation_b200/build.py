"""Build libsfvos.so (all CUDA kernels + the C ABI of include/sfvos.h) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA runtime + driver-entry-point code, loaded via ctypes.
    python <pkg>/build.py            # incremental
    python <pkg>/build.py --force
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libsfvos.so")
SOURCES = ["runtime.cu", "conv_umma.cu", "conv_tstack_umma.cu", "conv_pair_umma.cu", "wgrad_umma.cu", "wgrad_stack_umma.cu", "wgrad_halo_umma.cu", "wgrad_c32_umma.cu", "wgrad_pair_umma.cu", "conv_simt.cu", "elementwise.cu", "roi_align.cu", "mask_tail.cu", "box_tail.cu", "paste_masks.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "sfvos.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + HEADERS):
        return obj, ""
    r = subprocess.run([NVCC] + FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
