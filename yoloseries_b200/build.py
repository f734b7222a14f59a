"""Build libysb_postproc.so in-tree (yoloseries_b200/_lib/) with nvcc for sm_100a.

Plain nvcc, no torch headers: the library is a C-ABI (include/ysb_postproc.h) that any host language can bind.
-fmad=false and no fast-math are REQUIRED for parity (see csrc/ysb_internal.cuh).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(OUT_DIR, "libysb_postproc.so")
SOURCES = ["c_abi.cu", "filter_kernels.cu", "nms_kernel.cu", "decode_kernels.cu", "iou_kernels.cu", "gather_kernels.cu", "eval_kernels.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ysb_postproc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


VARIANTS_LIB = os.path.join(OUT_DIR, "libysb_postproc_variants.so")


def variants_fresh():
    """True when the profiling build exists and is at least as new as every source it was built from."""
    return not _stale(VARIANTS_LIB)


def build_variants(verbose=False):
    """The profiling build: same ABI plus the filter-kernel variants selected by YSB_FILTER_VARIANT / YSB_BULK_PPT
    (csrc/filter_variants.cuh).  Load it with YSB_LIBRARY=<path>; never the default library."""
    return build(force=True, verbose=verbose, extra_flags=["-DYSB_PROFILING_VARIANTS"], lib=VARIANTS_LIB, tag=".var")


def build(force=False, verbose=False, extra_flags=(), lib=LIB, tag=""):
    if not force and not _stale():
        return lib
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", tag + ".o"))
        cmd = [NVCC, *FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = lib + f".tmp{os.getpid()}"   # link next to the target, then rename: other processes never see a partial file
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib)
    return lib


ADAPTER_SRC = os.path.join(CSRC, "torch_adapter.cpp")
ADAPTER_LIB = os.path.join(OUT_DIR, "libysb_torch.so")


def build_torch_adapter(force=False):
    """The thin PyTorch C++ extension over the C ABI (csrc/torch_adapter.cpp -> _lib/libysb_torch.so): host C++ only,
    compiled with g++ against torch's headers, in-tree.  It does NOT link libysb_postproc.so: its ysb_* symbols resolve
    against whichever build of the C-ABI library yoloseries_b200._lib loaded (RTLD_GLOBAL) before it -- the product
    library, or a profiling build selected with YSB_LIBRARY."""
    deps = [ADAPTER_SRC, os.path.join(ROOT, "include", "ysb_postproc.h")]
    if not force and os.path.exists(ADAPTER_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(ADAPTER_LIB) for d in deps):
        return ADAPTER_LIB
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    tmp = ADAPTER_LIB + f".tmp{os.getpid()}"
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-I", os.path.join(ROOT, "include"),
           *[a for p in ce.include_paths() for a in ("-isystem", p)], "-isystem", os.path.join(cuda_home, "include"),
           ADAPTER_SRC, "-o", tmp,
           *[a for p in ce.library_paths() for a in ("-L" + p, "-Wl,-rpath," + p)],
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for torch_adapter.cpp:\n{r.stdout}\n{r.stderr[-4000:]}")
    os.replace(tmp, ADAPTER_LIB)
    return ADAPTER_LIB


def build_all(force=False):
    lib = build(force=force)
    build_torch_adapter(force=force)
    return lib


if __name__ == "__main__":
    if "--variants" in sys.argv:
        print(build_variants(verbose="-v" in sys.argv))
    else:
        print(build(force=True, verbose="-v" in sys.argv, extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
        print(build_torch_adapter(force=True))
