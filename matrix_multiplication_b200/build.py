"""In-tree build of the native pieces (no JIT cache: the built .so files travel with the tree).

  libspmm_b200.so     CUDA kernels + C ABI (include/spmm_b200.h), nvcc, sm_100a only
  custom_mm*.so       pybind11 module with the reference's 16 entry points
                      (reference src/custom_mm.cpp:393-416), g++ against libtorch,
                      linked to libspmm_b200.so with an $ORIGIN rpath

Mirrors what the reference's src/setup.py:9-21 does with CUDAExtension, but with the
sm_100a flags spelled out (`-gencode arch=compute_100a,code=sm_100a -lineinfo`).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libspmm_b200.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
CU_SOURCES = ["capi.cu", "spmm_fwd.cu", "sddmm.cu", "transpose.cu", "dense_to_csr.cu", "tc_gemm.cu", "spmm_slab.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without CUDA")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, log: Path | None = None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(" ".join(map(str, cmd)) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"command failed ({r.returncode}): {' '.join(map(str, cmd))}")
    return r.stdout


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", ROOT / "include" / "spmm_b200.h"]
    nvcc = _nvcc()

    def one(src: str):
        s = CSRC / src
        o = OBJ / (src + ".o")
        if force or _newer(o, [s, *headers]):
            _run([nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)], log=OBJ / (src + ".log"))
            if verbose:
                print(f"[build] nvcc {src}")
        return o

    with ThreadPoolExecutor(max_workers=len(CU_SOURCES)) as ex:
        objs = list(ex.map(one, CU_SOURCES))
    if force or _newer(LIB, objs):
        _run([nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode",
              "arch=compute_100a,code=sm_100a"])
        if verbose:
            print(f"[build] linked {LIB.name}")
    return LIB


def ext_path() -> Path:
    return PKG / ("custom_mm" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ext(force: bool = False, verbose: bool = False) -> Path:
    """pybind11 module `custom_mm` (host C++ only; device code lives in libspmm_b200.so)."""
    import torch
    from torch.utils import cpp_extension as ce

    build_lib(force=force, verbose=verbose)
    src = CSRC / "custom_mm.cpp"
    out = ext_path()
    if not (force or _newer(out, [src, ROOT / "include" / "spmm_b200.h", LIB])):
        return out
    inc = []
    for p in ce.include_paths("cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(True):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"], "-I", "/usr/local/cuda/include", "-I", str(ROOT / "include")]
    torch_lib = Path(torch.__file__).parent / "lib"
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cuda_lib = "/usr/local/cuda/lib64"
    cmd = [
        os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-attributes",
        f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", "-DTORCH_EXTENSION_NAME=custom_mm",
        "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, str(src), "-o", str(out),
        "-L", str(torch_lib), "-L", str(PKG), "-L", cuda_lib,
        "-lspmm_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
        "-lcublas", "-lcudart",
        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}", "-Wl,--no-as-needed",
    ]
    _run(cmd, log=OBJ / "custom_mm.log")
    if verbose:
        print(f"[build] g++ {out.name}")
    return out


def build_all(force: bool = False, verbose: bool = False):
    build_lib(force=force, verbose=verbose)
    build_ext(force=force, verbose=verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
