"""Builds the REFERENCE's own extension, unmodified, into oracle/_ref/ (git-ignored).

TEST / BENCH INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); the GPU
box only uses the prebuilt files that travel with the tree.  Nothing is copied into the repo's
history: the sources are compiled where they lie and only binaries land in oracle/_ref/:

  oracle/_ref/custom_mm.so   the reference's pybind module (src/custom_mm.cpp, baseline_mm.cu,
                             sparse_mm.cu, naive_sparse_mm.cu — the file list of src/setup.py:12-17)
  oracle/_ref/matmuls_ref.bin    byte-compiled /root/reference/matmuls.py (its autograd wrappers)

The single deviation from `python setup.py install` in src/: CUDA 12 removed the enum
CUSPARSE_MM_ALG_DEFAULT used at src/baseline_mm.cu:308,313, so the build passes
-DCUSPARSE_MM_ALG_DEFAULT=CUSPARSE_SPMM_ALG_DEFAULT (the documented replacement).  The hot
path under comparison (cusparse_mm_wrapper, src/baseline_mm.cu:167-216) is not touched by it.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import py_compile
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("REFERENCE_ROOT", "/root/reference"))
SOURCES = ["custom_mm.cpp", "baseline_mm.cu", "sparse_mm.cu", "naive_sparse_mm.cu"]


def available() -> bool:
    return (OUT / "custom_mm.so").exists() and (OUT / "matmuls_ref.bin").exists()


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds a usable build (fresh or existing)."""
    if available() and not force:
        return True
    if not (REF / "src" / "custom_mm.cpp").exists():
        return False
    import torch
    from torch.utils import cpp_extension as ce

    OUT.mkdir(exist_ok=True)
    obj = OUT / "obj"
    obj.mkdir(exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"], "-I", str(REF / "src")]
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-O2", "-std=c++17", f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", "-DTORCH_EXTENSION_NAME=custom_mm",
              "-DTORCH_API_INCLUDE_EXTENSION_H", "-DCUSPARSE_MM_ALG_DEFAULT=CUSPARSE_SPMM_ALG_DEFAULT", *inc]

    def one(src: str) -> Path:
        o = obj / (src + ".o")
        s = REF / "src" / src
        if src.endswith(".cu"):
            cmd = ["/usr/local/cuda/bin/nvcc", *common, "-gencode", "arch=compute_100a,code=sm_100a",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-w", "-c", str(s), "-o", str(o)]
        else:
            cmd = ["g++", *common, "-fPIC", "-w", "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        (obj / (src + ".log")).write_text(" ".join(cmd) + "\n" + r.stdout)
        if r.returncode != 0:
            raise RuntimeError(f"reference build failed on {src}:\n{r.stdout[-3000:]}")
        if verbose:
            print(f"[build_ref] compiled {src}", flush=True)
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, SOURCES))
    torch_lib = Path(torch.__file__).parent / "lib"
    link = ["g++", "-shared", "-o", str(OUT / "custom_mm.so"), *map(str, objs), "-L", str(torch_lib),
            "-L", "/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-lcusparse", "-lcublas", "-lcudart", f"-Wl,-rpath,{torch_lib}"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference link failed:\n{r.stdout[-3000:]}")
    py_compile.compile(str(REF / "matmuls.py"), cfile=str(OUT / "matmuls_ref.bin"), doraise=True)
    if verbose:
        print(f"[build_ref] wrote {OUT}/custom_mm.so and matmuls_ref.bin", flush=True)
    return True


def load():
    """Imports the reference's (custom_mm, matmuls) from oracle/_ref.  GPU box only (the module
    initialises CUDA-dependent globals lazily, but every entry point needs a device)."""
    import importlib.machinery
    import importlib.util

    if not available():
        raise ImportError("oracle/_ref is not built (run `python oracle/build_ref.py` in the build container)")
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    saved = {k: sys.modules.get(k) for k in ("custom_mm", "matmuls")}
    try:
        spec = importlib.util.spec_from_file_location("custom_mm", OUT / "custom_mm.so")
        cm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cm)
        sys.modules["custom_mm"] = cm  # the reference's matmuls does `import custom_mm`
        loader = importlib.machinery.SourcelessFileLoader("matmuls", str(OUT / "matmuls_ref.bin"))
        spec2 = importlib.util.spec_from_loader("matmuls", loader)
        mm = importlib.util.module_from_spec(spec2)
        loader.exec_module(mm)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return cm, mm


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference build:", "ok" if ok else "unavailable (/root/reference missing)")
