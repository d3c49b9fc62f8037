"""`python setup.py build_ext --inplace` / `pip install -e .` — the extension build of the drop-in.

Mirrors the reference's two setup scripts (src/setup.py:9-21 builds the CUDAExtension `custom_mm`;
setup.py:3-15 installs `matmuls` as a module), now targeting sm_100a only:

  libspmm_b200.so   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo   (kernels + C ABI)
  custom_mm*.so     g++ against libtorch, linked to libspmm_b200.so with an $ORIGIN rpath    (pybind11, 16 names)

The native pieces are built in-tree by matrix_multiplication_b200/build.py (no JIT cache), so the same command
works from a source checkout and from `__graft_entry__.build()`.
"""
from setuptools import Command, setup
from setuptools.command.build_ext import build_ext as _build_ext
from setuptools.command.build_py import build_py as _build_py


def _native(force=False):
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location(
        "b200_build", pathlib.Path(__file__).parent / "matrix_multiplication_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_all(force=force, verbose=True)


class build_ext(_build_ext):
    def run(self):
        _native(force=getattr(self, "force", False))


class build_py(_build_py):
    def run(self):
        _native()
        super().run()


setup(
    name="matrix_multiplication_b200",
    version="0.1.0",
    description="B200-native fp32 CSR SpMM engine behind the custom_mm / matmuls surface of smoorjani/matrix-multiplication",
    packages=["matrix_multiplication_b200"],
    py_modules=["matmuls", "custom_mm"],
    package_data={"matrix_multiplication_b200": ["*.so", "csrc/*"]},
    cmdclass={"build_ext": build_ext, "build_py": build_py},
    python_requires=">=3.10",
)
