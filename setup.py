"""Packaging for horizonml_b200.

``python setup.py build_ext --inplace`` (or ``pip install -e . --no-build-isolation``) compiles the CUDA extension
IN-TREE as ``horizonml_b200/_C.so`` with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` — the same
``horizonml_b200.ops._ext.build()`` that ``__graft_entry__.build()`` calls; no GPU is needed to build."""
import os
import sys

from setuptools import Command, find_packages, setup

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class build_ext(Command):
    description = "compile csrc/*.cu for sm_100a and link horizonml_b200/_C.so in-tree"
    user_options = [("inplace", "i", "ignored: the extension is always built in-tree"), ("force", "f", "rebuild")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace, self.force = True, False

    def finalize_options(self):
        pass

    def run(self):
        from horizonml_b200.ops import _ext
        print("built", _ext.build(force=bool(self.force), verbose=True))


setup(
    name="horizonml_b200",
    version="0.1.0",
    description="Blackwell-native (sm_100a) data / pipeline / tensor / hybrid-parallel ResNet training framework",
    packages=find_packages(include=["horizonml_b200", "horizonml_b200.*"]),
    package_data={"horizonml_b200": ["_C.so"]},
    py_modules=["data_parallel_train", "layer_model_parallel_train", "tensor_parallel_train", "hybrid_parallel_train",
                "main", "train", "analyze_results", "bench"],
    python_requires=">=3.10",
    install_requires=["torch", "numpy", "pandas"],
    cmdclass={"build_ext": build_ext},
    entry_points={"console_scripts": [
        "hz-data-parallel=horizonml_b200.cli:data_parallel_main",
        "hz-layer-parallel=horizonml_b200.cli:layer_parallel_main",
        "hz-tensor-parallel=horizonml_b200.cli:tensor_parallel_main",
        "hz-benchmark=horizonml_b200.bench_suite:main",
        "hz-analyze=horizonml_b200.bench_suite:analyze_main",
    ]},
)
