"""Build / load the in-tree CUDA extension ``horizonml_b200/_C*.so``.

``build()`` compiles ``csrc/*.cu`` with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo``
(cross-compiles on a GPU-less box), ``csrc/bindings.cpp`` with the host compiler against the torch
headers, and links them into ``horizonml_b200/_C.so`` — in-tree, so the binary travels with the
repo snapshot to the GPU box and no JIT happens there.  ``load()`` imports that file; on a GPU box
a missing extension is a hard error (never a silent PyTorch fallback).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import subprocess
import sys
import sysconfig
from typing import List, Optional

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_CSRC = os.path.join(_ROOT, "csrc")
_PKG = os.path.join(_ROOT, "horizonml_b200")
_SO = os.path.join(_PKG, "_C.so")
_BUILD = os.path.join(_CSRC, "build")
_CU = ["elementwise.cu", "depthwise.cu", "comm.cu", "conv_gemm.cu", "tp_fused.cu"]
_CPP = ["bindings.cpp"]
_HDRS = ["common.cuh", "tc05.cuh", "igemm_common.cuh", "dw_core.cuh", "launchers.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

_mod = None


def _sources() -> List[str]:
    return [f for f in _CU + _CPP + _HDRS if os.path.exists(os.path.join(_CSRC, f))]


def source_hash() -> str:
    h = hashlib.sha256()
    for f in sorted(_sources()):
        with open(os.path.join(_CSRC, f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def _nvcc() -> Optional[str]:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def is_built() -> bool:
    stamp = os.path.join(_BUILD, "stamp")
    return os.path.exists(_SO) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link ``_C.so``. Returns the .so path."""
    if is_built() and not force:
        return _SO
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build the sm_100a extension")
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(_BUILD, exist_ok=True)
    cuda_home = os.path.dirname(os.path.dirname(nvcc))
    objs, procs = [], []
    for f in _CU:
        src = os.path.join(_CSRC, f)
        if not os.path.exists(src):
            continue
        obj = os.path.join(_BUILD, f + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-I", _CSRC, "-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include"), _CSRC]:
        inc += ["-I", p]
    cxx = os.environ.get("CXX", "g++")
    for f in _CPP:
        obj = os.path.join(_BUILD, f + ".o")
        objs.append(obj)
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
               f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + inc + \
              ["-c", os.path.join(_CSRC, f), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out.decode(errors="replace") + "\n")
        if p.returncode != 0:
            raise RuntimeError(f"compile failed: {' '.join(cmd[:3])} ...")
    libdirs = ce.library_paths() + [os.path.join(cuda_home, "lib64")]
    link = [cxx, "-shared", "-o", _SO] + objs
    for d in libdirs:
        link += ["-L", d, f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace"))
        raise RuntimeError("link failed")
    with open(os.path.join(_BUILD, "stamp"), "w") as fh:
        fh.write(source_hash())
    return _SO


def load(required: bool = True):
    """Import the prebuilt extension (building it only when nvcc is available and it is missing)."""
    global _mod
    if _mod is not None:
        return _mod
    import torch  # noqa: F401  (must be imported before the extension: symbol resolution)
    if not os.path.exists(_SO):
        if _nvcc() is not None and os.environ.get("HZ_NO_AUTOBUILD", "0") != "1":
            try:
                build()
            except Exception as e:  # noqa: BLE001
                if required:
                    raise
                return None
        elif required:
            raise RuntimeError(
                f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first. "
                "The native sm_100a kernels are mandatory on a GPU box (no silent PyTorch fallback).")
        else:
            return None
    spec = importlib.util.spec_from_file_location("horizonml_b200._C", _SO)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:
        if required:
            raise
        return None
    sys.modules["horizonml_b200._C"] = mod
    _mod = mod
    return mod
