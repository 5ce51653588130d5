"""Static check of every call into the CUDA extension: for each `C.<function>(...)`, `C.PeerComm(...)` and PeerComm method
call in the package, the tools, the GPU tests and bench.py, the number of arguments (and the keyword names) must fit the
signature pybind11 reports for that binding.  Catches the kind of error that otherwise only shows up on a GPU box — a
binding that grew an argument while one of its callers did not (the peer-memory / fused tensor-parallel entry points
cannot be reached by the CPU shim of tests/test_cpu_native_plumbing.py: they need CUDA IPC handles)."""
import ast
import glob
import os
import re

import pytest

from horizonml_b200.ops import _ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEER_METHODS = {"allreduce", "allreduce_adam", "blocks_for", "export_handles", "heap_bytes", "heap_ptr", "heap_tensor",
                "import_handles", "link_local", "set_block_cap", "set_multicast", "symm_bytes", "zero1_shard", "zero1_step"}


def _signatures(doc: str):
    """[(min_args, max_args, names)] for every overload in a pybind docstring (self excluded)"""
    out = []
    for line in (doc or "").splitlines():
        m = re.match(r"\s*(?:\d+\.\s*)?\w+\((.*)\)\s*->", line)
        if not m:
            continue
        params, depth, cur = [], 0, ""
        for ch in m.group(1):
            depth += ch in "[("
            depth -= ch in "])"
            if ch == "," and depth == 0:
                params.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            params.append(cur.strip())
        params = [p for p in params if not p.startswith("self:")]
        names = [p.split(":")[0].strip() for p in params]
        out.append((sum(1 for p in params if "=" not in p.split(":", 1)[-1]), len(params), names))
    return out


def test_every_extension_call_site_fits_its_binding():
    C = _ext.load(required=False)
    if C is None:
        pytest.skip("extension not built")
    classes = {n for n in dir(C) if isinstance(getattr(C, n), type)}
    funcs = {n: _signatures(getattr(C, n).__doc__) for n in dir(C)
             if callable(getattr(C, n)) and not n.startswith("_") and n not in classes}
    methods = {n: _signatures(getattr(C.PeerComm, n).__doc__) for n in dir(C.PeerComm) if not n.startswith("_")}
    assert PEER_METHODS <= set(methods), PEER_METHODS - set(methods)
    files = (glob.glob(os.path.join(ROOT, "horizonml_b200", "**", "*.py"), recursive=True) +
             glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "test_gpu*.py")) +
             [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"), os.path.join(ROOT, "tests", "test_multigpu.py")])
    problems, checked = [], 0
    for f in sorted(files):
        for node in ast.walk(ast.parse(open(f).read())):
            if not isinstance(node, ast.Call) or not isinstance(node.func, ast.Attribute):
                continue
            name, base = node.func.attr, ast.unparse(node.func.value)
            is_c = base in ("C", "_C") or base.endswith(".C")
            if is_c and name in classes:
                sigs = _signatures(getattr(C, name).__init__.__doc__)
            elif is_c:
                if name not in funcs and not name.startswith("_"):
                    problems.append(f"{os.path.relpath(f, ROOT)}:{node.lineno}: C.{name} does not exist")
                sigs = funcs.get(name)
            else:
                sigs = methods.get(name) if name in PEER_METHODS else None
            if not sigs or any(isinstance(a, ast.Starred) for a in node.args):
                continue
            checked += 1
            npos, kws = len(node.args), [k.arg for k in node.keywords]
            if not any(lo <= npos + len(kws) <= hi and all(k in names for k in kws) for lo, hi, names in sigs):
                problems.append(f"{os.path.relpath(f, ROOT)}:{node.lineno}: {base}.{name} called with {npos} positional + {kws}; "
                                f"the binding takes {[(lo, hi) for lo, hi, _ in sigs]}")
    assert not problems, "\n".join(problems)
    assert checked > 100, checked
