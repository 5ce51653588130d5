"""Pipeline (one stage, 4 micro-batches) and tensor-parallel (world 1) engines on the shimmed native backend — run as a
script by tests/test_cpu_native_plumbing.py (own process: it creates a gloo process group of one rank).  Prints one JSON
line: per-backend step statistics, fallbacks, calls per binding."""
import json
import os
import socket
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(which: str, handoff: bool):
    import torch.distributed as dist
    import horizonml_b200.models.resnet as R
    import horizonml_b200.ops as ops
    import horizonml_b200.ops.native_backend as nb
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.ops import _ext
    from horizonml_b200.ops import functional as fn
    from horizonml_b200.trainers.common import Runtime
    from test_cpu_native_plumbing import BF16, ShimC, tb
    shim = ShimC(_ext.load(required=True))
    nb.C, nb._dev, nb._STRICT = shim, (lambda t: True), True
    state = {"native": False}
    dev = torch.device("cpu")
    fn._be = lambda t: nb if state["native"] else tb
    fn.step_begin = ops.step_begin = lambda device=None: nb.step_begin(dev) if state["native"] else None
    fn.step_end = ops.step_end = lambda: nb.step_end() if state["native"] else None
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    g = torch.Generator().manual_seed(3)
    xs = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, generator=g)
    ys = torch.randint(0, 10, (16,), generator=g)
    out = {"which": which, "handoff": handoff}
    for native in (False, True):
        state["native"] = native
        R._BN_BWD_IN_DGRAD = handoff and native
        be = "native" if native else "torch"
        kw = dict(world_size=1, batch_size=16, device="cpu", dtype="bf16", backend=be, quiet=True, cuda_graph=False)
        if which == "pp":
            from horizonml_b200.trainers.pp import PPEngine
            eng = PPEngine(TrainConfig(strategy="layer", microbatches=4, **kw), Runtime(0, 1, dev, BF16, be, "gloo"))
        else:
            from horizonml_b200.trainers.tp import TPEngine
            eng = TPEngine(TrainConfig(strategy="tensor", **kw), Runtime(0, 1, dev, BF16, be, "gloo"))
        for _ in range(2):
            eng.step(xs, ys)
        out[be] = [float(v) for v in eng.stats.buf.tolist()]
    out["fallbacks"] = dict(nb.FALLBACKS)
    out["calls"] = dict(shim.calls)
    print("SHIM_ENGINE " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2 and sys.argv[2] == "handoff")
