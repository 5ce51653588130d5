"""Pipeline and tensor-parallel engines on the shimmed native backend — run as a script by
tests/test_cpu_native_plumbing.py (own processes: gloo process group of WORLD ranks, one spawned process per rank).
    python tests/helpers_shim_engines.py dp|dpz|pp|tp WORLD [handoff]
Prints one JSON line per rank: per-backend step statistics, fallbacks, calls per binding."""
import json
import os
import socket
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def worker(rank: int, world: int, port: int, which: str, handoff: bool, q):
    import torch.distributed as dist
    import horizonml_b200.models.resnet as R
    import horizonml_b200.ops as ops
    import horizonml_b200.ops.native_backend as nb
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.ops import _ext
    from horizonml_b200.ops import functional as fn
    from horizonml_b200.trainers.common import Runtime
    from test_cpu_native_plumbing import BF16, ShimC, tb
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // (2 * world)))
    shim = ShimC(_ext.load(required=True))
    nb.C, nb._dev, nb._STRICT = shim, (lambda t: True), True
    state = {"native": False}
    dev = torch.device("cpu")
    fn._be = lambda t: nb if state["native"] else tb
    fn.step_begin = ops.step_begin = lambda device=None: nb.step_begin(dev) if state["native"] else None
    fn.step_end = ops.step_end = lambda: nb.step_end() if state["native"] else None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    xs = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, generator=g)
    ys = torch.randint(0, 10, (16,), generator=g)
    out = {"which": which, "world": world, "rank": rank, "handoff": handoff}
    for native in (False, True):
        state["native"] = native
        R._BN_BWD_IN_DGRAD = handoff and native
        be = "native" if native else "torch"
        kw = dict(world_size=world, batch_size=16, device="cpu", dtype="bf16", backend=be, quiet=True, cuda_graph=False)
        if which in ("dp", "dpz"):
            from horizonml_b200.trainers.dp import DPEngine
            xs_r, ys_r = xs[rank::world], ys[rank::world]                # this rank's shard of the batch
            kw["batch_size"] = xs_r.shape[0]
            cfg = TrainConfig(strategy="data", zero1=(which == "dpz"), zero1_impl="fused", **kw)
            eng = DPEngine(cfg, Runtime(rank, world, dev, BF16, be, "gloo"))
            for _ in range(2):
                eng.step(xs_r, ys_r)
        elif which == "pp":
            from horizonml_b200.trainers.pp import PPEngine
            eng = PPEngine(TrainConfig(strategy="layer", microbatches=4, **kw), Runtime(rank, world, dev, BF16, be, "gloo"))
        else:
            from horizonml_b200.trainers.tp import TPEngine
            eng = TPEngine(TrainConfig(strategy="tensor", **kw), Runtime(rank, world, dev, BF16, be, "gloo"))
        for _ in range(2 if which in ("pp", "tp") else 0):
            eng.step(xs, ys)
        out[be] = [float(v) for v in eng.stats.buf.tolist()]
        dist.barrier()
    out["fallbacks"] = dict(nb.FALLBACKS)
    out["calls"] = dict(shim.calls)
    q.put(out)
    dist.destroy_process_group()


def main(which: str, world: int, handoff: bool):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, which, handoff, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    res, t0 = [], time.time()
    while len(res) < world and time.time() - t0 < 540:
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):          # a rank died: do not wait for its result
                break
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.terminate()
    for o in sorted(res, key=lambda o: o["rank"]):
        print("SHIM_ENGINE " + json.dumps(o), flush=True)
    return 0 if (len(res) == world and all(p.exitcode == 0 for p in procs)) else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "handoff"))
