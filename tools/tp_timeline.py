"""Where does a fused GEMM+all-reduce kernel spend its time?  In-kernel clock64 stamps (csrc/tp_fused.cu HZ_STAMP) of
every CTA, taken from the LAST of a train of back-to-back launches replayed as a CUDA graph (no host in the loop, the
ranks are in lock-step), for the pull and the NVSwitch-multicast variant.  Run under torchrun on >= 2 GPUs."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PHASES = ["prologue", "wait_upstream", "tma+mma", "write_partial+fence", "post_arrival", "wait_peers", "pull", "store+stats",
          "epoch"]


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tp_timeline.json"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from horizonml_b200 import ops
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.parallel.symm import SymmHeap
    from horizonml_b200.parallel.tp import FusedTP
    ops.set_backend("native")
    heap = SymmHeap(dev, 64 << 20)
    res = {"world": world, "heap": heap.describe(), "rows": []}
    mhz = torch.cuda.clock_rate() / 1e3 if hasattr(torch.cuda, "clock_rate") else 1965.0
    for label, kind, xs, cout, stride in [("layer3.conv2.fwd", 0, (64, 256, 2, 2), 256, 1), ("layer4.conv2.fwd", 0, (64, 512, 1, 1), 512, 1),
                                          ("layer3.0.conv1.dgrad", 1, (64, 128, 4, 4), 256, 2)]:
        n, cin, h, w = xs
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        g = torch.Generator().manual_seed(1)
        if kind == 0:
            k = cin // world
            a = cl((torch.randn(n, k, h, w, generator=g) * 0.5).to(dev).bfloat16())
            wr = cl((torch.randn(cout, k, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).bfloat16())
            sx, sw = (n, k, h, w), (cout, k, 3, 3)
        else:
            k = cout // world
            a = cl((torch.randn(n, k, ho, wo, generator=g) * 0.5).to(dev).bfloat16())
            wr = cl((torch.randn(k, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).bfloat16())
            sx, sw = xs, (k, cin, 3, 3)
        for variant in (["pull", "nvls"] if heap.nvls else ["pull"]):
            fz = FusedTP(dev, heap=heap)
            fz.nvls = variant == "nvls"
            op = fz.allreduce_conv(kind, sx, sw, stride, 1)
            dbg = torch.zeros(op.tiles * 16, dtype=torch.int64, device=dev)
            for _ in range(3):
                op(a, wr)
            torch.cuda.synchronize(); dist.barrier()
            nb.C.tp_set_debug(dbg)
            gr = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                with torch.cuda.graph(gr):
                    for _ in range(16):
                        y = op(a, wr)
            nb.C.tp_set_debug(None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); dist.barrier()
            for _ in range(3):
                gr.replay()
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            per_call = e0.elapsed_time(e1) / (5 * 16) * 1e3
            st = dbg.view(op.tiles, 16)[:, :10].double().cpu()
            d = (st[:, 1:] - st[:, :-1]) / 1965.0           # us at 1965 MHz
            row = {"case": label, "variant": variant, "tiles": op.tiles, "graph_us_per_call": per_call,
                   "phase_us_mean": {p: round(float(d[:, i].mean()), 2) for i, p in enumerate(PHASES)},
                   "phase_us_max": {p: round(float(d[:, i].max()), 2) for i, p in enumerate(PHASES)},
                   "kernel_us_mean": round(float(((st[:, 9] - st[:, 0]) / 1965.0).mean()), 2)}
            t = torch.tensor([per_call], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row["graph_us_per_call_max_over_ranks"] = t.item()
            res["rows"].append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
        # the dense conv of the same shard + NCCL all-reduce, also as a graph train
        def unfused():
            yy = nb.conv_fwd(a, wr, 1, 1, True)[0] if kind == 0 else nb.conv_dgrad(a, wr, xs, stride, 1)
            dist.all_reduce(yy.permute(0, 2, 3, 1))
        for _ in range(3):
            unfused()
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(gr):
                for _ in range(16):
                    unfused()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier()
        for _ in range(3):
            gr.replay()
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 80 * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        row = {"case": label, "variant": "conv+nccl (graph)", "graph_us_per_call_max_over_ranks": t.item()}
        res["rows"].append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del gr
    dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        json.dump(res, open(out_path, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
