"""Multi-GPU collective check + micro-benchmark (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py out.json

* fused cast/scale peer all-reduce (one-shot / two-shot / NVLS; bf16 and fp32 wire) vs NCCL all-reduce
  over sizes 64 B … 48 MiB incl. the three DDP-like bucket sizes of SURVEY §2.4 (K2)
* back-to-back soak (flag / parity reuse) and CUDA-graph replay of the collective
* device-timed latency, max over ranks, and achieved fraction of the NVLink roofline
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/multigpu.json"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from horizonml_b200.parallel.comm import PeerAllReduce
    res = {"world": world, "cases": [], "timing": []}
    max_n = 12 * (1 << 20)          # 48 MiB fp32
    ars = {}
    for wire in ("bf16", "fp32"):
        try:
            ars[wire] = PeerAllReduce(max_n, dev, algo="auto", wire=wire)
            res[f"nvls_{wire}"] = ars[wire].has_nvls
        except Exception as e:  # noqa: BLE001
            res[f"init_error_{wire}"] = repr(e)
    if rank == 0:
        print("init:", {k: v for k, v in res.items() if k != "cases"}, flush=True)

    def rmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sizes = [16, 1024, 65536, 2364416 // 8 * 8, 6622208, 2191360, 11181696, max_n]
    for wire, ar in ars.items():
        algos = ["oneshot", "twoshot"] + (["nvls"] if ar.has_nvls else [])
        for n in sizes:
            g = torch.Generator(device="cpu").manual_seed(100 + rank)
            x = torch.randn(n, generator=g).to(dev)
            ref = x.clone()
            if wire == "bf16":
                ref = (ref / world).bfloat16().float()
                dist.all_reduce(ref)
            else:
                dist.all_reduce(ref)
                ref /= world
            for algo in algos + (["ll"] if (wire == "bf16" and n <= 512 * 1024) else []):
                y = x.clone()
                try:
                    for _ in range(3):           # back-to-back: flags / parity buffers are reused
                        y.copy_(x)
                        ar.allreduce_avg_(y, algo)
                    torch.cuda.synchronize()
                    err = ((y - ref).abs().max() / (ref.abs().max() + 1e-9)).item()
                    gathered = [torch.empty_like(y) for _ in range(world)]
                    dist.all_gather(gathered, y)
                    same = all(torch.equal(gathered[0], t) for t in gathered)
                    tol = 2e-2 if wire == "bf16" else 1e-5
                    ok = err < tol and same and ar.handle.error() == 0
                    case = {"wire": wire, "algo": algo, "n": n, "rel_err": err, "identical": same, "ok": bool(ok)}
                except Exception as e:  # noqa: BLE001
                    case = {"wire": wire, "algo": algo, "n": n, "ok": False, "exc": repr(e)}
                res["cases"].append(case)
                if rank == 0:
                    print(("PASS " if case["ok"] else "FAIL ") + json.dumps(case), flush=True)

    # ---- timing: fused (fp32 grad -> bf16 wire) vs NCCL (bf16 tensor all-reduce + separate cast/scale passes)
    def time_fn(fn, iters=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return rmax(e0.elapsed_time(e1) / iters * 1e3)        # us, max over ranks

    if "bf16" in ars:
        ar = ars["bf16"]
        for n in (16384, 262144, 2191360, 6622208, 11181696):
            x = torch.randn(n, device=dev)
            xb = torch.empty(n, device=dev, dtype=torch.bfloat16)
            row = {"n": n, "wire_bytes": n * 2}
            for algo in ["oneshot", "twoshot"] + (["nvls"] if ar.has_nvls else []) + (["ll"] if n <= 512 * 1024 else []):
                if algo == "oneshot" and n * 2 * world > (256 << 20):
                    continue
                row[algo + "_us"] = time_fn(lambda a=algo: ar.allreduce_avg_(x, a))

            def nccl_path():
                torch.mul(x, 1.0 / world, out=x)          # scale
                xb.copy_(x)                               # cast
                dist.all_reduce(xb)
                x.copy_(xb)                               # back to the fp32 grad buffer
            row["nccl_cast_scale_us"] = time_fn(nccl_path)
            row["nccl_only_bf16_us"] = time_fn(lambda: dist.all_reduce(xb))
            # roofline: two-shot moves 2*(W-1)/W*S per GPU per direction; NVLS ~S each way (770 GB/s measured peer bw)
            s = n * 2
            row["roofline_twoshot_us"] = 2 * (world - 1) / world * s / 770e9 * 1e6
            row["roofline_nvls_us"] = s / 770e9 * 1e6
            best = min(v for k, v in row.items() if k.endswith("_us") and k.split("_")[0] in ("oneshot", "twoshot", "nvls"))
            row["best_fused_us"] = best
            row["frac_of_twoshot_roofline"] = row["roofline_twoshot_us"] / best
            res["timing"].append(row)
            if rank == 0:
                print("TIME " + json.dumps(row), flush=True)
        # CUDA-graph replay of the collective (device-side counters make it replayable)
        try:
            x = torch.randn(2191360, device=dev)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                ar.allreduce_avg_(x, "twoshot")
            torch.cuda.synchronize(); dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ar.allreduce_avg_(x, "twoshot")
            x0 = torch.randn(2191360, device=dev)
            for _ in range(5):
                x.copy_(x0)
                g.replay()
            torch.cuda.synchronize()
            ref = (x0 / world).bfloat16().float()
            dist.all_reduce(ref)
            res["graph_replay_ok"] = bool(((x - ref).abs().max() / ref.abs().max()).item() < 2e-2 and ar.handle.error() == 0)
        except Exception as e:  # noqa: BLE001
            res["graph_replay_ok"] = False
            res["graph_exc"] = repr(e)
    res["n_fail"] = sum(1 for c in res["cases"] if not c["ok"])
    dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        with open(out_path, "w") as fh:
            json.dump(res, fh, indent=1)
        print(f"multigpu: {len(res['cases'])} cases, {res['n_fail']} failed, graph={res.get('graph_replay_ok')}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
