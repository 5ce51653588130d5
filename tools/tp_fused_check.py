"""Fused tensor-parallel kernels vs (native conv + NCCL collective), run under torchrun on real GPUs:
GEMM→all-reduce (row-parallel conv2 forward, column-parallel conv1 dgrad incl. stride 2) with the NVSwitch-multicast
(``multimem``) and the peer-pull variant, GEMM→reduce-scatter, all-gather→GEMM, the one-kernel classifier head and the
stand-alone bf16 all-reduce: correctness against an fp32 reference, bit-identity across ranks, re-launch / CUDA-graph
replay, device-timed latency (max over ranks) and roofline fraction (the slower of the FLOPs at the measured bf16 peak
and the bytes this rank must receive over NVLink at 900 GB/s)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PEAK_FLOPS = 1433.5e12
NVLINK_BPS = 900e9


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tp_fused.json"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from horizonml_b200 import ops
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    from horizonml_b200.parallel.symm import SymmHeap
    from horizonml_b200.parallel.tp import FusedTP, padded_classes
    ops.set_backend("native")
    heap = SymmHeap(dev, 192 << 20)
    f_mc = FusedTP(dev, heap=heap)                      # multimem variant when the heap is multicast-mapped
    f_pull = FusedTP(dev, heap=heap)
    f_pull.nvls = False                                  # same heap, peer-pull variant
    variants = [("pull", f_pull)] + ([("nvls", f_mc)] if heap.nvls else [])
    res = {"world": world, "heap": heap.describe(), "cases": [], "timing": []}
    if rank == 0:
        print("heap " + json.dumps(res["heap"]), flush=True)

    def rmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def time_fn(fn, iters=40):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return rmax(e0.elapsed_time(e1) / iters * 1e3)

    def report(name, got, ref, tol=3e-2, extra=None, check_same=True):
        err = ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-6)).item()
        same = None
        if check_same:
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got.contiguous())
            same = all(torch.equal(gathered[0], t) for t in gathered)
        c = {"name": name, "rel_err": err, "identical_across_ranks": same, "ok": bool(err < tol)}
        c.update(extra or {})
        res["cases"].append(c)
        if rank == 0:
            print(("PASS " if c["ok"] else "FAIL ") + json.dumps(c), flush=True)

    # (kind, dense x_shape [N,Cin,H,W], dense Cout, stride, label): the reduction points of ResNet-18's layer3 / layer4
    # blocks at batch 64 (kind 0 = row-parallel conv2 forward: Cin split; kind 1 = column-parallel conv1 dgrad: Cout split)
    cases = [(0, (64, 256, 2, 2), 256, 1, "layer3.conv2.fwd"), (0, (64, 512, 1, 1), 512, 1, "layer4.conv2.fwd"),
             (1, (64, 256, 2, 2), 256, 1, "layer3.1.conv1.dgrad"), (1, (64, 128, 4, 4), 256, 2, "layer3.0.conv1.dgrad"),
             (1, (64, 256, 2, 2), 512, 2, "layer4.0.conv1.dgrad"), (1, (64, 512, 1, 1), 512, 1, "layer4.1.conv1.dgrad"),
             (0, (64, 512, 4, 4), 128, 1, "wide.conv.fwd")]
    for kind, xs, cout, stride, label in cases:
        n, cin, h, w = xs
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        g = torch.Generator().manual_seed(100)           # same full tensors on every rank, each takes its shard
        W_full = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        if kind == 0:
            k = cin // world
            full = torch.randn(n, cin, h, w, generator=g) * 0.5
            a = cl(full[:, rank * k:(rank + 1) * k].to(dev).bfloat16())
            wr = cl(W_full[:, rank * k:(rank + 1) * k].to(dev).bfloat16())
            ref, _ = tb.conv_fwd(a.float(), wr.float(), 1, 1, False)
            shard_x, shard_w, out_shape = (n, k, h, w), (cout, k, 3, 3), (n, cout, h, w)
            dense = lambda: nb.conv_fwd(a, wr, 1, 1, True)[0]                       # noqa: E731
        else:
            k = cout // world
            full = torch.randn(n, cout, ho, wo, generator=g) * 0.5
            a = cl(full[:, rank * k:(rank + 1) * k].to(dev).bfloat16())
            wr = cl(W_full[rank * k:(rank + 1) * k].to(dev).bfloat16())
            ref = tb.conv_dgrad(a.float(), wr.float(), xs, stride, 1)
            shard_x, shard_w, out_shape = xs, (k, cin, 3, 3), xs
            dense = lambda: nb.conv_dgrad(a, wr, xs, stride, 1)                     # noqa: E731
        ref = ref.contiguous()
        dist.all_reduce(ref)
        if not f_pull.supported(kind, shard_x, shard_w[0], stride):
            continue
        row = {"case": label, "kind": kind, "x": list(xs), "cout": cout, "stride": stride, "shard_k": k}
        for vname, fz in variants:
            op = fz.allreduce_conv(kind, shard_x, shard_w, stride, 1)
            st = torch.zeros(2, out_shape[1], device=dev) if kind == 0 else None
            for _ in range(3):
                if st is not None:
                    st.zero_()
                y = op(a, wr, None, st)
            torch.cuda.synchronize()
            report(f"{label}.allreduce.{vname}", y, ref)
            if st is not None:
                yf = y.float()
                sref = torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])
                report(f"{label}.bn_sums.{vname}", st, sref, tol=1e-3, check_same=False)
            row[f"fused_{vname}_us"] = time_fn(lambda: op(a, wr, None, None))
            row["tiles"] = op.tiles
        buf = torch.empty(out_shape, dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)

        def unfused():
            yy = dense()
            dist.all_reduce(yy.permute(0, 2, 3, 1))
        row["unfused_conv_plus_nccl_us"] = time_fn(unfused)
        row["conv_only_us"] = time_fn(dense)
        row["nccl_allreduce_only_us"] = time_fn(lambda: dist.all_reduce(buf.permute(0, 2, 3, 1)))
        row["peer_allreduce_kernel_only_us"] = time_fn(lambda: f_mc.allreduce_bf16(buf))
        flops = 2.0 * out_shape[0] * (ho if kind == 0 else h) * (wo if kind == 0 else w) * out_shape[1] * k * 9
        if kind == 1 and stride == 2:
            flops /= 4
        out_bytes = buf.numel() * 2
        for vname, _ in variants:
            nv = out_bytes * (1 if vname == "nvls" else world - 1)
            roof = max(flops / PEAK_FLOPS, nv / NVLINK_BPS) * 1e6
            row[f"roofline_{vname}_us"] = roof
            row[f"frac_of_roofline_{vname}"] = roof / row[f"fused_{vname}_us"]
        best = min(row[f"fused_{v}_us"] for v, _ in variants)
        row["speedup_vs_conv_plus_nccl"] = row["unfused_conv_plus_nccl_us"] / best
        res["timing"].append(row)
        if rank == 0:
            print("TIME " + json.dumps(row), flush=True)

    # ---- GEMM -> reduce-scatter (tile t kept by rank t % W)
    n, cin, h, w, cout = 64, 512, 2, 2, 256
    if cin % world == 0:
        g = torch.Generator().manual_seed(7)
        k = cin // world
        full = torch.randn(n, cin, h, w, generator=g) * 0.5
        W_full = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        a = cl(full[:, rank * k:(rank + 1) * k].to(dev).bfloat16())
        wr = cl(W_full[:, rank * k:(rank + 1) * k].to(dev).bfloat16())
        ref, _ = tb.conv_fwd(a.float(), wr.float(), 1, 1, False)
        ref = ref.contiguous(); dist.all_reduce(ref)
        for vname, fz in variants[:1]:
            op = fz.reduce_scatter_conv(0, (n, k, h, w), (cout, k, 3, 3), 1, 1)
            for _ in range(3):
                y = op(a, wr)
            torch.cuda.synchronize()
            worst = 0.0
            for mt in range(2):
                for nt in range(4):
                    if (nt * 2 + mt) % world == rank:
                        got, want = y[32 * mt:32 * mt + 32, 64 * nt:64 * nt + 64], ref[32 * mt:32 * mt + 32, 64 * nt:64 * nt + 64]
                        worst = max(worst, ((got.float() - want).abs().max() / (want.abs().max() + 1e-6)).item())
            worst = rmax(worst)
            c = {"name": f"reduce_scatter_conv.{vname}", "rel_err": worst, "ok": bool(worst < 3e-2)}
            res["cases"].append(c)
            row = {"case": "reduce_scatter_conv", "x": [n, cin, h, w], "cout": cout,
                   f"fused_{vname}_us": time_fn(lambda: op(a, wr))}
            res["timing"].append(row)
            if rank == 0:
                print(("PASS " if c["ok"] else "FAIL ") + json.dumps(c), flush=True)
                print("TIME " + json.dumps(row), flush=True)

    # ---- all-gather -> GEMM: x image-sharded across ranks (in the symmetric heap), w local
    for (n, c, h, w, nout) in [(64, 64, 8, 8, 64), (64, 128, 4, 4, 128)]:
        g = torch.Generator().manual_seed(77 + rank)
        nl = n // world
        x_off, xbuf = f_mc.ag_buffer((nl, c, h, w))
        xs_ = cl((torch.randn(nl, c, h, w, generator=g) * 0.5).to(dev).bfloat16())
        xbuf.copy_(xs_)
        wl = cl((torch.randn(nout, c, 3, 3, generator=g) / (c * 9) ** 0.5).to(dev).bfloat16())
        op = f_mc.ag_conv(x_off, (n, c, h, w), (nout, c, 3, 3))
        parts = [torch.empty_like(xs_) for _ in range(world)]
        dist.all_gather(parts, xs_)
        xfull = cl(torch.cat(parts, dim=0))
        ref, _ = tb.conv_fwd(xfull.float(), wl.float(), 1, 1, False)
        torch.cuda.synchronize(); dist.barrier()
        for _ in range(3):
            y = op(None, wl)
        torch.cuda.synchronize()
        report(f"ag_conv_{(n, c, h, w, nout)}", y, ref, check_same=False)

        def unfused_ag():
            dist.all_gather(parts, xs_)
            nb.conv_fwd(cl(torch.cat(parts, dim=0)), wl, 1, 1, False)
        t_f = time_fn(lambda: op(None, wl))
        t_u = time_fn(unfused_ag)
        nv = (world - 1) / world * n * h * w * c * 2
        roof = max(2.0 * n * h * w * nout * c * 9 / PEAK_FLOPS, nv / NVLINK_BPS) * 1e6
        row = {"case": "ag_conv", "x": [n, c, h, w], "cout": nout, "fused_us": t_f, "nccl_allgather_plus_conv_us": t_u,
               "roofline_us": roof, "frac_of_roofline": roof / t_f, "speedup_vs_nccl_plus_conv": t_u / t_f}
        res["timing"].append(row)
        if rank == 0:
            print("TIME " + json.dumps(row), flush=True)

    # ---- the tensor-parallel head in one kernel vs PyTorch ops + NCCL
    n, c, classes = 64, 512, 10
    kpad = padded_classes(classes, world)
    kl = kpad // world
    g = torch.Generator().manual_seed(3)
    feat = cl(torch.randn(n, c, 1, 1, generator=g).to(dev).bfloat16())
    Wf, bf = torch.zeros(kpad, c), torch.zeros(kpad)
    Wf[:classes] = torch.randn(classes, c, generator=g) / c ** 0.5
    bf[:classes] = torch.randn(classes, generator=g) * 0.1
    labels = torch.randint(0, classes, (n,), generator=g).to(dev)
    Wl, bl = Wf[rank * kl:(rank + 1) * kl].contiguous().to(dev), bf[rank * kl:(rank + 1) * kl].contiguous().to(dev)
    pooled = feat.float().view(n, c).requires_grad_(True)
    logits = pooled @ Wf[:classes].to(dev).t() + bf[:classes].to(dev)
    lref = torch.nn.functional.cross_entropy(logits, labels)
    lref.backward()
    for vname, fz in variants:
        hop = fz.head(n, c, kl)
        dW, db = torch.zeros(kl, c, device=dev), torch.zeros(kl, device=dev)

        def run_head():
            return hop(feat, Wl, bl, labels, 1.0, classes, dW, db, False, True, None)
        for _ in range(3):
            loss, correct, dfeat, lg = run_head()
        torch.cuda.synchronize()
        report(f"tp_head.dfeat.{vname}", dfeat.view(n, c), pooled.grad, tol=1e-2)
        report(f"tp_head.logits.{vname}", lg[:, :classes].contiguous(), logits.detach(), tol=1e-3)
        c_ = {"name": f"tp_head.loss.{vname}", "rel_err": abs(loss.item() - lref.item()) / abs(lref.item()),
              "ok": bool(abs(loss.item() - lref.item()) < 2e-3 * abs(lref.item()))}
        res["cases"].append(c_)
        row = {"case": "tp_head", "variant": vname, "fused_us": time_fn(run_head)}
        res["timing"].append(row)
        if rank == 0:
            print(("PASS " if c_["ok"] else "FAIL ") + json.dumps(c_), flush=True)
            print("TIME " + json.dumps(row), flush=True)

    def eager_head():
        pl = feat.float().mean(dim=(2, 3))
        local = pl @ Wl.t() + bl
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        lg_ = torch.cat(parts, dim=1)
        lg_ = lg_.masked_fill(torch.arange(kpad, device=dev) >= classes, float("-inf"))
        p = torch.softmax(lg_, dim=1)
        p = p.scatter_add(1, labels.view(-1, 1), -torch.ones(n, 1, device=dev))
        dl = (p / n)[:, rank * kl:(rank + 1) * kl]
        dp = (dl @ Wl).contiguous()
        dist.all_reduce(dp)
        return dl.t() @ pl, dp
    row = {"case": "tp_head", "variant": "pytorch_ops_plus_nccl", "us": time_fn(eager_head)}
    res["timing"].append(row)
    if rank == 0:
        print("TIME " + json.dumps(row), flush=True)

    # ---- CUDA-graph replay of a fused kernel
    try:
        n, cs, h, w, nout = 64, 256 // world if 256 // world >= 8 else 8, 2, 2, 256
        g = torch.Generator().manual_seed(5 + rank)
        x = cl((torch.randn(n, cs, h, w, generator=g) * 0.5).to(dev).bfloat16())
        wf = cl((torch.randn(nout, cs, 3, 3, generator=g) / (cs * 9 * world) ** 0.5).to(dev).bfloat16())
        op = f_mc.allreduce_conv(0, (n, cs, h, w), (nout, cs, 3, 3), 1, 1)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            op(x, wf)
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            yg = op(x, wf)
        for _ in range(4):
            gr.replay()
        torch.cuda.synchronize()
        ref, _ = tb.conv_fwd(x.float(), wf.float(), 1, 1, False)
        ref = ref.contiguous(); dist.all_reduce(ref)
        res["graph_replay_ok"] = bool(((yg.float() - ref).abs().max() / ref.abs().max()).item() < 3e-2)
    except Exception as e:  # noqa: BLE001
        res["graph_replay_ok"] = False
        res["graph_exc"] = repr(e)
    res["n_fail"] = sum(1 for c in res["cases"] if not c["ok"])
    dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        json.dump(res, open(out_path, "w"), indent=1)
        print(f"tp_fused: {len(res['cases'])} cases, {res['n_fail']} failed, graph={res.get('graph_replay_ok')}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
