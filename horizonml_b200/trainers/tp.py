"""Tensor-parallel trainer (reference: tensor_parallel_train.py:155-296).

Every rank sees the same (seeded) batches.  Sharded parameters (column-parallel classifier,
channel-parallel layer3/4 convs) live in their own flat store and are never averaged; replicated
parameters are averaged with the bucketed fused all-reduce (the reference does 62 blocking
per-parameter all-reduces, tensor_…:215-218).  ``avg_bandwidth`` keeps the reference meaning (bytes
moved per step).
"""
from __future__ import annotations

import time
from typing import List

import torch
import torch.distributed as dist

from .. import checkpoint, ops
from ..config import TrainConfig
from ..data import BatchLoader, build_dataset
from ..metrics import EpochRecorder, write_summary
from ..models.flat import FlatAdam, FlatParams
from ..models.resnet import resnet18
from ..parallel.comm import make_grad_allreduce
from ..parallel.dp import GradReducer
from ..parallel.tp import TensorParallelResNet, TPComm
from .common import (DeviceStats, FaultInjector, GraphedStep, Heartbeat, Runtime, allreduce_max_scalar,
                     gpu_mem_mb, probe_step_regions, setup_runtime, split_compute_comm)


class TPEngine:
    """Tensor-parallel engine.  With a ``mesh`` (parallel/mesh.py) the tensor-parallel group is one row of a
    DP × TP process mesh: TP collectives run on ``mesh.tp_group``, replicated-parameter gradients are averaged over
    the whole world (identical inside a TP group, different across replicas) and sharded-parameter gradients over
    the rank's data-parallel group."""

    def __init__(self, cfg: TrainConfig, rt: Runtime, mesh=None):
        if cfg.model != "resnet18":
            # the reference's tensor-parallel script partitions torchvision's ResNet-18 and nothing else; MobileNetV2 is the
            # legacy container path's second model (train.py MODEL_TYPE) and runs on the data-parallel engine
            raise ValueError(f"--model {cfg.model}: the tensor-parallel strategy partitions ResNet-18 only "
                             "(use data_parallel_train.py / train.py for mobilenet)")
        self.cfg, self.rt, self.mesh = cfg, rt, mesh
        fused = None
        tp_group = mesh.tp_group if mesh is not None else None
        tp_world = mesh.tp if mesh is not None else rt.world
        if rt.device.type == "cuda" and rt.backend == "native" and tp_world > 1:
            from ..parallel.tp import FusedTP
            # one symmetric heap (+ NVSwitch multicast mapping when available) per tensor-parallel group: under a
            # DP × TP mesh every row gets its own
            fused = FusedTP(rt.device, group=tp_group)
        self.fused = fused
        self.comm = TPComm(group=tp_group, fused=fused)
        dense = resnet18(cfg.num_classes, seed=cfg.seed)
        self.model = TensorParallelResNet(dense, self.comm, cfg.tp_conv_split).to(rt.device)
        self.model.train()
        rep, shd = self.model.split_params()
        self.flat_rep = FlatParams(rep, rt.device, rt.dtype, cfg.bucket_mb)
        self.flat_shd = FlatParams(shd, rt.device, rt.dtype, cfg.bucket_mb)
        self.opt_rep = FlatAdam(self.flat_rep, lr=cfg.lr)
        self.opt_shd = FlatAdam(self.flat_shd, lr=cfg.lr)
        self.ar = make_grad_allreduce(cfg.allreduce, self.flat_rep.total, rt.device) if rt.world > 1 else None
        self.reducer = GradReducer(self.flat_rep, self.ar, cfg.overlap) if self.ar is not None else None
        self.ar_shd = None
        if mesh is not None and mesh.dp > 1:
            kind = cfg.allreduce if cfg.allreduce not in ("auto", "nvls") else "twoshot"
            self.ar_shd = make_grad_allreduce(kind, self.flat_shd.total, rt.device, group=mesh.dp_group)
        self.stats = DeviceStats(rt.device)
        # wgrad kernels depend only on (dy, x): side stream, off the dgrad / fused-reduction critical path
        ops.enable_side_stream(rt.device.type == "cuda" and rt.backend == "native")
        self.prev_grad = torch.zeros_like(self.flat_rep.grad) if cfg.grad_divergence else None
        # collectives inside the step (NCCL all-gather/all-reduce) are graph-capturable on CUDA
        self._graphed = GraphedStep(self._step_impl, rt.device, cfg.cuda_graph and rt.backend == "native")
        self.global_step = 0

    def _step_impl(self, images, labels):
        x = images
        if x.dtype == torch.uint8:
            x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
        ops.step_begin(self.rt.device)
        self.flat_rep.begin_step(); self.flat_shd.begin_step()
        if self.reducer is not None:
            self.reducer.begin_step()
        loss, correct = self.model.forward_loss(x, labels)
        ops.backward(loss)
        ops.join_side()
        if self.reducer is not None:
            self.reducer.finish()
        if self.ar_shd is not None:
            for bk in self.flat_shd.buckets:
                self.ar_shd.allreduce_avg_(self.flat_shd.grad[bk.start:bk.end])
        diff = self.opt_rep.step(prev_grad=self.prev_grad)
        self.opt_shd.step()
        self.stats.add_step(loss, correct, labels.shape[0], diff)
        ops.step_end()

    def step(self, images, labels):
        self._graphed(images, labels)
        self.global_step += 1

    def probe_regions(self, images, labels, iters: int = 10):
        """Device-timed step regions (trainers/common.py ``probe_step_regions``).  Forward and backward contain the fused
        GEMM+all-reduce / head kernels — the tensor-parallel traffic is *inside* those regions by construction —
        ``allreduce_ms`` is the replicated-parameter gradient averaging, ``exposed_comm_ms`` what the overlapped step
        costs beyond forward + backward + optimizer.  All ranks call it together."""
        dev = self.rt.device
        state = [self.flat_rep.master, self.flat_rep.grad, self.flat_shd.master, self.flat_shd.grad, self.opt_rep.m,
                 self.opt_rep.v, self.opt_rep.step_t, self.opt_shd.m, self.opt_shd.v, self.opt_shd.step_t, self.stats.buf,
                 self.stats.has_prev]
        for fl in (self.flat_rep, self.flat_shd):
            if fl.shadow is not None:
                state.append(fl.shadow)
        if self.prev_grad is not None:
            state.append(self.prev_grad)
        state += [b for b in self.model.buffers()]

        def fwd(x, y):
            if x.dtype == torch.uint8:
                x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
            ops.step_begin(dev)
            self.flat_rep.begin_step(); self.flat_shd.begin_step()
            return self.model.forward_loss(x, y)

        def comm():
            if self.ar is not None:
                for bk in self.flat_rep.buckets:
                    self.ar.allreduce_avg_(self.flat_rep.grad[bk.start:bk.end])
            if self.ar_shd is not None:
                for bk in self.flat_shd.buckets:
                    self.ar_shd.allreduce_avg_(self.flat_shd.grad[bk.start:bk.end])

        def opt():
            self.opt_rep.step(prev_grad=self.prev_grad)
            self.opt_shd.step()

        def restore():
            self.flat_rep.begin_step(); self.flat_shd.begin_step()

        return probe_step_regions(dev, state, [self.reducer], self._graphed, fwd, comm, opt, restore, images, labels, iters)

    def bytes_per_step(self) -> int:
        rep = sum(self.ar.wire_bytes(b.end - b.start) for b in self.flat_rep.buckets) if self.ar else 0
        return rep


def _native_fallbacks(rt) -> dict:
    if rt.backend != "native":
        return {}
    from ..ops import native_backend as nb
    return dict(nb.FALLBACKS)


def train_tensor_parallel(rank: int, world: int, cfg: TrainConfig, device: str):
    rt = setup_runtime(rank, world, cfg, device)
    logs_dir = cfg.resolved_logs_dir()
    images, labels = build_dataset(cfg.sample_size, cfg.synthetic, cfg.data_dir, cfg.seed)
    if rank == 0 and not cfg.quiet:
        print("Worker 0 generated the synthetic dataset." if cfg.synthetic else
              "Worker 0 downloaded the dataset.", flush=True)
    mesh = None
    if cfg.dp_replicas > 1:
        from ..parallel.mesh import DeviceMesh
        if world % cfg.dp_replicas:
            raise ValueError(f"--dp_replicas {cfg.dp_replicas} does not divide world_size {world}")
        mesh = DeviceMesh(world, rank, dp=cfg.dp_replicas, tp=world // cfg.dp_replicas)
    sampler = None                       # all ranks of a TP group: same data (tensor_…:143); replicas: disjoint shards
    if mesh is not None:
        from ..data import ShardedSampler
        sampler = ShardedSampler(len(labels), mesh.dp, mesh.coord.dp, shuffle=False, seed=cfg.seed)
    loader = BatchLoader(images, labels, cfg.batch_size, rt.device, sampler=sampler)
    eng = TPEngine(cfg, rt, mesh)
    rec = EpochRecorder("tensor", rank, logs_dir, cfg.sample_size)
    hb = Heartbeat(cfg.heartbeat_dir, rank)
    fault = FaultInjector(cfg.inject_fault, rank)
    tag = f"tp_rank{rank}of{world}" if mesh is None else f"tp_rank{mesh.coord.tp}of{mesh.tp}"
    saver = mesh is None or mesh.coord.dp == 0
    start_epoch = 0
    if cfg.resume:
        payload = checkpoint.load(cfg.resume, tag, eng.model, None)
        if payload is not None:
            start_epoch, eng.global_step = payload["epoch"], payload["global_step"]
            ex = payload.get("extra", {})
            if "opt_rep" in ex:
                eng.opt_rep.load_state_dict(ex["opt_rep"]); eng.opt_shd.load_state_dict(ex["opt_shd"])
            eng.flat_rep.sync_shadow(); eng.flat_shd.sync_shadow()
    if not cfg.quiet:
        print(f"Worker {rank} is starting training...", flush=True)
    cuda = rt.device.type == "cuda"
    probe_xy, regions = None, {}
    for epoch in range(start_epoch, cfg.epochs):
        t_epoch = time.time()
        t0 = time.time()
        if world > 1:
            dist.barrier()
        rec.total_idle += time.time() - t0
        step_times: List[float] = []
        nsteps, tp_bytes = 0, 0
        if cuda:
            torch.cuda.reset_peak_memory_stats(rt.device)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for bi, (x, y) in enumerate(loader):
            if cfg.max_steps and bi >= cfg.max_steps:
                break
            if probe_xy is None or (x.shape[0] == cfg.batch_size and probe_xy[0].shape[0] != cfg.batch_size):
                probe_xy = (x.clone(), y.clone())
            ts = time.time()
            rec.host.sample()
            fault.maybe_fail(eng.global_step)
            eng.step(x, y)
            if cfg.step_barrier and world > 1:
                ti = time.time()
                if cuda:
                    torch.cuda.synchronize()
                dist.barrier()
                rec.total_idle += time.time() - ti
            step_times.append(time.time() - ts)
            nsteps += 1
            if bi % 50 == 0:
                hb.beat(epoch, eng.global_step)
        if cuda:
            ev1.record()
            torch.cuda.synchronize()
            dev_s = ev0.elapsed_time(ev1) / 1e3
        else:
            dev_s = time.time() - t_epoch
        tp_bytes = eng.comm.take_bytes()
        s = eng.stats.read_and_reset()
        epoch_time = time.time() - t_epoch
        steps = max(int(s["steps"]), 1)
        loss = s["loss_sum"] / steps
        acc = 100.0 * s["correct"] / max(s["seen"], 1)
        if s["grad_div_n"] > 0:
            rec.grad_divs = [s["grad_div_sum"] / s["grad_div_n"]]
        dev_s_max = allreduce_max_scalar(dev_s, rt.device)
        if epoch == start_epoch and cfg.region_probe:
            have = allreduce_max_scalar(0.0 if probe_xy is not None else 1.0, rt.device) == 0.0
            if have:
                try:
                    regions = eng.probe_regions(*probe_xy)
                except Exception as e:  # noqa: BLE001 — a diagnostics feature must never take the run down
                    if rank == 0:
                        print(f"[probe] region breakdown unavailable: {e!r}", flush=True)
                    regions = {}
        comp_s, comm_s, src = split_compute_comm(dev_s, regions)
        rec.total_compute += comp_s
        rec.total_comm += comm_s
        bytes_per_step = eng.bytes_per_step() + (tp_bytes / steps if eng._graphed.graph is None else 0)
        seen = s["seen"] * (mesh.dp if mesh is not None else 1)
        ext = {"images_per_sec": seen / dev_s_max if dev_s_max > 0 else 0, "steps": nsteps,
               "gpu_mem_MB": gpu_mem_mb(rt.device),
               "nvlink_GBps": bytes_per_step * nsteps / dev_s_max / 1e9 if dev_s_max > 0 else 0}
        ext.update({k: v for k, v in regions.items() if k != "step_ms"})
        ext["split_source"] = src
        if cuda:
            step_times = [dev_s / max(nsteps, 1)] * nsteps
        rec.end_epoch(epoch + 1, loss, acc, epoch_time, step_times, avg_bandwidth=bytes_per_step, ext=ext)
        if not cfg.quiet:   # the reference prints the epoch line on every rank (tensor_…:264)
            print(f"Epoch [{epoch+1}/{cfg.epochs}], Loss: {loss:.4f}, Accuracy: {acc:.2f}%, "
                  f"Time: {epoch_time:.2f}s", flush=True)
        if cfg.save_dir and saver and ((cfg.save_every and (epoch + 1) % cfg.save_every == 0) or epoch + 1 == cfg.epochs):
            checkpoint.save(cfg.save_dir, tag, eng.model, None, epoch + 1, eng.global_step,
                            extra={"opt_rep": eng.opt_rep.state_dict(), "opt_shd": eng.opt_shd.state_dict()})
        if world > 1:
            dist.barrier()
    if rank == 0:
        write_summary(logs_dir, f"summary_{cfg.sample_size}.json", {
            "strategy": "tensor", "world_size": world, "backend": rt.backend, "dtype": str(rt.dtype),
            "conv_split": eng.model.conv_split, "mesh": mesh.describe() if mesh is not None else None,
            "fused_tp": eng.fused.describe() if eng.fused is not None else None,
            "fused_tp_ops": eng.fused.ops if eng.fused is not None else None,
            "library_collectives_in_step": eng.comm.library_collectives,
            "native_fallbacks": _native_fallbacks(rt),
            "replicated_grad_allreduce": getattr(eng.ar, "name", None),
            "graph": eng._graphed.graph is not None, "graph_error": eng._graphed.capture_error,
            "final": rec.rows[-1] if rec.rows else None})
    # a captured graph that contains NCCL kernels must be gone before the communicator is torn down
    eng._graphed.graph = None
    if cuda:
        torch.cuda.synchronize()
    from ..launch import shutdown_distributed
    shutdown_distributed()
    return rec.frame()
