"""Data-parallel trainer.

Reference call stack: data_parallel_train.py ``worker`` → ``train`` (:76-189, :192-230): per step
``zero_grad → model(images) → CE → backward (DDP bucket all-reduce) → Adam.step → loss.item() →
argmax accuracy → grad clone+cat divergence → dist.barrier()``.

B200 step (``DPEngine.step``): one CUDA-graph replay containing
``stem-im2col(normalise+cast) → 20× [tcgen05 conv(+BN sums) → BN/ReLU/residual] → fused
FC+CE+accuracy(+its own backward) → 20× [BN-bwd reduce/apply → wgrad (into the bucket) → dgrad] →
per-bucket fused cast/scale peer all-reduce on the comm stream (overlapped) → fused Adam (+bf16
shadow refresh) → grad-divergence``; loss/accuracy accumulate on the device and are read once per
epoch.  No host barrier per step (flag ``--step_barrier`` restores the reference behaviour).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import checkpoint, ops
from ..config import TrainConfig
from ..data import BatchLoader, ShardedSampler, build_dataset
from ..metrics import EpochRecorder, write_summary
from ..models.flat import FlatAdam, FlatParams
from ..models.resnet import resnet18
from ..parallel.comm import make_grad_allreduce
from ..parallel.dp import GradReducer
from .common import (DeviceStats, FaultInjector, GraphedStep, Heartbeat, Runtime, allreduce_max_scalar,
                     enable_nvtx, gpu_mem_mb, nvtx_range, probe_step_regions, profile_steps, setup_runtime,
                     split_compute_comm)


def build_model(cfg: TrainConfig, device, class_pad_to: int = 1):
    if cfg.model == "mobilenet":
        from ..models.mobilenet import mobilenet_v2
        return mobilenet_v2(cfg.num_classes, seed=cfg.seed).to(device)
    return resnet18(cfg.num_classes, seed=cfg.seed, class_pad_to=class_pad_to).to(device)


class DPEngine:
    """Public data-parallel training engine: ``engine.step(images_u8_nhwc, labels)``.

    ``images`` uint8 NHWC on CUDA (normalisation fused into the stem) or normalised fp32 on CPU."""

    def __init__(self, cfg: TrainConfig, rt: Runtime):
        self.cfg, self.rt = cfg, rt
        dev = rt.device
        self.model = build_model(cfg, dev)
        self.model.train()
        live = self.model.live_tap_masks(32) if (cfg.skip_dead_taps and hasattr(self.model, "live_tap_masks")) else None
        by_live = live is not None and (bool(cfg.bucket_by_live) or os.environ.get("HZ_BUCKET_LIVE", "0") == "1")
        self.zero1 = bool(cfg.zero1) and rt.world > 1
        # ZeRO-1 as ONE peer-memory kernel per bucket (reduce -> Adam on the owned shard -> bf16 parameter push) where
        # the peer kernels run; "fused" on CPU selects the same sharding on torch.distributed (tests / plumbing)
        peer_ok = (rt.world > 1 and dev.type == "cuda" and cfg.allreduce != "nccl" and rt.backend == "native"
                   and rt.dtype == torch.bfloat16)
        self.zero_fused = self.zero1 and (cfg.zero1_impl == "fused" or (cfg.zero1_impl == "auto" and peer_ok))
        # Adam inside the all-reduce kernel (csrc/comm.cu AdamFuse): peer kernels only, replicated optimizer state
        self.fused_adam = (bool(cfg.fused_adam) and rt.world > 1 and not self.zero1 and dev.type == "cuda"
                           and cfg.allreduce != "nccl" and rt.backend == "native" and not cfg.overlap_adam
                           and os.environ.get("HZ_OVERLAP_ADAM", "0") != "1")
        layout = cfg.bucket_layout
        if layout == "auto":
            # layer-group buckets: every group's all-reduce starts the moment its backward is done and the bucket that
            # is only complete after the very last gradient kernel (layer1 + stem, 157 k elements) is small enough for
            # the flag-in-data latency protocol — measured 0.601 vs 0.605 ms/step at 2 GPUs against DDP-like size caps
            peer = (rt.world > 1 and dev.type == "cuda" and cfg.allreduce != "nccl" and rt.backend == "native"
                    and (not self.zero1 or self.zero_fused))
            layout = "layers" if (peer and cfg.model == "resnet18") else "size"
        starts = ("layer3.", "layer2.", "layer1.") if layout == "layers" else None
        self.flat = FlatParams(list(self.model.named_parameters()), dev, rt.dtype,
                               cfg.live_bucket_mb if by_live else cfg.bucket_mb, live_masks=live, bucket_by_live=by_live,
                               pad_multiple=64 * rt.world if self.zero1 else 64, bucket_starts=starts)
        if rt.world > 1:   # K1: make replicas identical (same seed already does; belt and braces)
            dist.broadcast(self.flat.master, src=0)
            for b in self.model.buffers():
                if b.dtype.is_floating_point:
                    dist.broadcast(b, src=0)
            self.flat.sync_shadow()
        kind = cfg.allreduce
        if self.zero_fused:
            from ..parallel.zero import FusedShardedAdam
            self.ar = make_grad_allreduce(kind, self.flat.total, dev) if peer_ok else None   # the kernel's communicator
            self.opt = FusedShardedAdam(self.flat, self.ar if hasattr(self.ar, "handle") else None, lr=cfg.lr,
                                        with_prev=bool(cfg.grad_divergence))
        elif self.zero1:
            from ..parallel.zero import ShardedFlatAdam
            self.opt = ShardedFlatAdam(self.flat, lr=cfg.lr)      # reduce-scatter + sharded Adam + all-gather
            self.ar = None
        else:
            self.opt = FlatAdam(self.flat, lr=cfg.lr)
            self.ar = make_grad_allreduce(kind, self.flat.total, dev) if rt.world > 1 else None
        self.stats = DeviceStats(dev)
        ops.enable_side_stream(dev.type == "cuda" and rt.backend == "native")
        self.prev_grad = None
        if cfg.grad_divergence and not self.zero_fused:         # (the fused ZeRO optimizer keeps its own shards of it)
            self.prev_grad = torch.zeros_like(self.opt.gshard if self.zero1 else self.flat.grad)
        # bucket-wise optimizer: Adam for a bucket runs right behind that bucket's all-reduce (or, on one GPU, as
        # soon as its gradients are final) and overlaps the rest of backward
        self.bucket_adam = (bool(cfg.overlap_adam) or os.environ.get("HZ_OVERLAP_ADAM", "0") == "1") and not self.zero1
        self._diff_acc = (torch.zeros((), dtype=torch.float32, device=dev)
                          if (self.prev_grad is not None or (self.zero_fused and cfg.grad_divergence)) else None)
        self.fused_adam = self.fused_adam and hasattr(self.ar, "allreduce_adam_")
        if self.fused_adam and self._diff_acc is None and self.prev_grad is not None:
            self._diff_acc = torch.zeros((), dtype=torch.float32, device=dev)
        self.reducer = None
        if self.ar is not None or self.bucket_adam:
            fused = self._fused_bucket if self.fused_adam else (self._zero_bucket if self.zero_fused else None)
            self.reducer = GradReducer(self.flat, self.ar, cfg.overlap,
                                       post_bucket=self._adam_bucket if self.bucket_adam else None, fused_bucket=fused)
        self._graphed = GraphedStep(self._step_impl, dev, cfg.cuda_graph)
        self.global_step = 0

    def _fused_bucket(self, b: int, last: bool):
        """One kernel: average bucket ``b``'s gradients over the ranks and apply Adam to its parameters."""
        f, o = self.flat, self.opt
        bk = f.buckets[b]
        sl = slice(bk.start, bk.end)
        return self.ar.allreduce_adam_(f.grad[sl], f.master[sl], o.m[sl], o.v[sl],
                                       f.shadow[sl] if f.shadow is not None else None,
                                       self.prev_grad[sl] if self.prev_grad is not None else None,
                                       self._diff_acc if self.prev_grad is not None else None, o.step_t, o.lr,
                                       o.betas[0], o.betas[1], o.eps, bump=last, live=f.bucket_live[b])

    def _zero_bucket(self, b: int, last: bool):
        """One kernel: reduce bucket ``b`` over the ranks, Adam on the shard this rank owns, new bf16 parameters to all."""
        return self.opt.step_bucket(b, last, self._diff_acc)

    def _adam_bucket(self, b: int, first: bool) -> None:
        self.opt.step_bucket(b, first, diff_out=self._diff_acc, prev_grad=self.prev_grad)

    # one full training step; everything inside is stream-ordered and graph-capturable
    def _step_impl(self, images, labels) -> None:
        cfg = self.cfg
        x = images
        if x.dtype == torch.uint8:
            x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
        ops.step_begin(self.rt.device)
        self.flat.begin_step()
        if self.reducer is not None:
            self.reducer.begin_step()
        with nvtx_range("forward"):
            loss, correct = self.model.forward_loss(x, labels)
        with nvtx_range("backward"):
            ops.backward(loss)
            ops.join_side()
        with nvtx_range("allreduce_join"):
            if self.reducer is not None:
                self.reducer.finish()
        with nvtx_range("optimizer"):
            if self.bucket_adam or self.fused_adam:
                diff = self._diff_acc          # every bucket's Adam has been joined by reducer.finish()
            elif self.zero_fused:
                if self.reducer is None:       # plumbing path (no peer communicator): buckets in gradient-ready order
                    nb_ = len(self.flat.buckets)
                    for b in range(nb_):
                        self.opt.step_bucket(b, b == nb_ - 1, self._diff_acc)
                diff = self._diff_acc
            else:
                diff = self.opt.step(prev_grad=self.prev_grad)
        self.stats.add_step(loss, correct, labels.shape[0], diff)
        ops.step_end()

    def step(self, images, labels) -> None:
        self._graphed(images, labels)
        self.global_step += 1

    # ------------------------------------------------------------------------------------------
    # device-timed breakdown of the step (SURVEY §5.1 / §5.5 extended columns)
    # ------------------------------------------------------------------------------------------
    def _state_tensors(self) -> List[torch.Tensor]:
        if self.zero_fused:
            ts = [self.flat.master, self.flat.grad] + self.opt.state_tensors() + [self.stats.buf, self.stats.has_prev]
        else:
            ts = [self.flat.master, self.flat.grad, self.opt.m, self.opt.v, self.opt.step_t, self.stats.buf,
                  self.stats.has_prev]
        if self.zero1 and not self.zero_fused:
            ts.append(self.opt.gshard)
        if self.flat.shadow is not None:
            ts.append(self.flat.shadow)
        if self.prev_grad is not None:
            ts += [self.prev_grad, self._diff_acc]
        elif self._diff_acc is not None:
            ts.append(self._diff_acc)
        ts += [b for b in self.model.buffers()]
        return ts

    def probe_regions(self, images, labels, iters: int = 10) -> Dict[str, float]:
        """{fwd_ms, bwd_ms, allreduce_ms, optimizer_ms, exposed_comm_ms, step_ms}: where one training step spends its
        time (trainers/common.py ``probe_step_regions``).  All ranks must call it together."""
        dev = self.rt.device

        def fwd(x, y):
            ops.step_begin(dev)
            self.flat.begin_step()
            if x.dtype == torch.uint8:
                x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
            return self.model.forward_loss(x, y)

        def comm():
            if self.zero_fused:                  # reduction and sharded optimizer are one kernel per bucket
                nb_ = len(self.flat.buckets)
                for b in range(nb_):
                    self.opt.step_bucket(b, b == nb_ - 1, self._diff_acc)
                return
            if self.ar is None:
                return
            for bk in self.flat.buckets:
                self.ar.allreduce_avg_(self.flat.grad[bk.start:bk.end], live=self.flat.bucket_live[bk.index])

        def restore():
            self.flat.begin_step()

        opt_step = (lambda: None) if self.zero_fused else (lambda: self.opt.step(prev_grad=self.prev_grad))
        return probe_step_regions(dev, self._state_tensors(), [self.reducer], self._graphed, fwd, comm,
                                  opt_step, restore, images, labels, iters)

    def input_buffers(self):
        """Static (images, labels) buffers of the captured step, or None (eager mode / before capture)."""
        return self._graphed.input_buffers()

    def allreduce_bytes_per_step(self) -> int:
        if self.ar is None:
            return 0
        return sum(self.ar.wire_bytes(b.end - b.start if self.flat.bucket_live[b.index] is None
                                      else self.flat.bucket_live[b.index].numel() * 64) for b in self.flat.buckets)


def train_data_parallel(rank: int, world: int, cfg: TrainConfig, device: str):
    """Per-rank worker (the reference's ``worker`` + ``train``)."""
    rt = setup_runtime(rank, world, cfg, device)
    if cfg.profile:
        enable_nvtx(True)                      # --profile: NVTX ranges around forward / backward / collectives / optimizer
    logs_dir = cfg.resolved_logs_dir()
    images, labels = build_dataset(cfg.sample_size, cfg.synthetic, cfg.data_dir, cfg.seed)
    if rank == 0 and not cfg.quiet:
        print("Worker 0 downloaded the dataset." if not cfg.synthetic else
              "Worker 0 generated the synthetic dataset.", flush=True)
    sampler = ShardedSampler(len(labels), world, rank, shuffle=True, seed=cfg.seed)
    loader = BatchLoader(images, labels, cfg.batch_size, rt.device, sampler)
    eng = DPEngine(cfg, rt)
    rec = EpochRecorder("data", rank, logs_dir, cfg.sample_size)
    hb = Heartbeat(cfg.heartbeat_dir, rank)
    fault = FaultInjector(cfg.inject_fault, rank)
    start_epoch = 0
    if cfg.resume:
        payload = checkpoint.load(cfg.resume, "dp", eng.model, eng.opt)
        if payload is not None:
            start_epoch = payload["epoch"]
            eng.global_step = payload["global_step"]
    if not cfg.quiet:
        print(f"Worker {rank} is starting training...", flush=True)
    cuda = rt.device.type == "cuda"
    df = None
    probe_xy = None          # one full-size batch for the region probe
    regions: Dict[str, float] = {}
    dev_idle = dev_stamps = None
    if cfg.step_barrier and world > 1 and hasattr(getattr(eng, "ar", None), "handle"):
        dev_stamps = torch.zeros(2, dtype=torch.int64, device=rt.device)
        dev_idle = torch.zeros((), dtype=torch.int64, device=rt.device)
    for epoch in range(start_epoch, cfg.epochs):
        sampler.set_epoch(epoch)
        t_epoch = time.time()
        t0 = time.time()
        if world > 1:
            dist.barrier()
        rec.total_idle += time.time() - t0
        step_times: List[float] = []
        ev_start = ev_end = None
        if cuda:
            torch.cuda.reset_peak_memory_stats(rt.device)
            ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_start.record()
        nsteps = 0
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
                if dev_idle is not None:
                    # device-side flag barrier over NVLink: no host sync; idle = release − arrival
                    eng.ar.handle.barrier(dev_stamps)
                    dev_idle += (dev_stamps[1] - dev_stamps[0])
                else:
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
            ev_end.record()
            torch.cuda.synchronize()
            dev_s = ev_start.elapsed_time(ev_end) / 1e3
        else:
            dev_s = time.time() - t_epoch
        s = eng.stats.read_and_reset()
        if hasattr(eng.ar, "check_error"):
            eng.ar.check_error()              # a rank that never arrived at a peer barrier is an error, not a hang
        if dev_idle is not None:
            rec.total_idle += float(dev_idle.item()) * 1e-9
            dev_idle.zero_()
        epoch_time = time.time() - t_epoch
        steps = max(int(s["steps"]), 1)
        loss = s["loss_sum"] / steps
        acc = 100.0 * s["correct"] / max(s["seen"], 1)
        if s["grad_div_n"] > 0:
            rec.grad_divs = [s["grad_div_sum"] / s["grad_div_n"]]
        dev_s_max = allreduce_max_scalar(dev_s, rt.device)
        seen_all = s["seen"] * world
        ar_bytes = eng.allreduce_bytes_per_step()
        ext = {"images_per_sec": seen_all / dev_s_max if dev_s_max > 0 else 0,
               "gpu_mem_MB": gpu_mem_mb(rt.device), "steps": nsteps,
               "nvlink_GBps": (ar_bytes * nsteps / dev_s_max / 1e9) if dev_s_max > 0 else 0}
        if cuda:
            step_times = [dev_s / max(nsteps, 1)] * nsteps
        if epoch == start_epoch and cfg.region_probe:
            # every rank takes the same decision (the probe contains collectives)
            have = allreduce_max_scalar(0.0 if probe_xy is not None else 1.0, rt.device) == 0.0
            if have:
                try:
                    regions = eng.probe_regions(*probe_xy)
                except Exception as e:  # noqa: BLE001 — a diagnostics feature must never take the run down
                    if rank == 0:
                        print(f"[probe] region breakdown unavailable: {e!r}", flush=True)
                    regions = {}
        ext.update({k: v for k, v in regions.items() if k != "step_ms"})
        # reference columns (Q8: "compute_time" = forward, "comm_time" = backward + all-reduce + optimizer), filled from
        # the device-timed regions — not a fixed ratio
        comp_s, comm_s, src = split_compute_comm(dev_s, regions)
        rec.total_compute += comp_s
        rec.total_comm += comm_s
        ext["split_source"] = src
        rec.end_epoch(epoch + 1, loss, acc, epoch_time, step_times, ext=ext)
        if rank == 0 and not cfg.quiet:
            print(f"Epoch [{epoch+1}/{cfg.epochs}], Loss: {loss:.4f}, Accuracy: {acc:.2f}%, "
                  f"Time: {epoch_time:.2f}s", flush=True)
        if cfg.profile and epoch == start_epoch and nsteps > 0:
            # the profiled steps are real optimizer steps: snapshot the training state and put it back afterwards, so a
            # profiled run trains (and checkpoints) exactly like an unprofiled one
            state = eng._state_tensors()
            snap, gs = [t.clone() for t in state], eng.global_step
            profile_steps(eng.step, x, y, f"{logs_dir}/profile_rank{rank}.json")
            if cuda:
                torch.cuda.synchronize()
            for t, s_ in zip(state, snap):
                t.copy_(s_)
            eng.global_step = gs
        if cfg.save_dir and ((cfg.save_every and (epoch + 1) % cfg.save_every == 0) or epoch + 1 == cfg.epochs):
            # ZeRO-1: the moments are sharded — every rank takes part in gathering them, rank 0 writes the file
            optim_state = eng.opt.gather_state() if eng.zero1 else None
            if rank == 0:
                checkpoint.save(cfg.save_dir, "dp", eng.model, eng.opt, epoch + 1, eng.global_step,
                                optim_state=optim_state)
        if world > 1:
            dist.barrier()
    df = rec.frame()
    if rank == 0:
        write_summary(logs_dir, f"summary_{cfg.sample_size}.json", {
            "strategy": "data", "world_size": world, "backend": rt.backend, "dtype": str(rt.dtype),
            "comm": rt.comm_backend, "allreduce": getattr(eng.ar, "name", None),
            "allreduce_detail": eng.ar.describe() if hasattr(eng.ar, "describe") else None,
            "fused_adam": eng.fused_adam,
            "zero1": ("fused-kernel" if getattr(eng.opt, "native", False) else "fused-dist") if eng.zero_fused
                     else ("nccl" if eng.zero1 else None),
            "bucket_algos": eng.reducer.algos if eng.reducer is not None else None,
            "buckets": [[b.names[0], b.names[-1], b.end - b.start] for b in eng.flat.buckets],
            "graph": eng._graphed.graph is not None, "graph_error": eng._graphed.capture_error,
            "final": rec.rows[-1] if rec.rows else None})
    eng._graphed.graph = None
    if cuda:
        torch.cuda.synchronize()
    from ..launch import shutdown_distributed
    shutdown_distributed()
    return df
