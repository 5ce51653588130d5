"""Layer / pipeline-parallel trainer (reference: layer_model_parallel_train.py:134-334).

Each rank owns a contiguous chunk of the five atomic blocks (``partition_blocks`` — the reference's
``SplitResNet`` rule), its own flat parameter store and its own fused Adam.  Per optimizer step the
batch is split into ``--microbatches`` micro-batches run 1F1B (``parallel.pp.PipelineRunner``):
activations down, gradients up, over NCCL p2p.  Reference CSV conventions are kept: only the last
stage reports loss/accuracy/grad_divergence (others write 0, layer_…:304-317), ``avg_bandwidth`` is
bytes sent per step, ``comm_time`` is time spent in p2p (blocking share).
"""
from __future__ import annotations

import time
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .. import checkpoint, ops
from ..config import TrainConfig
from ..data import BatchLoader, build_dataset
from ..metrics import EpochRecorder, write_summary
from ..models.flat import FlatAdam, FlatParams
from ..models.partition import boundary_shape, partition_blocks
from ..models.resnet import resnet18
from ..parallel.pp import GraphedMicroBatch, OverlappedPipelineRunner, P2PChannels, PipelineRunner
from .common import (DeviceStats, FaultInjector, Heartbeat, Runtime, allreduce_max_scalar, gpu_mem_mb,
                     setup_runtime)


class PPEngine:
    """One pipeline stage.  With a ``mesh`` (parallel/mesh.py) the pipeline is one row of a DP × PP process mesh:
    stage index / neighbours come from the mesh and the stage's gradients are averaged over its data-parallel
    group (fused peer all-reduce on GPUs) between the 1F1B schedule and the optimizer pass."""

    def __init__(self, cfg: TrainConfig, rt: Runtime, mesh=None):
        if cfg.model != "resnet18":
            # the reference's pipeline script partitions torchvision's ResNet-18 and nothing else; MobileNetV2 is the
            # legacy container path's second model (train.py MODEL_TYPE) and runs on the data-parallel engine
            raise ValueError(f"--model {cfg.model}: the pipeline strategy partitions ResNet-18 only "
                             "(use data_parallel_train.py / train.py for mobilenet)")
        self.cfg, self.rt, self.mesh = cfg, rt, mesh
        S, s = (mesh.pp, mesh.coord.pp) if mesh is not None else (rt.world, rt.rank)
        self.S, self.s = S, s
        self.first_blk, self.last_blk = partition_blocks(S)[s]
        self.is_first, self.is_last = s == 0, s == S - 1
        full = resnet18(cfg.num_classes, seed=cfg.seed)     # identical init on every stage (seeded)
        self.model = full.to(rt.device)
        self.model.train()
        names = set()
        groups = self.model.block_param_names()
        for b in range(self.first_blk, self.last_blk + 1):
            names.update(groups[b])
        mine = [(n, p) for n, p in self.model.named_parameters() if n in names]
        live = self.model.live_tap_masks(32) if cfg.skip_dead_taps else None
        # parameters of other stages are dropped (freed) — this stage never touches them
        for n, p in self.model.named_parameters():
            if n not in names:
                p.requires_grad_(False)
                p.data = torch.empty(0, device=rt.device)
        self.flat = FlatParams(mine, rt.device, rt.dtype, cfg.bucket_mb, live_masks=live)
        self.opt = FlatAdam(self.flat, lr=cfg.lr)
        self.stats = DeviceStats(rt.device)
        self.prev_grad = torch.zeros_like(self.flat.grad) if (cfg.grad_divergence and self.is_last) else None
        self.labels_mb: List[torch.Tensor] = []
        self.mb_frac: List[float] = []
        hw = 32
        self.runner = PipelineRunner(
            s, S, self._fwd,
            in_shape=lambda n: boundary_shape(self.first_blk - 1, n, hw),
            out_shape=lambda n: boundary_shape(self.last_blk, n, hw),
            dtype=rt.dtype, device=rt.device, peers=mesh.pp_neighbours() if mesh is not None else None)
        self.dp_ar = None
        if mesh is not None and mesh.dp > 1:
            from ..parallel.comm import make_grad_allreduce
            kind = cfg.allreduce if cfg.allreduce not in ("auto", "nvls") else "twoshot"
            self.dp_ar = make_grad_allreduce(kind, self.flat.total, rt.device, group=mesh.dp_group)
        self.global_step = 0
        # CUDA-graphed micro-batches (GPU + native kernels): built lazily after one eager step
        self.use_graphs = cfg.cuda_graph and rt.device.type == "cuda" and rt.backend == "native"
        self.slots = None
        self.slot_sizes = None
        # overlapped 1F1B (GPU + graphs): one NCCL communicator and one CUDA stream per neighbour and direction, so
        # receives are posted ahead of the compute that needs them and sends fire from events (parallel/pp.py)
        self.channels = None
        self.overlapped = None
        if self.use_graphs and S > 1 and rt.comm_backend == "nccl" and bool(getattr(cfg, "pp_overlap", True)):
            if mesh is not None:
                rows = [[mesh.rank_of(d, st, t) for st in range(mesh.pp)] for d in range(mesh.dp) for t in range(mesh.tp)]
            else:
                rows = [list(range(S))]
            self.channels = P2PChannels(rt.rank, rows, rt.device)
        ops.enable_side_stream(rt.device.type == "cuda" and rt.backend == "native")

    def _fwd(self, x, i):
        if self.is_first and x.dtype == torch.uint8:
            x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
        f = self.model.features(x, self.first_blk, self.last_blk)
        if self.is_last:
            return ops.head_loss(f, self.model.fc.weight, self.model.fc.bias, self.labels_mb[i],
                                 loss_scale=self.mb_frac[i], n_valid=self.cfg.num_classes)
        return f

    # ---- graphed micro-batches ------------------------------------------------------------------
    def _stage_fwd_static(self, x, labels):
        """Forward of this stage on static buffers (captured): returns y, or (loss, correct) on the last stage."""
        if self.is_first and x.dtype == torch.uint8:
            x = ops.stem_prepare(x.permute(0, 3, 1, 2), dtype=self.rt.dtype)
        f = self.model.features(x, self.first_blk, self.last_blk)
        if self.is_last:
            return ops.head_loss(f, self.model.fc.weight, self.model.fc.bias, labels,
                                 loss_scale=self._graph_frac, n_valid=self.cfg.num_classes)
        return f

    def _build_slots(self, sizes, images):
        from ..ops import native_backend as nb
        S, s = self.S, self.s
        n = sizes[0]
        self._graph_frac = 1.0 / len(sizes)
        # 1F1B keeps min(S - s, M) micro-batches in flight; one spare slot lets the next activation be received
        # (straight into the slot's static input) while the in-flight ones are still being computed
        nslots = max(1, min(S - s + (1 if self.channels is not None else 0), len(sizes)))
        nb.ARENA.active = False                      # captured micro-batches are replayed several times per step:
        for p in self.flat.params:                   # they must not share pre-zeroed scratch, and every gradient
            p._acc = True                            # write accumulates (the optimizer pass clears the buffer)
        slots = []
        torch.cuda.synchronize()
        for _ in range(nslots):
            # one private memory pool per slot: activations saved by slot A's forward graph are returned to the
            # pool when its backward is captured, and must not be handed to slot B (both are in flight at once)
            g = GraphedMicroBatch(self._stage_fwd_static, boundary_shape(self.first_blk - 1, n, 32) if not self.is_first else None,
                                  self.rt.dtype, self.rt.device, self.is_first, self.is_last,
                                  label_shape=(n,), image_like=images[0] if self.is_first else None, pool=None)
            g.capture()
            slots.append(g)
        torch.cuda.synchronize()
        self.slots, self.slot_sizes = slots, list(sizes)
        if self.channels is not None:
            for g in slots:      # payloads travel as the dense NHWC storage of the static buffers: no staging copies
                for t in ([] if self.is_first else [g.x, g.dx]) + ([] if self.is_last else [g.out, g.dy]):
                    assert t is not None and t.permute(0, 2, 3, 1).is_contiguous(), "stage boundary tensor is not NHWC-dense"
            self.overlapped = OverlappedPipelineRunner(s, S, slots, self.channels, self.rt.device,
                                                       timing=bool(self.cfg.region_probe))

    def _fwd_graphed(self, x, i):
        slot = self.slots[i % len(self.slots)]
        return slot.run_fwd(x, self.labels_mb[i] if self.is_last else None)

    def _bwd_graphed(self, i, dout):
        return self.slots[i % len(self.slots)].run_bwd(dout)

    def step(self, images, labels):
        B = labels.shape[0]
        M = max(1, min(self.cfg.microbatches, B))
        sizes = [len(c) for c in np.array_split(np.arange(B), M)]
        self.labels_mb = list(torch.split(labels, sizes))
        self.mb_frac = [n / B for n in sizes]
        imgs = list(torch.split(images, sizes)) if self.is_first else None
        ops.step_begin(self.rt.device)
        self.flat.begin_step()
        uniform = len(set(sizes)) == 1
        if self.use_graphs and uniform and self.global_step >= 2:
            if self.slots is None or self.slot_sizes != list(sizes):
                try:
                    self._build_slots(sizes, imgs)
                except Exception as e:  # noqa: BLE001
                    print(f"[pp] graph capture failed on stage {self.rt.rank}, staying eager: {e!r}", flush=True)
                    self.use_graphs, self.slots = False, None
                    torch.cuda.synchronize()
        graphed = self.slots is not None and uniform and self.slot_sizes == list(sizes)
        if graphed:
            for p in self.flat.params:
                p._acc = True
            self.runner.fwd_fn, self.runner.bwd_fn = self._fwd_graphed, self._bwd_graphed
        else:
            self.runner.fwd_fn, self.runner.bwd_fn = self._fwd, None
        if graphed and self.overlapped is not None:
            loss, correct = self.overlapped.run(M, imgs, self.labels_mb)
        else:
            loss, correct = self.runner.run(sizes, imgs)
        ops.join_side()
        if self.dp_ar is not None:                   # hybrid DP × PP: average this stage's gradients over its replicas
            for bk in self.flat.buckets:
                self.dp_ar.allreduce_avg_(self.flat.grad[bk.start:bk.end], live=self.flat.bucket_live[bk.index])
        diff = self.opt.step(prev_grad=self.prev_grad)
        sent = self.runner.p2p.end_step() + (self.channels.end_step() if self.channels is not None else 0)
        ops.step_end()
        if self.is_last:
            self.stats.add_step(loss, correct, B, diff)
        else:
            self.stats.buf[5] += 1
        self.global_step += 1
        return sent


def train_model_parallel(rank: int, world: int, cfg: TrainConfig, device: str):
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
        mesh = DeviceMesh(world, rank, dp=cfg.dp_replicas, pp=world // cfg.dp_replicas)
    # every stage iterates the same (seeded, unshuffled) subset — layer_…:103-131; replicas of a DP × PP mesh
    # take disjoint shards of it
    sampler = None
    if mesh is not None:
        from ..data import ShardedSampler
        sampler = ShardedSampler(len(labels), mesh.dp, mesh.coord.dp, shuffle=False, seed=cfg.seed)
    loader = BatchLoader(images, labels, cfg.batch_size, rt.device, sampler=sampler)
    eng = PPEngine(cfg, rt, mesh)
    eng.runner.p2p.timing = bool(cfg.region_probe)
    stages = mesh.pp if mesh is not None else world
    rec = EpochRecorder("layer", rank, logs_dir, cfg.sample_size)
    hb = Heartbeat(cfg.heartbeat_dir, rank)
    fault = FaultInjector(cfg.inject_fault, rank)
    tag = f"pp_stage{eng.s}of{stages}"
    saver = mesh is None or mesh.coord.dp == 0          # replicas hold identical state: one copy per stage
    start_epoch = 0
    if cfg.resume:
        payload = checkpoint.load(cfg.resume, tag, eng.model, eng.opt)
        if payload is not None:
            start_epoch, eng.global_step = payload["epoch"], payload["global_step"]
    if not cfg.quiet:
        print(f"Worker {rank} is starting training...", flush=True)
    cuda = rt.device.type == "cuda"
    for epoch in range(start_epoch, cfg.epochs):
        t_epoch = time.time()
        t0 = time.time()
        if world > 1:
            dist.barrier()
        rec.total_idle += time.time() - t0
        step_times: List[float] = []
        sent_total, nsteps = 0, 0
        if cuda:
            torch.cuda.reset_peak_memory_stats(rt.device)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for bi, (x, y) in enumerate(loader):
            if cfg.max_steps and bi >= cfg.max_steps:
                break
            ts = time.time()
            rec.host.sample()
            fault.maybe_fail(eng.global_step)
            sent_total += eng.step(x, y)
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
        s = eng.stats.read_and_reset()
        epoch_time = time.time() - t_epoch
        steps = max(nsteps, 1)
        if eng.is_last:
            loss = s["loss_sum"] / steps
            acc = 100.0 * s["correct"] / max(s["seen"], 1)
            if s["grad_div_n"] > 0:
                rec.grad_divs = [s["grad_div_sum"] / s["grad_div_n"]]
        else:
            loss, acc = 0, 0
        # reference columns: its layer script's "comm_time" is the time blocked in send/recv incl. waiting for the
        # neighbour (layer_model_parallel_train.py:188-209) — here: the measured time this stage's compute stream was
        # held up behind activation / gradient arrivals (exposed p2p + pipeline bubble); the rest is compute
        p2p_total_ms = 0.0
        if cfg.region_probe:
            p2p_total_ms = eng.runner.p2p.take_time_ms() + (eng.overlapped.stall_ms() if eng.overlapped is not None else 0.0)
        comm_s = min(p2p_total_ms / 1e3, dev_s)
        rec.total_compute += dev_s - comm_s
        rec.total_comm += comm_s
        dev_s_max = allreduce_max_scalar(dev_s, rt.device)
        n_img = nsteps * cfg.batch_size if nsteps else 0
        n_img = min(n_img, len(labels)) if not cfg.max_steps else n_img
        if mesh is not None:
            n_img = min(nsteps * cfg.batch_size * mesh.dp, len(labels)) if not cfg.max_steps else n_img * mesh.dp
        ext = {"images_per_sec": n_img / dev_s_max if dev_s_max > 0 else 0, "steps": nsteps,
               "gpu_mem_MB": gpu_mem_mb(rt.device),
               "nvlink_GBps": sent_total / dev_s_max / 1e9 if dev_s_max > 0 else 0}
        if cuda:
            step_times = [dev_s / steps] * nsteps
        if cfg.region_probe:
            # time per step this stage's compute stream is blocked behind activation / gradient exchanges (includes
            # the pipeline bubble, like the reference's blocking send/recv "comm_time", layer_…:188-195,209)
            ext["p2p_ms"] = p2p_total_ms / steps
            ext["exposed_comm_ms"] = ext["p2p_ms"]
            ext["split_source"] = "device-timed stalls behind p2p arrivals"
        rec.end_epoch(epoch + 1, loss, acc, epoch_time, step_times,
                      avg_bandwidth=sent_total / steps, ext=ext)
        if eng.is_last and saver and not cfg.quiet:
            print(f"Epoch [{epoch+1}/{cfg.epochs}], Loss: {loss:.4f}, Accuracy: {acc:.2f}%, "
                  f"Time: {epoch_time:.2f}s", flush=True)
        if cfg.save_dir and saver and ((cfg.save_every and (epoch + 1) % cfg.save_every == 0) or epoch + 1 == cfg.epochs):
            checkpoint.save(cfg.save_dir, tag, eng.model, eng.opt, epoch + 1, eng.global_step)
        if world > 1:
            dist.barrier()
    if eng.is_last and saver:
        write_summary(logs_dir, f"summary_{cfg.sample_size}.json", {
            "strategy": "layer", "world_size": world, "backend": rt.backend, "dtype": str(rt.dtype),
            "microbatches": cfg.microbatches, "partition": partition_blocks(stages),
            "mesh": mesh.describe() if mesh is not None else None,
            "final": rec.rows[-1] if rec.rows else None})
    from ..launch import shutdown_distributed
    shutdown_distributed()
    return rec.frame()
