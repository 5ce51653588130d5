"""Run configuration + the CLI flag system.

The reference exposes exactly three flags per trainer (``--world_size 5 --epochs 5
--sample_size 1000``; data_parallel_train.py:293-300, layer_model_parallel_train.py:425-432,
tensor_parallel_train.py:387-394) and hard-codes everything else (batch 64, Adam 1e-3, gloo,
timeouts).  We keep those three flags and their defaults and expose the hard-coded literals
as optional flags whose defaults equal the reference's literals.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
from dataclasses import dataclass, field
from typing import Optional


@dataclass
class TrainConfig:
    # ---- reference flags (same names, same defaults) ----
    world_size: int = 5
    epochs: int = 5
    sample_size: int = 1000
    # ---- reference literals, now configurable ----
    batch_size: int = 64              # data_parallel_train.py:196
    lr: float = 1e-3                  # data_parallel_train.py:205
    num_classes: int = 10
    model: str = "resnet18"           # train.py:60-68 also allows mobilenet
    logs_dir: Optional[str] = None    # default depends on the strategy
    data_dir: str = "./data"
    # ---- new knobs ----
    strategy: str = "data"            # data | layer | tensor
    device: str = "auto"              # auto | cuda | cpu
    dtype: str = "auto"               # auto (bf16 on cuda, fp32 on cpu) | bf16 | fp32
    backend: str = "auto"             # op backend: auto | native | torch
    comm: str = "auto"                # process-group backend: auto | nccl | gloo
    allreduce: str = "auto"           # auto | oneshot | twoshot | nvls | ll | bulk (one-shot with cp.async.bulk pulls) | nccl
    bucket_mb: float = 25.0           # DDP-like bucket cap (MiB); reference uses DDP default 25
    overlap: bool = True              # overlap bucket all-reduce with backward
    zero1: bool = False               # ZeRO-1: shard Adam's moments over the data-parallel group (parallel/zero.py)
    zero1_impl: str = "auto"          # fused: one peer-memory kernel per bucket (csrc/comm.cu zero1_kernel; on CPU its
                                      # torch.distributed plumbing form) | nccl: reduce-scatter + Adam + all-gather |
                                      # auto: fused where the peer kernels run (CUDA, native backend, bf16), else nccl
    overlap_adam: bool = False        # bucket-wise Adam right behind each bucket's all-reduce (measured: no gain on B200)
    fused_adam: bool = False          # world > 1, peer all-reduce: Adam applied inside the bucket's all-reduce kernel
                                      # (measured on B200: 0.620 vs 0.601 ms/step at 2 GPUs — the update inside the comm
                                      # kernel keeps the comm stream busy longer than the separate full-bandwidth pass costs)
    bucket_layout: str = "auto"       # auto (layer groups on GPUs with the peer all-reduce, else DDP-like size caps) | layers | size
    bucket_by_live: bool = False      # with dead-tap elision, size buckets by LIVE elements using live_bucket_mb
    live_bucket_mb: float = 2.0       # (fp32 MiB of live gradient per bucket; the last bucket's collective is exposed)
    microbatches: int = 4             # pipeline micro-batches (1F1B)
    pp_overlap: bool = True           # GPUs: per-direction NCCL channels + streams, prefetched receives (no host wait between graphs)
    dp_replicas: int = 1              # hybrid DP x PP / DP x TP mesh: number of data-parallel replicas of the pipeline / TP group
    tp_conv_split: bool = True        # channel-split layer3/4 convs in tensor-parallel mode
    synthetic: bool = True            # no network in this environment: synthetic CIFAR-shaped data
    seed: int = 1234
    grad_divergence: bool = True      # reference metric (data_parallel_train.py:132-145)
    step_barrier: bool = False        # reference does a host barrier per step (:150-152)
    cuda_graph: bool = True
    skip_dead_taps: bool = True       # optimizer / all-reduce skip conv taps that only ever see padding (exact)
    watchdog_s: float = 0.0           # 0 → reference rule max(120, 120*N/1000)
    save_dir: Optional[str] = None
    resume: Optional[str] = None
    save_every: int = 0               # epochs; 0 = only at the end when save_dir is set
    profile: bool = False
    region_probe: bool = True         # fill fwd_ms/bwd_ms/allreduce_ms/optimizer_ms/exposed_comm_ms once per run
    inject_fault: Optional[str] = None   # "rank:step" → that rank aborts at that global step
    heartbeat_dir: Optional[str] = None
    quiet: bool = False
    max_steps: int = 0                # >0: stop each epoch after this many steps (smoke tests)

    def resolved_logs_dir(self) -> str:
        if self.logs_dir:
            return self.logs_dir
        return {"data": "data_parallel_logs", "layer": "model_parallel_logs",
                "tensor": "tensor_parallel_logs"}[self.strategy]

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), sort_keys=True)

    @staticmethod
    def from_json(s: str) -> "TrainConfig":
        return TrainConfig(**json.loads(s))

    def replace(self, **kw) -> "TrainConfig":
        return dataclasses.replace(self, **kw)


def _str2bool(v: str) -> bool:
    return str(v).lower() in ("1", "true", "yes", "y", "on")


def add_train_flags(p: argparse.ArgumentParser, strategy: str) -> argparse.ArgumentParser:
    """Reference flags first (identical help/defaults), then the extended set."""
    d = TrainConfig(strategy=strategy)
    p.add_argument('--world_size', type=int, default=d.world_size, help='Number of processes to spawn')
    p.add_argument('--epochs', type=int, default=d.epochs, help='Number of epochs to train')
    p.add_argument('--sample_size', type=int, default=d.sample_size, help='Number of samples to use')
    g = p.add_argument_group("horizonml_b200 extensions (defaults = reference literals)")
    g.add_argument('--batch_size', type=int, default=d.batch_size)
    g.add_argument('--lr', type=float, default=d.lr)
    g.add_argument('--model', default=d.model, choices=["resnet18", "mobilenet"])
    g.add_argument('--logs_dir', default=None)
    g.add_argument('--data_dir', default=d.data_dir)
    g.add_argument('--device', default=d.device, choices=["auto", "cuda", "cpu"])
    g.add_argument('--dtype', default=d.dtype, choices=["auto", "bf16", "fp32"])
    g.add_argument('--backend', default=d.backend, choices=["auto", "native", "torch"],
                   help="op backend: hand-written sm_100a kernels or the PyTorch oracle")
    g.add_argument('--comm', default=d.comm, choices=["auto", "nccl", "gloo"])
    g.add_argument('--allreduce', default=d.allreduce,
                   choices=["auto", "oneshot", "twoshot", "nvls", "ll", "bulk", "nccl"])
    g.add_argument('--bucket_mb', type=float, default=d.bucket_mb)
    g.add_argument('--no_overlap', dest='overlap', action='store_false')
    g.add_argument('--overlap_adam', action='store_true')
    g.add_argument('--fused_adam', action='store_true',
                   help='apply Adam inside each bucket\'s all-reduce kernel instead of a separate pass')
    g.add_argument('--bucket_layout', default=d.bucket_layout, choices=["auto", "layers", "size"])
    g.add_argument('--zero1', action='store_true', help='shard the optimizer state over the data-parallel ranks')
    g.add_argument('--zero1_impl', default=d.zero1_impl, choices=["auto", "fused", "nccl"],
                   help='ZeRO-1 implementation: one fused peer-memory kernel per bucket, or NCCL reduce-scatter/all-gather')
    g.add_argument('--bucket_by_live', action='store_true')
    g.add_argument('--live_bucket_mb', type=float, default=d.live_bucket_mb)
    g.add_argument('--microbatches', type=int, default=d.microbatches)
    g.add_argument('--no_pp_overlap', dest='pp_overlap', action='store_false',
                   help='pipeline: blocking batch_isend_irecv exchanges on the compute stream (round-1 schedule)')
    g.add_argument('--dp_replicas', type=int, default=d.dp_replicas,
                   help='hybrid mesh: replicate the pipeline / tensor-parallel group this many times (world_size = dp_replicas x stages)')
    g.add_argument('--no_tp_conv_split', dest='tp_conv_split', action='store_false')
    g.add_argument('--real_data', dest='synthetic', action='store_false',
                   help="read CIFAR-10 from --data_dir instead of synthetic data")
    g.add_argument('--seed', type=int, default=d.seed)
    g.add_argument('--no_grad_divergence', dest='grad_divergence', action='store_false')
    g.add_argument('--step_barrier', action='store_true')
    g.add_argument('--no_cuda_graph', dest='cuda_graph', action='store_false')
    g.add_argument('--no_skip_dead_taps', dest='skip_dead_taps', action='store_false')
    g.add_argument('--watchdog_s', type=float, default=d.watchdog_s)
    g.add_argument('--save_dir', default=None)
    g.add_argument('--resume', default=None)
    g.add_argument('--save_every', type=int, default=0)
    g.add_argument('--profile', action='store_true')
    g.add_argument('--no_region_probe', dest='region_probe', action='store_false')
    g.add_argument('--inject_fault', default=None, help='"rank:step" fault-injection test hook')
    g.add_argument('--heartbeat_dir', default=None)
    g.add_argument('--quiet', action='store_true')
    g.add_argument('--max_steps', type=int, default=0)
    return p


def config_from_args(args: argparse.Namespace, strategy: str) -> TrainConfig:
    names = {f.name for f in dataclasses.fields(TrainConfig)}
    kw = {k: v for k, v in vars(args).items() if k in names}
    kw["strategy"] = strategy
    return TrainConfig(**kw)
