"""Communication layer.

Reference: ``torch.distributed`` ProcessGroupGloo over TCP loopback, WORLD group only, blocking
calls only (K1-K12).  Here:

* ``torch.distributed`` (NCCL on GPUs, gloo on CPU) is the bootstrap, the p2p transport for the
  pipeline and the *measured baseline*;
* gradient all-reduce on GPUs goes through :class:`PeerAllReduce` — hand-written sm_100a kernels
  (csrc/comm.cu) that fuse the fp32→bf16 cast and the 1/W scale into a one-shot / two-shot
  reduction over CUDA-IPC peer buffers, plus an NVLS (``multimem``) variant over a multicast
  mapping; device-side flag barriers replace the reference's per-step host barrier (K4/K10).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> int:
    return dist.get_world_size() if dist_ready() else 1


def rank() -> int:
    return dist.get_rank() if dist_ready() else 0


# size thresholds (bytes of wire data) for algorithm selection; refined from measurements
ONESHOT_MAX_BYTES = 1 << 20
LL_MAX_ELEMS = 512 * 1024                 # csrc/comm.cu kLLMaxElems
LL_MAX_INGRESS_BYTES = int(os.environ.get("HZ_LL_MAX_INGRESS", str(6 << 20)))


def pick_allreduce_algo(numel: int, world: int, wire: str = "bf16", has_nvls: bool = False) -> str:
    """Algorithm of the fused peer all-reduce for a bucket of ``numel`` wire elements (the rule behind
    ``PeerAllReduce(algo="auto")``; measured crossover points in profiles/README.md):

    * ``ll``      — flag-in-data push, no staging pass / barrier: small buckets.  Every rank receives
      ``world x numel x 4`` bytes, so the bound is on ingress bytes as well as on the slot size;
    * ``oneshot`` — everyone reads all copies: 2 ranks, or up to 1 MiB of wire data;
    * ``nvls``    — in-switch reduction (``multimem.ld_reduce`` / ``multimem.st``) when the buffers are multicast-mapped;
    * ``twoshot`` — reduce-scatter + all-gather over peer pointers otherwise."""
    wb = numel * (2 if wire == "bf16" else 4)
    if wire == "bf16" and numel <= LL_MAX_ELEMS and world * numel * 4 <= LL_MAX_INGRESS_BYTES:
        return "ll"
    if world <= 2 or wb <= ONESHOT_MAX_BYTES:
        return "oneshot"
    return "nvls" if has_nvls else "twoshot"


class GradAllReduce:
    """In-place *averaging* all-reduce of a slice of the flat fp32 gradient buffer."""

    name = "base"

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist_ready() else 1
        self.rank = dist.get_rank(group) if dist_ready() else 0
        self.bytes_per_call: List[int] = []

    def allreduce_avg_(self, t: torch.Tensor) -> None:   # pragma: no cover - interface
        raise NotImplementedError

    def wire_bytes(self, numel: int) -> int:
        return numel * 4


class TorchDistAllReduce(GradAllReduce):
    """NCCL / gloo all-reduce (the baseline path; also the CPU plumbing path)."""

    name = "nccl"

    def allreduce_avg_(self, t: torch.Tensor, algo=None, live=None) -> None:
        if self.world == 1:
            return
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        t.div_(self.world)


class PeerAllReduce(GradAllReduce):
    """Fused cast/scale + peer-memory all-reduce (csrc/comm.cu).

    ``algo``: oneshot | twoshot | nvls | auto.  ``wire``: bf16 (default, the north-star's fused
    cast) or fp32 (bit-comparable with a fp32 NCCL all-reduce up to summation order)."""

    name = "peer"

    def __init__(self, max_numel: int, device, group=None, algo: str = "auto", wire: str = "bf16",
                 max_blocks: int = 96):
        super().__init__(group)
        from ..ops import _ext
        self.C = _ext.load(required=True)
        self.device = torch.device(device)
        self.algo, self.wire = algo, wire
        self.max_numel = int(max_numel)
        self.max_blocks = max_blocks
        wire_bytes = 2 if wire == "bf16" else 4
        self.handle = self.C.PeerComm(self.rank, self.world, self.device.index or 0,
                                      self.max_numel * wire_bytes, max_blocks)
        self.has_nvls = False
        self.early_blocks = int(os.environ.get("HZ_COMM_BLOCKS_EARLY", "24"))
        self.tail_blocks = int(os.environ.get("HZ_COMM_BLOCKS_TAIL", "0"))
        self.plain_blocks = int(os.environ.get("HZ_COMM_BLOCKS", "32"))   # measured: 32 CTAs 0.601, 96 CTAs 0.608 ms/step (2 GPUs)
        self.nvls_error: Optional[str] = None
        if self.world > 1:
            # 64-byte cudaIpcMemHandle of this rank's region -> everyone (opaque host bytes)
            objs: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(objs, bytes(self.handle.export_handles()), group=group)
            self.handle.import_handles([bytes(o) for o in objs])
            if algo in ("nvls", "auto"):
                self.has_nvls = self._try_setup_nvls(group)
            dist.barrier(group=group)
        if algo == "nvls" and not self.has_nvls and self.world > 1:
            raise RuntimeError("NVLS multicast is not available on this system")

    def _try_setup_nvls(self, group) -> bool:
        """Multicast (NVLS) mapping via torch symmetric memory; kernels are ours."""
        if os.environ.get("HZ_DISABLE_NVLS", "0") == "1":
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = int(self.handle.symm_bytes())     # stage[2] + out[2] + latency-protocol slots
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            hdl = symm.rendezvous(buf, group=group if group is not None else dist.group.WORLD)
            buf.zero_()                                 # LL flags start at 0
            torch.cuda.synchronize(self.device)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            ok = torch.tensor([1 if mc else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                return False
            self._symm_buf, self._symm_hdl = buf, hdl
            self.handle.set_multicast(mc, buf.data_ptr(), nbytes)
            return True
        except Exception as e:  # noqa: BLE001 — recorded (bench line / summaries show why NVLS is off), never silent
            self.nvls_error = repr(e)
            if os.environ.get("HZ_DEBUG"):
                print(f"[comm] NVLS setup failed: {e!r}")
            return False

    def pick(self, numel: int) -> str:
        if self.algo != "auto":
            return self.algo
        return pick_allreduce_algo(numel, self.world, self.wire, self.has_nvls)

    def wire_bytes(self, numel: int) -> int:
        return numel * (2 if self.wire == "bf16" else 4)

    def describe(self) -> dict:
        return {"kind": "peer", "wire": self.wire, "algo": self.algo, "nvls": self.has_nvls, "nvls_error": self.nvls_error,
                "max_blocks": self.max_blocks}

    def allreduce_avg_(self, t: torch.Tensor, algo: Optional[str] = None, live: Optional[torch.Tensor] = None,
                       background: bool = False) -> None:
        """``live``: int32 indices (relative to ``t``) of the 64-element blocks to reduce — the rest of ``t`` is
        known to be identically zero on every rank (dead conv taps) and never touches the wire.  ``background``: the
        call runs in the shadow of other kernels (a gradient bucket during backward): cap its grid (HZ_COMM_BLOCKS)."""
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_numel
        n_wire = t.numel() if live is None else live.numel() * 64
        a = algo or self.pick(n_wire)
        # "ll" buckets are the latency-critical ones (small / last); the staged algorithms run in the shadow of the
        # remaining backward kernels and get a CTA cap so that they do not crowd them
        self.handle.set_block_cap(self.plain_blocks if (background and a != "ll") else 0)
        self.handle.allreduce(t, a, self.wire == "bf16", 1.0 / self.world, live)

    def allreduce_adam_(self, t: torch.Tensor, master, m, v, shadow, prev, diff_out, step_t, lr, b1, b2, eps,
                        bump: bool, algo: Optional[str] = None, live: Optional[torch.Tensor] = None) -> str:
        """Averaging all-reduce of gradient bucket ``t`` **and** the Adam update of the bucket's parameters in ONE
        kernel (csrc/comm.cu ``AdamFuse``): the reduced values feed the update directly, the gradient slice is
        cleared.  All tensors are the bucket's slices of their flat buffers.  Returns the algorithm used."""
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_numel
        n_wire = t.numel() if live is None else live.numel() * 64
        a = algo or self.pick(n_wire)
        # a bucket reduced in the shadow of the remaining backward needs few CTAs (it has ~100 us of slack and must
        # not take the SMs from the latency-bound backward kernels); the last bucket is on the critical path
        self.handle.set_block_cap(self.tail_blocks if bump else self.early_blocks)
        self.handle.allreduce_adam(t, a, self.wire == "bf16", 1.0 / self.world, live, master, m, v, shadow, prev,
                                   diff_out, step_t, float(lr), float(b1), float(b2), float(eps), bool(bump))
        return a

    def barrier(self) -> None:
        self.handle.barrier(None)

    def check_error(self) -> None:
        """Raise if a kernel of this communicator ever gave up waiting for a peer (it also traps, which surfaces as a
        CUDA error at the next synchronisation; this is the explicit check trainers run once per epoch)."""
        if self.handle.error():
            raise RuntimeError("peer all-reduce: a rank never arrived at a flag barrier (HZ_COMM_TIMEOUT_S to extend)")


def make_grad_allreduce(kind: str, max_numel: int, device, group=None, wire: str = "bf16") -> GradAllReduce:
    dev = torch.device(device)
    if kind == "auto":
        kind = "peer-auto" if dev.type == "cuda" else "nccl"
        try:
            from ..ops import _ext
            if dev.type == "cuda" and _ext.load(required=False) is None:
                kind = "nccl"
        except Exception:
            kind = "nccl"
    if kind == "nccl" or dev.type != "cuda":
        return TorchDistAllReduce(group)
    algo = {"peer-auto": "auto"}.get(kind, kind)
    return PeerAllReduce(max_numel, dev, group, algo=algo, wire=wire)
