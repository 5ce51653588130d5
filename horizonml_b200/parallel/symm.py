"""Symmetric heap: one device buffer per rank of a process group, same layout everywhere, every peer's
buffer mapped into this process — and, when the fabric supports it, the NVSwitch **multicast** mapping of the
same memory (``multimem.red`` / ``multimem.ld_reduce`` operands of the fused kernels in csrc/tp_fused.cu).

Three providers, picked in this order:

* ``symm``  — ``torch.distributed._symmetric_memory`` (CUDA VMM allocation + fabric/fd handle exchange +
  multicast object); works for sub-groups, so every tensor-parallel row of a DP × TP mesh gets its own heap;
* ``ipc``   — one ``cudaMalloc`` per rank exchanged with ``cudaIpc*`` handles (csrc/comm.cu); no multicast;
* ``local`` — ``world`` *virtual ranks* inside one process on one GPU (``SymmHeap.virtual``): the peers' heaps are
  plain tensors of the same device.  The multi-rank protocol of the kernels (counters, parity slots, pulls) is
  identical, which is what lets the single-GPU test tier cover the fused tensor-parallel kernels.

The reference has no counterpart (its only transport is gloo over TCP loopback)."""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

_DT = {"bf16": torch.bfloat16, "f32": torch.float32, "i32": torch.int32, "u8": torch.uint8}


class SymmHeap:
    def __init__(self, device, nbytes: int, group=None, kind: str = "auto"):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nbytes = (int(nbytes) + 4095) // 4096 * 4096
        self.off = 0
        self.mc_ptr = 0
        self.kind = None
        self.why_not_symm: Optional[str] = None
        self._keep = []
        if self.world == 1:
            self._local_single()
            return
        if kind in ("auto", "symm") and os.environ.get("HZ_DISABLE_SYMM", "0") != "1":
            try:
                self._setup_symm()
            except Exception as e:  # noqa: BLE001 — recorded and reported (summary / bench line), never silent
                self.why_not_symm = repr(e)
                if kind == "symm":
                    raise
        if self.kind is None:
            self._setup_ipc()
        # every rank must agree on whether the multicast path exists (a kernel using it needs all peers to)
        flag = torch.tensor([1 if self.mc_ptr else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.mc_ptr = 0
        dist.barrier(group=group)

    # ---- providers -------------------------------------------------------------------------------
    def _local_single(self):
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.ptrs = [self.local.data_ptr()]
        self.kind = "local"

    def _setup_symm(self):
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(buf, group=self.group if self.group is not None else dist.group.WORLD)
        buf.zero_()
        torch.cuda.synchronize(self.device)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
        mc = 0
        if os.environ.get("HZ_DISABLE_NVLS", "0") != "1":
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self.local, self.ptrs, self.mc_ptr, self.kind = buf, ptrs, mc, "symm"
        self._keep += [hdl]

    def _setup_ipc(self):
        from ..ops import _ext
        C = _ext.load(required=True)
        comm = C.PeerComm(self.rank, self.world, self.device.index or 0, 1024, 8, self.nbytes)
        objs = [None] * self.world
        dist.all_gather_object(objs, bytes(comm.export_handles()), group=self.group)
        comm.import_handles([bytes(o) for o in objs])
        self.ptrs = [int(comm.heap_ptr(r)) for r in range(self.world)]
        self.local = comm.heap_tensor(0, [int(comm.heap_bytes())], [1], "u8")
        self.nbytes = int(comm.heap_bytes())
        self.kind = "ipc"
        self._keep += [comm]

    @classmethod
    def virtual(cls, world: int, device, nbytes: int) -> "List[SymmHeap]":
        """``world`` virtual ranks on one device, one process (tests, tools): launch the ranks' kernels on ``world``
        different streams so that they are co-resident while they wait for each other."""
        device = torch.device(device)
        nbytes = (int(nbytes) + 4095) // 4096 * 4096
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=device) for _ in range(world)]
        heaps = []
        for r in range(world):
            h = cls.__new__(cls)
            h.device, h.group, h.world, h.rank, h.nbytes, h.off, h.mc_ptr = device, None, world, r, nbytes, 0, 0
            h.kind, h.why_not_symm, h._keep = "local", None, [bufs]
            h.local, h.ptrs = bufs[r], [b.data_ptr() for b in bufs]
            heaps.append(h)
        return heaps

    # ---- allocation ------------------------------------------------------------------------------
    def alloc(self, nbytes: int, align: int = 1024) -> int:
        """Byte offset of a fresh region (identical on every rank as long as every rank allocates in the same order)."""
        off = (self.off + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted ({self.nbytes} bytes)")
        self.off = off + int(nbytes)
        return off

    def tensor(self, off: int, sizes, strides, dtype: str = "bf16") -> torch.Tensor:
        """View of the LOCAL heap at ``off`` (element strides)."""
        dt = _DT[dtype]
        numel_span = 1 + sum((s - 1) * st for s, st in zip(sizes, strides))
        raw = self.local[off: off + numel_span * torch.empty((), dtype=dt).element_size()].view(dt)
        return raw.as_strided(list(sizes), list(strides))

    @property
    def nvls(self) -> bool:
        return bool(self.mc_ptr)

    def describe(self) -> dict:
        return {"provider": self.kind, "multicast": self.nvls, "bytes": self.nbytes, "used": self.off,
                "symm_error": self.why_not_symm}
