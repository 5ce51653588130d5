"""Layer (pipeline) parallelism: real 1F1B.

Reference (layer_model_parallel_train.py:172-273): a *forward-only* blocking chain — each stage
``recv``s a 32-byte shape header + the activation into a fresh leaf tensor, runs its segment and
``send``s on; only the last stage has a loss/optimizer, no gradient ever travels upstream (Q1).

Here every stage trains: activations go down and gradients come back over ``torch.distributed``
p2p (NCCL on GPUs — the north-star keeps PP traffic on NCCL p2p; gloo on CPU), shapes are static
(no header), payloads are in the compute dtype (bf16), and the batch is cut into micro-batches
scheduled **1F1B** so a stage's p2p for micro-batch *i* overlaps its neighbours' compute.  Sends and
receives that face each other are issued as one ``batch_isend_irecv`` group (one NCCL group
kernel), which is what makes the steady state deadlock-free.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def one_f_one_b(stage: int, num_stages: int, num_micro: int) -> List[Tuple[str, int]]:
    """The 1F1B action list for ``stage``: [('F', i) | ('B', i)].  Used for tests / tracing;
    ``PipelineRunner`` executes the same order."""
    warm = min(num_stages - stage - 1, num_micro)
    acts: List[Tuple[str, int]] = [("F", i) for i in range(warm)]
    f, b = warm, 0
    for _ in range(num_micro - warm):
        acts.append(("F", f)); f += 1
        acts.append(("B", b)); b += 1
    while b < num_micro:
        acts.append(("B", b)); b += 1
    return acts


class P2P:
    """Static-shape activation/gradient exchange with the neighbouring stages."""

    def __init__(self, stage: int, num_stages: int, group=None, peers=None):
        """``peers`` = (global rank of the previous stage, of the next stage) when the pipeline is one row of a
        process mesh (parallel/mesh.py); default: stage index == global rank."""
        self.stage, self.S, self.group = stage, num_stages, group
        self.prev = stage - 1 if stage > 0 else None
        self.next = stage + 1 if stage < num_stages - 1 else None
        if peers is not None:
            self.prev, self.next = peers
        self.bytes_sent = 0
        self._keep: List = []
        self.timing = False          # trainers switch this on: per-exchange spans, resolved once per epoch
        self._spans: List = []

    def _run(self, ops_: List[dist.P2POp]):
        if not ops_:
            return
        t0 = self._clock() if self.timing else None
        reqs = dist.batch_isend_irecv(ops_)
        for r in reqs:
            r.wait()
        if t0 is not None:
            self._spans.append((t0, self._clock()))

    @staticmethod
    def _clock():
        """A CUDA event on the compute stream (which ``req.wait()`` blocks behind the transfer) or the host clock:
        the span is the time this stage's compute stream is held up by the exchange — exposed p2p time."""
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        import time
        return time.perf_counter()

    def take_time_ms(self) -> float:
        """Σ of the exchange spans since the last call (synchronises the device once: call at epoch end)."""
        spans, self._spans = self._spans, []
        if not spans:
            return 0.0
        if not isinstance(spans[0][0], float):
            torch.cuda.synchronize()
            return float(sum(a.elapsed_time(b) for a, b in spans))
        return float(sum(b - a for a, b in spans) * 1e3)

    def _send(self, t, peer):
        self.bytes_sent += t.numel() * t.element_size()
        self._keep.append(t)
        return dist.P2POp(dist.isend, t, peer, self.group)

    def _recv(self, t, peer):
        return dist.P2POp(dist.irecv, t, peer, self.group)

    def exchange(self, send_fwd=None, send_bwd=None, recv_fwd_like=None, recv_bwd_like=None):
        """Issue up to four p2p ops as ONE group. ``*_like`` = (shape, dtype, device) to receive."""
        ops_: List[dist.P2POp] = []
        got_f = got_b = None
        # payloads travel in their physical NHWC order (a dense contiguous tensor for c10d)
        def phys(t):
            return t.detach().contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)

        def alloc(like):
            (n, c, h, w), dtype, dev = like
            return torch.empty((n, h, w, c), dtype=dtype, device=dev)

        if send_fwd is not None and self.next is not None:
            ops_.append(self._send(phys(send_fwd), self.next))
        if send_bwd is not None and self.prev is not None:
            ops_.append(self._send(phys(send_bwd), self.prev))
        if recv_fwd_like is not None and self.prev is not None:
            got_f = alloc(recv_fwd_like)
            ops_.append(self._recv(got_f, self.prev))
        if recv_bwd_like is not None and self.next is not None:
            got_b = alloc(recv_bwd_like)
            ops_.append(self._recv(got_b, self.next))
        self._run(ops_)
        return (got_f.permute(0, 3, 1, 2) if got_f is not None else None,
                got_b.permute(0, 3, 1, 2) if got_b is not None else None)

    def end_step(self):
        self._keep.clear()
        b, self.bytes_sent = self.bytes_sent, 0
        return b


class PipelineRunner:
    """Executes one optimizer step's worth of micro-batches on this stage with the 1F1B order.

    ``fwd_fn(x_or_images, mb_index) -> out`` (for the last stage: returns (loss, correct));
    stage boundaries are tensors of ``in_shape(mb_size)`` / ``out_shape(mb_size)``."""

    def __init__(self, stage: int, num_stages: int, fwd_fn: Callable, in_shape: Callable,
                 out_shape: Callable, dtype, device, group=None, bwd_fn: Optional[Callable] = None, peers=None):
        self.bwd_fn = bwd_fn      # optional (i, dout) -> dx override (CUDA-graphed micro-batches)
        self.stage, self.S = stage, num_stages
        self.first, self.last = stage == 0, stage == num_stages - 1
        self.fwd_fn, self.in_shape, self.out_shape = fwd_fn, in_shape, out_shape
        self.dtype, self.device = dtype, device
        self.p2p = P2P(stage, num_stages, group, peers)
        self.trace: List[Tuple[str, int]] = []

    def _like_in(self, n):
        return None if self.first else (self.in_shape(n), self.dtype, self.device)

    def _like_out(self, n):
        return None if self.last else (self.out_shape(n), self.dtype, self.device)

    def run(self, mb_sizes: Sequence[int], first_inputs: Optional[Sequence] = None):
        """Returns (sum of micro losses, sum of correct) on the last stage, else (None, None)."""
        M = len(mb_sizes)
        warm = min(self.S - self.stage - 1, M)
        remaining = M - warm
        inputs: List = []
        outputs: List = []
        self.trace = []
        loss_sum = correct_sum = None
        f_idx = b_idx = 0

        def do_fwd(x, i):
            nonlocal loss_sum, correct_sum
            if not self.first:
                x.requires_grad_(True)
            out = self.fwd_fn(x if not self.first else first_inputs[i], i)
            inputs.append(x)
            outputs.append(out)
            self.trace.append(("F", i))
            if self.last:
                loss, correct = out
                # clone: with graphed micro-batches `loss` / `correct` are static buffers that the next replay overwrites
                loss_sum = loss.detach().clone() if loss_sum is None else loss_sum + loss.detach()
                correct_sum = correct.detach().clone() if correct_sum is None else correct_sum + correct.detach()
            return out

        def do_bwd(dout, i):
            x, out = inputs[i], outputs[i]
            if self.bwd_fn is not None:
                dx = self.bwd_fn(i, dout)
                self.trace.append(("B", i))
                inputs[i] = outputs[i] = None
                return dx
            if self.last:
                from ..ops import backward as _bw
                _bw(out[0])
            else:
                torch.autograd.backward(out, dout)
            self.trace.append(("B", i))
            inputs[i] = outputs[i] = None
            return None if self.first else x.grad

        # ---- warm-up forwards
        for _ in range(warm):
            x, _ = self.p2p.exchange(recv_fwd_like=self._like_in(mb_sizes[f_idx]))
            out = do_fwd(x, f_idx)
            self.p2p.exchange(send_fwd=None if self.last else out)
            f_idx += 1
        x = None
        if remaining > 0:
            x, _ = self.p2p.exchange(recv_fwd_like=self._like_in(mb_sizes[f_idx]))
        # ---- steady state 1F1B
        for k in range(remaining):
            out = do_fwd(x, f_idx)
            f_idx += 1
            _, dout = self.p2p.exchange(send_fwd=None if self.last else out,
                                        recv_bwd_like=self._like_out(mb_sizes[b_idx]))
            dx = do_bwd(dout, b_idx)
            b_idx += 1
            if k == remaining - 1:
                self.p2p.exchange(send_bwd=dx)
                x = None
            else:
                x, _ = self.p2p.exchange(send_bwd=dx, recv_fwd_like=self._like_in(mb_sizes[f_idx]))
        # ---- cool-down backwards
        while b_idx < M:
            _, dout = self.p2p.exchange(recv_bwd_like=self._like_out(mb_sizes[b_idx]))
            dx = do_bwd(dout, b_idx)
            self.p2p.exchange(send_bwd=dx)
            b_idx += 1
        return loss_sum, correct_sum


class GraphedMicroBatch:
    """One in-flight micro-batch slot of a pipeline stage captured as two CUDA graphs (forward, backward)
    over static buffers — the 1F1B loop then replays graphs instead of launching ~130 kernels per
    micro-batch from Python.  Stage ``s`` of ``S`` needs ``min(S - s, M)`` slots (1F1B in-flight bound)."""

    def __init__(self, fwd, in_shape, in_dtype, device, first: bool, last: bool, label_shape=None,
                 image_like=None, pool=None):
        self.first, self.last = first, last
        dev = device
        if first:
            self.x = torch.zeros_like(image_like)
        else:
            n, c, h, w = in_shape
            self.x = torch.zeros((n, h, w, c), dtype=in_dtype, device=dev).permute(0, 3, 1, 2).requires_grad_(True)
        self.labels = torch.zeros(label_shape, dtype=torch.long, device=dev) if last else None
        self.fwd = fwd
        self.gf, self.gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self.pool = pool

    def capture(self):
        from .. import ops
        with torch.cuda.graph(self.gf, pool=self.pool):
            self.out = self.fwd(self.x, self.labels)
        self.pool = self.gf.pool()
        if self.last:
            loss, correct = self.out
            self.dy = None
            with torch.cuda.graph(self.gb, pool=self.pool):
                ops.backward(loss)
                ops.join_side()
        else:
            self.dy = torch.zeros_like(self.out)
            with torch.cuda.graph(self.gb, pool=self.pool):
                torch.autograd.backward(self.out, self.dy)
                ops.join_side()
        self.dx = None if self.first else self.x.grad
        return self.pool

    def run_fwd(self, x, labels=None):
        self.x.detach().copy_(x, non_blocking=True)
        if self.last:
            self.labels.copy_(labels, non_blocking=True)
        self.gf.replay()
        return self.out

    def run_bwd(self, dout):
        if not self.last:
            self.dy.copy_(dout, non_blocking=True)
        self.gb.replay()
        return self.dx
