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


# ----------------------------------------------------------------------------------------------------------
# Overlapped 1F1B on GPUs: the pipeline as a pipeline
# ----------------------------------------------------------------------------------------------------------
class P2PChannels:
    """One NCCL communicator **and** one CUDA stream per (neighbour, direction) of a pipeline stage.

    A channel is used in one direction only (activations down *or* gradients up), in micro-batch order on both
    ends, so a send can only ever wait for its own matching receive — which the receiver posts ahead of time —
    never for traffic of the opposite direction queued on the same communicator (the reason the blocking runner
    has to issue facing pairs as one ``batch_isend_irecv`` group).  A posted ``irecv`` parks on its own stream:
    nothing else is serialised behind it.  Every rank creates every channel group in the same order
    (``pipelines`` = the global ranks of every pipeline row, stage-major)."""

    def __init__(self, rank: int, pipelines: Sequence[Sequence[int]], device):
        self.device = torch.device(device)
        self.recv_fwd = self.send_fwd = self.recv_bwd = self.send_bwd = None      # (group, peer global rank)
        for row in pipelines:
            for s in range(len(row) - 1):
                a, b = row[s], row[s + 1]
                g_down = dist.new_group([a, b])          # activations a -> b
                g_up = dist.new_group([a, b])            # gradients  b -> a
                if rank == a:
                    self.send_fwd, self.recv_bwd = (g_down, b), (g_up, b)
                if rank == b:
                    self.recv_fwd, self.send_bwd = (g_down, a), (g_up, a)
        mk = lambda ch: torch.cuda.Stream(device=self.device) if ch is not None else None   # noqa: E731
        self.streams = {k: mk(getattr(self, k)) for k in ("recv_fwd", "send_fwd", "recv_bwd", "send_bwd")}
        self.bytes_sent = 0

    def post(self, kind: str, tensor: torch.Tensor, after: Sequence[torch.cuda.Event] = ()) -> torch.cuda.Event:
        """Enqueue one transfer on the channel's own stream (after ``after``); returns the event that fires when the
        payload has arrived (recv) / left the buffer (send).  Never blocks the host or the compute stream."""
        group, peer = getattr(self, kind)
        st = self.streams[kind]
        for ev in after:
            if ev is not None:
                st.wait_event(ev)
        with torch.cuda.stream(st):
            if kind.startswith("send"):
                self.bytes_sent += tensor.numel() * tensor.element_size()
                w = dist.isend(tensor, peer, group=group)
            else:
                w = dist.irecv(tensor, peer, group=group)
            w.wait()                      # orders the channel stream behind the NCCL kernel (no host block)
            done = torch.cuda.Event()
            done.record(st)
        return done

    def end_step(self) -> int:
        b, self.bytes_sent = self.bytes_sent, 0
        return b


class OverlappedPipelineRunner:
    """1F1B over CUDA-graphed micro-batch slots with every exchange on its own channel stream.

    Per micro-batch ``i`` (slot ``i % nslots``) the host only *enqueues*:

    * ``irecv`` of the activation **straight into the slot's static input** — posted as soon as the slot is free,
      i.e. one or more micro-batches ahead of its forward (one spare slot beyond the 1F1B in-flight bound makes
      that window exist), so the transfer overlaps the stage's own compute;
    * forward graph replay behind the arrival event; ``isend`` of the slot's static output fired from an event
      recorded right after the replay; ``irecv`` of the matching gradient into the slot's static ``dy``;
    * backward graph replay behind the gradient's arrival event; ``isend`` of the slot's static ``dx``.

    There is no host ``wait()`` and no copy between graph replays; the compute stream blocks only where the data
    dependency is real.  ``stall_ms()`` reports how long it was blocked behind arrivals (exposed p2p + pipeline
    bubble — the reference's blocking send/recv "comm_time", layer_model_parallel_train.py:188-209)."""

    def __init__(self, stage: int, num_stages: int, slots: List["GraphedMicroBatch"], channels: P2PChannels, device,
                 timing: bool = True):
        self.stage, self.S = stage, num_stages
        self.first, self.last = stage == 0, stage == num_stages - 1
        self.slots, self.ch, self.device = slots, channels, torch.device(device)
        self.timing = timing
        self._stalls: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self.trace: List[Tuple[str, int]] = []
        n = len(slots)
        self._bwd_done: List[Optional[torch.cuda.Event]] = [None] * n     # slot's last backward finished (inputs reusable)
        self._sent_fwd: List[Optional[torch.cuda.Event]] = [None] * n     # slot's output has left (forward may overwrite it)
        self._sent_bwd: List[Optional[torch.cuda.Event]] = [None] * n     # slot's dx has left (backward may overwrite it)

    @staticmethod
    def plan(stage: int, num_stages: int, num_micro: int, nslots: int) -> List[Tuple[str, int]]:
        """The host-side enqueue order of ``run`` as a list of (op, micro-batch) with op in ``recv_fwd`` (posted, not
        waited for), ``F``, ``send_fwd``, ``recv_bwd``, ``B``, ``send_bwd`` — the same rules as ``run`` (a receive is
        posted as soon as the slot's previous owner's backward has been enqueued), without any device work.  For
        tests, tracing and documentation."""
        first, last = stage == 0, stage == num_stages - 1
        out: List[Tuple[str, int]] = []
        posted = b_enq = 0

        def prefetch():
            nonlocal posted
            while not first and posted < num_micro and (posted < nslots or posted - nslots < b_enq):
                out.append(("recv_fwd", posted))
                posted += 1
        prefetch()
        for a, i in one_f_one_b(stage, num_stages, num_micro):
            if a == "F":
                out.append(("F", i))
                if not last:
                    out += [("send_fwd", i), ("recv_bwd", i)]
            else:
                out.append(("B", i))
                b_enq += 1
                if not first:
                    out.append(("send_bwd", i))
            prefetch()
        return out

    @staticmethod
    def _phys(t: torch.Tensor) -> torch.Tensor:
        """The dense NHWC storage view of a channels_last activation (what travels)."""
        return t.detach().permute(0, 2, 3, 1)

    def _wait(self, cur, ev):
        if ev is None:
            return
        if self.timing:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cur)
            cur.wait_event(ev)
            b.record(cur)
            self._stalls.append((a, b))
        else:
            cur.wait_event(ev)

    def stall_ms(self) -> float:
        """Σ time the compute stream was blocked behind arrivals since the last call (one device sync)."""
        st, self._stalls = self._stalls, []
        if not st:
            return 0.0
        torch.cuda.synchronize(self.device)
        return float(sum(a.elapsed_time(b) for a, b in st))

    def run(self, M: int, first_inputs: Optional[Sequence] = None, labels: Optional[Sequence] = None):
        ns = len(self.slots)
        cur = torch.cuda.current_stream(self.device)
        acts = one_f_one_b(self.stage, self.S, M)
        recv_f: List[Optional[torch.cuda.Event]] = [None] * M
        recv_b: List[Optional[torch.cuda.Event]] = [None] * M
        posted_f = b_enq = 0
        loss_sum = correct_sum = None
        self.trace = []

        def prefetch_fwd():
            """post the activation receive of every micro-batch whose slot is free: its previous owner's backward
            has been enqueued (``_bwd_done`` then holds that backward's event)"""
            nonlocal posted_f
            while not self.first and posted_f < M and (posted_f < ns or posted_f - ns < b_enq):
                i = posted_f
                recv_f[i] = self.ch.post("recv_fwd", self._phys(self.slots[i % ns].x), after=[self._bwd_done[i % ns]])
                posted_f += 1

        prefetch_fwd()
        for a, i in acts:
            s = i % ns
            slot = self.slots[s]
            if a == "F":
                if self.first:
                    slot.x.copy_(first_inputs[i], non_blocking=True)
                else:
                    self._wait(cur, recv_f[i])
                if self.last:
                    slot.labels.copy_(labels[i], non_blocking=True)
                if self._sent_fwd[s] is not None:
                    cur.wait_event(self._sent_fwd[s])
                slot.gf.replay()
                self.trace.append(("F", i))
                if self.last:
                    loss, correct = slot.out
                    loss_sum = loss.detach().clone() if loss_sum is None else loss_sum + loss.detach()
                    correct_sum = correct.detach().clone() if correct_sum is None else correct_sum + correct.detach()
                else:
                    done = torch.cuda.Event()
                    done.record(cur)
                    self._sent_fwd[s] = self.ch.post("send_fwd", self._phys(slot.out), after=[done])
                    # the gradient of this micro-batch comes back into the slot's static dy
                    recv_b[i] = self.ch.post("recv_bwd", self._phys(slot.dy), after=[self._bwd_done[s]])
            else:
                if not self.last:
                    self._wait(cur, recv_b[i])
                if self._sent_bwd[s] is not None:
                    cur.wait_event(self._sent_bwd[s])
                slot.gb.replay()
                self.trace.append(("B", i))
                done = torch.cuda.Event()
                done.record(cur)
                self._bwd_done[s] = done
                b_enq += 1
                if not self.first:
                    self._sent_bwd[s] = self.ch.post("send_bwd", self._phys(slot.dx), after=[done])
            prefetch_fwd()
        return loss_sum, correct_sum
