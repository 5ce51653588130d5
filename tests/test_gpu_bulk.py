"""One-shot all-reduce with cp.async.bulk pulls (csrc/comm.cu ``allreduce_bulk_kernel``, algo "bulk") on ONE GPU with
virtual ranks: same pack / flag barrier / parity protocol and the same rank-ordered fp32 sum as the register-staged
one-shot kernel, so the results must be bit-identical to it — for sizes that end inside a ring chunk, for dead-block
compaction, for back-to-back calls that wrap the 3-stage ring many times.  `late`: written after the round's GPU budget
was spent."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=4)]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nb():
    from horizonml_b200.ops import native_backend
    return native_backend


def _run(comms, streams, tensors, algo, wire_bf16, scale, live=None):
    torch.cuda.synchronize()
    for r, c in enumerate(comms):
        with torch.cuda.stream(streams[r]):
            c.allreduce(tensors[r], algo, wire_bf16, scale, live)
    torch.cuda.synchronize()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("wire_bf16", [True, False])
@pytest.mark.parametrize("n", [64, 4096 + 64, (1 << 18) + 192])
def test_bulk_allreduce_matches_oneshot_bitwise(nb, world, wire_bf16, n):
    C = nb.C
    cap = (1 << 18) + 256
    comA = [C.PeerComm(r, world, 0, cap * 4, 16) for r in range(world)]
    comB = [C.PeerComm(r, world, 0, cap * 4, 16) for r in range(world)]
    C.PeerComm.link_local(comA); C.PeerComm.link_local(comB)
    g = torch.Generator().manual_seed(n % 97 + world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(3):                                   # back-to-back calls: staging parity, ring phases
        grads = [torch.randn(n, generator=g).to(DEV) for _ in range(world)]
        a, b = [x.clone() for x in grads], [x.clone() for x in grads]
        _run(comA, streams, a, "bulk", wire_bf16, 1.0 / world)
        _run(comB, streams, b, "oneshot", wire_bf16, 1.0 / world)
        assert not any(c.error() for c in comA + comB)
        ref = torch.zeros(n, device=DEV)
        for x in grads:
            t = x * (1.0 / world)
            ref += t.bfloat16().float() if wire_bf16 else t
        for r in range(world):
            assert torch.equal(a[r], b[r]), (it, r, (a[r] - b[r]).abs().max().item())
            assert torch.equal(a[r], a[0])
            assert ((a[r] - ref).abs().max() / ref.abs().max()).item() < (1e-2 if wire_bf16 else 1e-5)


def test_bulk_allreduce_with_dead_block_compaction(nb):
    C = nb.C
    world, n = 4, 1 << 17
    g = torch.Generator().manual_seed(5)
    live = torch.nonzero(torch.rand(n // 64, generator=g) > 0.5).flatten().to(torch.int32).to(DEV)
    mask = torch.zeros(n // 64, dtype=torch.bool, device=DEV)
    mask[live.long()] = True
    emask = mask.repeat_interleave(64)
    comms = [C.PeerComm(r, world, 0, n * 2, 16) for r in range(world)]
    C.PeerComm.link_local(comms)
    streams = [torch.cuda.Stream() for _ in range(world)]
    grads = [torch.randn(n, generator=g).to(DEV) for _ in range(world)]
    work = [x.clone() for x in grads]
    _run(comms, streams, work, "bulk", True, 1.0 / world, live)
    ref = sum((x * (1.0 / world)).bfloat16().float() for x in grads)
    for r in range(world):
        assert torch.allclose(work[r][emask], ref[emask], atol=1e-6)
        assert torch.equal(work[r][~emask], grads[r][~emask])             # dead blocks never touched
