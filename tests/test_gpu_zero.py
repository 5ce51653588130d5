"""ZeRO-1 as one peer-memory kernel per bucket (csrc/comm.cu ``zero1_kernel``) on ONE GPU with virtual ranks: reduce ->
Adam on the owned shard -> bf16 parameter push must equal the all-reduce kernel followed by the replicated Adam kernel
(same fp32 values feed the same arithmetic), with moments that exist only as 1/W shards.

Written after the round's GPU budget was spent (`late`: collected after the hardware-verified tests); the sharding /
state layout is also exercised on CPU over gloo (tests/test_dist_cpu.py::test_zero1_per_bucket_fused_form_equivalence_gloo)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=3)]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nb():
    from horizonml_b200.ops import native_backend
    return native_backend


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("cap", [0, 1])       # grid of 2 blocks / capped to 1
def test_zero1_fused_kernel_virtual_ranks(nb, world, cap):
    C = nb.C
    n = 1 << 16
    g = torch.Generator().manual_seed(17)
    live_all = torch.nonzero(torch.rand(n // 64, generator=g) > 0.4).flatten()
    halves = []
    for lo, hi in ((0, n // 2), (n // 2, n)):                      # two "buckets"; the second is the last of the step
        lb = (live_all[(live_all >= lo // 64) & (live_all < hi // 64)] - lo // 64).to(torch.int32).to(DEV)
        idx = (lo + lb.long()[:, None] * 64 + torch.arange(64, device=DEV)[None, :]).reshape(-1)      # wire order
        halves.append((lo, hi, lb, idx))
    comA = [C.PeerComm(r, world, 0, n * 2, 16) for r in range(world)]
    comB = [C.PeerComm(r, world, 0, n * 2, 16) for r in range(world)]
    C.PeerComm.link_local(comA); C.PeerComm.link_local(comB)
    shard = [int(comA[0].zero1_shard(idx.numel())) for _, _, _, idx in halves]
    p0 = torch.randn(n, generator=torch.Generator().manual_seed(1)).to(DEV)

    def zeros(k, dt=torch.float32):
        return torch.zeros(k, device=DEV, dtype=dt)
    A = {"p": [p0.clone() for _ in range(world)], "sh": [p0.bfloat16() for _ in range(world)],
         "m": [[zeros(s) for s in shard] for _ in range(world)], "v": [[zeros(s) for s in shard] for _ in range(world)],
         "prev": [[zeros(s) for s in shard] for _ in range(world)], "diff": [zeros(()) for _ in range(world)],
         "step": [zeros(1) for _ in range(world)]}
    B = {"p": [p0.clone() for _ in range(world)], "sh": [p0.bfloat16() for _ in range(world)],
         "m": [zeros(n) for _ in range(world)], "v": [zeros(n) for _ in range(world)],
         "prev": [zeros(n) for _ in range(world)], "diff": [zeros(()) for _ in range(world)],
         "step": [zeros(1) for _ in range(world)]}
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(3):
        grads = [torch.randn(n, generator=g).to(DEV) for _ in range(world)]
        gA, gB = [x.clone() for x in grads], [x.clone() for x in grads]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                for bi, (lo, hi, lb, idx) in enumerate(halves):
                    comA[r].set_block_cap(cap)
                    comA[r].zero1_step(gA[r][lo:hi], 1.0 / world, lb, A["p"][r][lo:hi], A["m"][r][bi], A["v"][r][bi],
                                       A["sh"][r][lo:hi], A["prev"][r][bi], A["diff"][r], A["step"][r],
                                       1e-3, 0.9, 0.999, 1e-8, bi == 1)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                for bi, (lo, hi, lb, idx) in enumerate(halves):
                    comB[r].set_block_cap(cap)
                    # one-shot: the W bf16 copies are summed in fp32 and stored unrounded — the values Adam sees in
                    # the fused kernel (two-shot would round the sum to bf16 for its second hop)
                    comB[r].allreduce(gB[r][lo:hi], "oneshot", True, 1.0 / world, lb)
        torch.cuda.synchronize()
        for r in range(world):
            for bi, (lo, hi, lb, idx) in enumerate(halves):
                C.adam_step(B["p"][r][lo:hi], gB[r][lo:hi], B["m"][r][lo:hi], B["v"][r][lo:hi], B["sh"][r][lo:hi], B["step"][r],
                            1e-3, 0.9, 0.999, 1e-8, 1.0, B["prev"][r][lo:hi], B["diff"][r], True, lb, bi == 0, 0)
        torch.cuda.synchronize()
        assert not any(c.error() for c in comA + comB)
        for r in range(world):
            assert A["step"][r].item() == it + 1 == B["step"][r].item()
            # every rank reads the same bf16 parameters, and they are the replicated optimizer's
            assert torch.equal(A["sh"][r], A["sh"][0]), "replicas diverged"
            assert torch.allclose(A["sh"][r].float(), B["sh"][r].float(), rtol=1e-5, atol=1e-2)
            assert torch.equal(gA[r], gB[r])                       # live blocks cleared, dead ones untouched
            assert abs(A["diff"][r].item() - B["diff"][r].item()) <= 1e-4 * abs(B["diff"][r].item()) + 1e-6
            A["diff"][r].zero_()
            for bi, (lo, hi, lb, idx) in enumerate(halves):
                s = shard[bi]
                w_lo = min(r * s, idx.numel())
                own = idx[w_lo:min(w_lo + s, idx.numel())]         # flat offsets of the slice rank r owns
                k = own.numel()
                for name, full in (("m", B["m"][r]), ("v", B["v"][r]), ("prev", B["prev"][r])):
                    assert torch.allclose(A[name][r][bi][:k], full[own], rtol=1e-5, atol=1e-6), (name, r, bi, it)
                assert torch.allclose(A["p"][r][own], B["p"][r][own], rtol=1e-5, atol=1e-6)      # authoritative master slice
                assert (A["m"][r][bi][k:] == 0).all() and (A["v"][r][bi][k:] == 0).all()      # shard padding untouched


def test_fused_sharded_adam_class_on_virtual_ranks_state_layout(nb):
    """FusedShardedAdam's index bookkeeping (wire order, owned ranges) matches what the kernel uses: after one native
    step per virtual rank, the master elements at ``idx[lo:lo+cnt]`` are exactly the ones that moved."""
    C = nb.C
    from horizonml_b200.models.flat import Bucket
    from horizonml_b200.parallel.zero import FusedShardedAdam
    world, n = 2, 1 << 14

    class _Flat:                                                   # the slice of FlatParams the optimizer touches
        pass

    class _Ar:
        pass
    comms = [C.PeerComm(r, world, 0, n * 2, 16) for r in range(world)]
    C.PeerComm.link_local(comms)
    g = torch.Generator().manual_seed(2)
    live = torch.nonzero(torch.rand(n // 64, generator=g) > 0.3).flatten().to(torch.int32).to(DEV)
    opts, flats = [], []
    for r in range(world):
        f = _Flat()
        f.device, f.total = torch.device(DEV), n
        f.master = torch.ones(n, device=DEV)
        f.grad = torch.ones(n, device=DEV) * (r + 1)
        f.shadow = f.master.bfloat16()
        f.buckets, f.bucket_live = [Bucket(0, 0, n, ["w"])], [live]
        ar = _Ar()
        ar.world, ar.rank, ar.handle = world, r, comms[r]
        flats.append(f)
        opts.append(FusedShardedAdam(f, ar, lr=1e-2, with_prev=True))
        assert opts[-1].native
        f.grad.fill_(float(r + 1))                                 # (the constructor clears the gradient buffer)
    streams = [torch.cuda.Stream() for _ in range(world)]
    diffs = [torch.zeros((), device=DEV) for _ in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            opts[r].step_bucket(0, True, diffs[r])
    torch.cuda.synchronize()
    for r in range(world):
        lo, cnt = opts[r].own[0]
        own = opts[r].idx[0][lo:lo + cnt]
        moved = torch.nonzero(flats[r].master != 1.0).flatten()
        assert torch.equal(moved, own.sort().values), (moved.numel(), own.numel())
        assert (flats[r].grad[opts[r].idx[0]] == 0).all() and opts[r].step_t.item() == 1.0
        assert torch.equal(flats[r].shadow, flats[0].shadow)
        assert (flats[r].shadow[opts[r].idx[0]].float() < 1.0).all()     # every live parameter was updated everywhere
