"""Executable models of the peer-memory protocols (SURVEY §5.2 "race detection"): the buffer-reuse arguments of
csrc/comm.cu and csrc/tp_fused.cu checked under randomly interleaved schedules on the CPU.

The kernels synchronise GPUs with nothing but words in memory, so their safety arguments are about *which call may
still be reading a buffer when a later call overwrites it*.  compute-sanitizer cannot see across processes, and a
hazard window of a few hundred nanoseconds does not show up in functional tests; a small model that explores
arbitrary rank / block skew does.  Two models:

* ``simulate_ll``      — the flag-in-data ("LL") protocol of the gradient all-reduce's last bucket, the fused
  tensor-parallel GEMM+all-reduce, the TP head and the small bf16 all-reduce: every call pushes {data, epoch} words into
  slot ``[parity][me]`` of every rank and polls its own slots.  Claim: with two parities a word is never overwritten
  while a reader of an earlier call still waits for it, and no schedule deadlocks; with ONE buffer it is overwritten —
  the model finds that.
* ``simulate_staged``  — the staged (pack → block barrier → reduce) one-shot all-reduce with grids of different size
  back to back.  Claim: taking the buffer parity from ONE per-communicator call counter is
  safe; the round-1 scheme (a counter per block index) lets a block that did not exist in the previous, smaller call
  reuse that call's parity and overwrite a region a slow peer is still reading — the model reproduces that hazard.

Both return the number of schedules (out of ``trials``) in which a violation was observed.  Stores of one phase land in
any order (GPU stores are unordered without fences); the kernels of one rank run in stream order; blocks of one kernel
run concurrently; the scheduler is uniformly random over everything that can make a step."""
from __future__ import annotations

import random
from typing import Dict, List, Sequence


# ----------------------------------------------------------------------------------------------------------
# flag-in-data protocol
# ----------------------------------------------------------------------------------------------------------
def simulate_ll(world: int = 3, calls: int = 6, words: int = 3, parities: int = 2, trials: int = 200, seed: int = 0,
                max_steps: int = 200_000) -> int:
    """Violation = a rank polling for epoch ``e`` finds a NEWER epoch in the word (its data was overwritten before it
    was read — the real kernel would spin until its timeout), reads a payload that does not belong to ``e``, or the
    system stops making progress."""
    rng = random.Random(seed)
    bad = 0
    for _ in range(trials):
        slots = [[[[(0, 0)] * words for _ in range(world)] for _ in range(parities)] for _ in range(world)]   # [dst][par][src][w]
        state = [{"call": 1, "stores": [(d, w) for d in range(world) for w in range(words)], "reads": None} for _ in range(world)]
        live = list(range(world))
        violated, steps = False, 0
        while live and not violated:
            steps += 1
            if steps > max_steps:
                violated = True                              # nobody can finish: deadlock
                break
            r = rng.choice(live)
            st = state[r]
            e = st["call"]
            par = e % parities
            if st["stores"]:
                dst, w = st["stores"].pop(rng.randrange(len(st["stores"])))
                slots[dst][par][r][w] = (e, e)               # {epoch, payload} travel in one atomic word
                if not st["stores"]:
                    st["reads"] = [(s, w2) for s in range(world) for w2 in range(words)]
                continue
            s, w = st["reads"][0]
            ep, payload = slots[r][par][s][w]
            if ep > e or (ep == e and payload != e):
                violated = True
            elif ep == e:                                    # the poll succeeds (else: keep spinning)
                st["reads"].pop(0)
                if not st["reads"]:
                    st["call"] += 1
                    if st["call"] > calls:
                        live.remove(r)
                    else:
                        st["stores"] = [(d, w2) for d in range(world) for w2 in range(words)]
        bad += 1 if violated else 0
    return bad


# ----------------------------------------------------------------------------------------------------------
# staged all-reduce with varying grid sizes
# ----------------------------------------------------------------------------------------------------------
def _make_blocks(grid: int, region: int, parity_of) -> List[Dict]:
    per = (region + grid - 1) // grid
    blocks = []
    for b in range(grid):
        lo, hi = min(b * per, region), min(b * per + per, region)
        blocks.append({"id": b, "range": list(range(lo, hi)), "parity": parity_of(b), "phase": "pack",
                       "todo": list(range(lo, hi))})
    return blocks


def simulate_staged(grids: Sequence[int], world: int = 2, region: int = 12, per_block_parity: bool = False,
                    trials: int = 300, seed: int = 0) -> int:
    """One-shot flavour: in call ``k`` (grid ``grids[k]``) block ``b`` of every rank packs its share of ``region`` words
    into the rank's ``stage[parity]``, meets block ``b`` of every peer at a barrier, then reads the same share from every
    peer's ``stage[parity]`` — no trailing barrier.  Violation = a block reads a word that was not written by the call
    it belongs to."""
    rng = random.Random(seed)
    bad = 0
    ncalls, gmax = len(grids), max(grids)
    for _ in range(trials):
        stage = [[[-1] * region for _ in range(2)] for _ in range(world)]            # stage[rank][parity][word] = call id
        calls_done = [0] * world                                                     # ONE counter per communicator
        block_calls = [[0] * gmax for _ in range(world)]                             # round 1: one counter per block index
        arrived = [[[False] * world for _ in range(gmax)] for _ in range(ncalls)]    # barrier flags [call][block][rank]
        cur = [0] * world

        def parity_fn(r):
            return (lambda b: block_calls[r][b] & 1) if per_block_parity else (lambda b: calls_done[r] & 1)

        blocks = [_make_blocks(grids[0], region, parity_fn(r)) for r in range(world)]
        violated = False
        while not violated and any(c < ncalls for c in cur):
            r = rng.choice([q for q in range(world) if cur[q] < ncalls])
            k = cur[r]
            b = rng.choice([x for x in blocks[r] if x["phase"] != "done"])
            if b["phase"] == "pack":
                if b["todo"]:
                    w = b["todo"].pop(rng.randrange(len(b["todo"])))
                    stage[r][b["parity"]][w] = k
                if not b["todo"]:
                    arrived[k][b["id"]][r] = True                                    # flag published after the block's stores
                    b["phase"] = "barrier"
            elif b["phase"] == "barrier":
                if all(arrived[k][b["id"]]):
                    b["phase"], b["todo"] = "read", [(s, w) for s in range(world) for w in b["range"]]
                    if not b["todo"]:
                        b["phase"] = "done"
                        block_calls[r][b["id"]] += 1
            else:
                s, w = b["todo"].pop(rng.randrange(len(b["todo"])))
                if stage[s][b["parity"]][w] != k:
                    violated = True
                if not b["todo"]:
                    b["phase"] = "done"
                    block_calls[r][b["id"]] += 1
            if all(x["phase"] == "done" for x in blocks[r]):                         # kernel complete: the next call may start
                calls_done[r] += 1
                cur[r] += 1
                if cur[r] < ncalls:
                    blocks[r] = _make_blocks(grids[cur[r]], region, parity_fn(r))
        bad += 1 if violated else 0
    return bad


# ----------------------------------------------------------------------------------------------------------
# persistent convolution kernel: operand ring + two TMEM accumulators (csrc/conv_gemm.cu igemm_persist_kernel)
# ----------------------------------------------------------------------------------------------------------
class _MBar:
    """mbarrier with phase parity: ``try_wait(P)`` is true iff the phase of parity P has completed, i.e. iff the phase
    currently in progress has the other parity (a fresh barrier: phase 0 in progress, so parity 1 passes at once — the
    "wait on parity^1" idiom of empty barriers).  A waiter that falls two phases behind would alias; the protocol has
    to make that impossible, which is what the model checks."""

    def __init__(self, count: int):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self) -> None:
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def try_wait(self, parity: int) -> bool:
        return (self.phase & 1) != parity


def simulate_persistent_pipeline(tile_k: Sequence[int], stages: int = 6, epi_warps: int = 4, trials: int = 200,
                                 seed: int = 0, broken: str = "") -> int:
    """One persistent CTA working through tiles with ``tile_k[i]`` k-iterations each (0 = a class without taps: no
    accumulator is used, the epilogue stores zeros).  Three agents with exactly the kernel's index / parity formulas:

    * producer : ring slot ``it % stages``, waits ``empty[s]`` on parity ``((it // stages) & 1) ^ 1``, fills, ``full[s]``
      completes (TMA transaction);
    * MMA      : per tile with k > 0 takes accumulator ``acc_it & 1``, waits ``tmem_empty[acc]`` on parity
      ``((acc_it >> 1) & 1) ^ 1``, then per k waits ``full[s]`` on ``(it // stages) & 1``, consumes, commits
      ``empty[s]``; after the last k commits ``tmem_full[acc]`` (tcgen05.commit arrives only when the MMAs are done);
    * epilogue : ``epi_warps`` warps, each waits ``tmem_full[acc]`` on ``(acc_it >> 1) & 1``, reads the accumulator,
      arrives on ``tmem_empty[acc]`` (count = epi_warps).

    Violation = a slot filled before its previous content was consumed, an MMA reading a slot that does not hold the
    (tile, k) it expects, an accumulator overwritten before every epilogue warp has read it, an epilogue warp reading an
    accumulator that does not hold its tile, or no agent able to move before all tiles are done (deadlock).
    ``broken``: "acc_parity" drops the ^1 of the MMA's tmem_empty wait, "one_acc" uses a single accumulator without
    waiting — the model must catch both (sanity of the model itself).  Returns the number of violating schedules."""
    rng = random.Random(seed)
    bad = 0
    for _ in range(trials):
        full = [_MBar(1) for _ in range(stages)]
        empty = [_MBar(1) for _ in range(stages)]
        tfull = [_MBar(1) for _ in range(2)]
        tempty = [_MBar(epi_warps) for _ in range(2)]
        slot = [None] * stages                   # content tag (tile, k) or None = consumed / never filled
        acc = [None, None]                       # {"tile": t, "k": n accumulated, "readers": set of warps}
        violated = [False]

        def producer():
            it = 0
            for t, kt in enumerate(tile_k):
                for k in range(kt):
                    s = it % stages
                    while not empty[s].try_wait(((it // stages) & 1) ^ 1):
                        yield False
                    if slot[s] is not None:
                        violated[0] = True       # overwrote operands the MMA has not consumed
                    slot[s] = (t, k)
                    full[s].arrive()             # TMA completes the transaction bytes
                    it += 1
                    yield True

        def mma():
            it = acc_it = 0
            for t, kt in enumerate(tile_k):
                if kt == 0:
                    continue
                a = 0 if broken == "one_acc" else (acc_it & 1)
                par = (acc_it >> 1) & 1
                acc_it += 1
                if broken != "one_acc":
                    while not tempty[a].try_wait(par if broken == "acc_parity" else par ^ 1):
                        yield False
                if acc[a] is not None and len(acc[a]["readers"]) < epi_warps:
                    violated[0] = True           # accumulator overwritten while an epilogue warp still has to read it
                acc[a] = {"tile": t, "k": 0, "readers": set()}
                for k in range(kt):
                    s = it % stages
                    while not full[s].try_wait((it // stages) & 1):
                        yield False
                    if slot[s] != (t, k):
                        violated[0] = True       # wrong operands under the tensor core
                    slot[s] = None
                    acc[a]["k"] += 1
                    empty[s].arrive()            # tcgen05.commit -> empty[s]
                    it += 1
                    yield True
                tfull[a].arrive()                # tcgen05.commit -> tmem_full[acc]
                yield True

        def epilogue(wid):
            acc_it = 0
            for t, kt in enumerate(tile_k):
                if kt == 0:
                    yield True                   # zeros from registers, no accumulator involved
                    continue
                a = 0 if broken == "one_acc" else (acc_it & 1)
                par = (acc_it >> 1) & 1
                acc_it += 1
                while not tfull[a].try_wait(par):
                    yield False
                if acc[a] is None or acc[a]["tile"] != t or acc[a]["k"] != kt:
                    violated[0] = True           # read an accumulator that is not (all of) this tile
                else:
                    acc[a]["readers"].add(wid)
                tempty[a].arrive()
                yield True

        agents = [producer(), mma()] + [epilogue(w) for w in range(epi_warps)]
        alive = list(range(len(agents)))
        stuck = 0
        while alive and not violated[0]:
            i = rng.choice(alive)
            try:
                progressed = next(agents[i])
            except StopIteration:
                alive.remove(i)
                stuck = 0
                continue
            stuck = 0 if progressed else stuck + 1
            if stuck > 50 * len(agents) * (stages + 4):
                violated[0] = True               # nobody can move: deadlock
        bad += 1 if violated[0] else 0
    return bad


def simulate_bulk_ring(nchunks: int = 11, stages: int = 3, warps: int = 16, trials: int = 200, seed: int = 0,
                       prefetch: int = -1) -> int:
    """The shared-memory ring of the bulk-copy all-reduce (csrc/comm.cu allreduce_bulk_kernel): lane 0 of warp 0 both
    issues the copies (chunk ``ck + stages - 1`` at the top of iteration ``ck``, after waiting ``empty[slot]`` on parity
    ``((ck' // stages) & 1) ^ 1``) and consumes like every other warp (wait ``full[slot]`` on ``(ck // stages) & 1``, sum,
    arrive on ``empty[slot]``, count = warps).  Violation = a slot refilled before every warp has consumed it, a warp
    summing a slot that does not hold its chunk, or a deadlock (the issuing warp waiting for itself).
    ``prefetch`` overrides the distance (``stages`` instead of ``stages - 1`` makes warp 0 wait for its own arrival:
    the model must report the deadlock)."""
    rng = random.Random(seed)
    dist_ = stages - 1 if prefetch < 0 else prefetch
    bad = 0
    for _ in range(trials):
        full = [_MBar(1) for _ in range(stages)]
        empty = [_MBar(warps) for _ in range(stages)]
        slot = [None] * stages                  # {"chunk": ck, "readers": set()}
        violated = [False]

        def issue(ck):
            s = ck % stages
            while not empty[s].try_wait(((ck // stages) & 1) ^ 1):
                yield False
            if slot[s] is not None and len(slot[s]["readers"]) < warps:
                violated[0] = True
            slot[s] = {"chunk": ck, "readers": set()}
            full[s].arrive()                    # the copies land, complete_tx flips the phase
            yield True

        def warp(wid):
            if wid == 0:
                for ck in range(min(dist_, nchunks)):
                    yield from issue(ck)
            for ck in range(nchunks):
                if wid == 0 and ck + dist_ < nchunks:
                    yield from issue(ck + dist_)
                s = ck % stages
                while not full[s].try_wait((ck // stages) & 1):
                    yield False
                if slot[s] is None or slot[s]["chunk"] != ck:
                    violated[0] = True
                else:
                    slot[s]["readers"].add(wid)
                empty[s].arrive()
                yield True

        agents = [warp(w) for w in range(warps)]
        alive = list(range(warps))
        stuck = 0
        while alive and not violated[0]:
            i = rng.choice(alive)
            try:
                progressed = next(agents[i])
            except StopIteration:
                alive.remove(i)
                stuck = 0
                continue
            stuck = 0 if progressed else stuck + 1
            if stuck > 200 * warps:
                violated[0] = True
        bad += 1 if violated[0] else 0
    return bad
