"""Randomised model check of the mbarrier pipeline of the SM-pair GEMM (ops/csrc/gemm2cta_sm100.cu).

compute-sanitizer's racecheck cannot follow the orderings this kernel relies on (TMA transaction bytes completing on the
LEADER's barrier, `tcgen05.commit` multicast to both CTAs' barriers, remote arrives of the peer's epilogue warps), so the
protocol is checked as a model: the actors of one CTA pair -- two TMA producers, the leader's MMA thread, 2 x 4 epilogue warps --
run as coroutines under a random scheduler with random completion delays for every asynchronous operation (TMA landing, MMA
completion, remote arrival).  mbarriers are modelled with their real semantics (pending arrival count, transaction count that
may go negative, phase parity, `try_wait.parity`).  Invariants on every interleaving:
  * no deadlock;
  * when the MMA thread consumes ring stage s, BOTH CTAs' buffers hold exactly the (tile, k-block) it expects;
  * a ring stage is never refilled while an MMA that reads it is still in flight;
  * an accumulator stage is never written while one of the 8 epilogue warps still has to read the previous tile, and every
    epilogue warp reads exactly the tile it expects, complete.
Hardware memory-ordering effects are out of scope (the GPU numerics tests cover the real thing).
"""
import random

import pytest

STAGES, ACC = 3, 2            # smaller than the kernel's 6 / 2 so that wrap-arounds happen often
STAGE_BYTES = 32              # per CTA per stage (A + B half), arbitrary unit


class MBar:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.completed = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.completed += 1
            self.pending = self.init

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, nbytes):
        self.tx -= nbytes                      # may go negative before the expect_tx of the phase: allowed
        self._maybe_complete()

    def passed(self, parity):                  # try_wait.parity: has the phase with this parity completed?
        return (self.completed & 1) != parity


class Pair:
    def __init__(self, rng, tiles, kblocks):
        self.rng, self.tiles, self.kblocks = rng, tiles, kblocks
        self.full = [MBar(1) for _ in range(STAGES)]                    # leader only
        self.empty = [[MBar(1) for _ in range(STAGES)] for _ in range(2)]
        self.tmem_full = [[MBar(1) for _ in range(ACC)] for _ in range(2)]
        self.tmem_empty = [MBar(8) for _ in range(ACC)]                 # leader only
        self.smem = [[None] * STAGES for _ in range(2)]                 # (tile, kb) held by each CTA's ring stage
        self.smem_readers = [[0] * STAGES for _ in range(2)]            # MMAs in flight that read the stage
        self.acc = [[None] * ACC for _ in range(2)]                     # tile whose result the accumulator stage holds
        self.acc_unread = [[0] * ACC for _ in range(2)]                 # epilogue warps that still have to read it
        self.async_ops = []                                             # [remaining delay, callable]
        self.errors = []

    def later(self, fn):
        self.async_ops.append([self.rng.randint(0, 6), fn])


def producer(p, cta):
    stage = phase = 0
    for t in range(p.tiles):
        for kb in range(p.kblocks):
            while not p.empty[cta][stage].passed(phase ^ 1):
                yield
            if cta == 0:
                p.full[stage].arrive(expect_tx=2 * STAGE_BYTES)
            yield
            def land(s=stage, item=(t, kb)):
                if p.smem_readers[cta][s]:
                    p.errors.append(f"cta {cta} stage {s} refilled with {item} while an MMA still reads it")
                p.smem[cta][s] = item
                p.full[s].complete_tx(STAGE_BYTES)          # ALWAYS the leader's barrier
            p.later(land)
            stage += 1
            if stage == STAGES:
                stage, phase = 0, phase ^ 1
            yield


def mma(p):
    stage = phase = acc = acc_phase = 0
    for t in range(p.tiles):
        while not p.tmem_empty[acc].passed(acc_phase ^ 1):
            yield
        for c in range(2):
            if p.acc_unread[c][acc]:
                p.errors.append(f"tile {t}: accumulator stage {acc} of cta {c} overwritten with {p.acc_unread[c][acc]} readers left")
            p.acc[c][acc] = None
        for kb in range(p.kblocks):
            while not p.full[stage].passed(phase):
                yield
            for c in range(2):
                if p.smem[c][stage] != (t, kb):
                    p.errors.append(f"MMA of tile {t} k {kb} found {p.smem[c][stage]} in cta {c} stage {stage}")
                p.smem_readers[c][stage] += 1
            last = kb == p.kblocks - 1
            def retire(s=stage, a=acc, tile=t, last=last):
                # tcgen05.commit: fires when the MMAs issued so far have completed; multicast to both CTAs
                for c in range(2):
                    p.smem_readers[c][s] -= 1
                    p.empty[c][s].arrive()
                    if last:
                        p.acc[c][a] = tile
                        p.acc_unread[c][a] = 4
                        p.tmem_full[c][a].arrive()
            p.later(retire)
            stage += 1
            if stage == STAGES:
                stage, phase = 0, phase ^ 1
            yield
        acc += 1
        if acc == ACC:
            acc, acc_phase = 0, acc_phase ^ 1


def epilogue(p, cta, warp):
    acc = acc_phase = 0
    for t in range(p.tiles):
        while not p.tmem_full[cta][acc].passed(acc_phase):
            yield
        for _ in range(p.rng.randint(1, 4)):                # draining the accumulator takes a while
            yield
        if p.acc[cta][acc] != t:
            p.errors.append(f"epilogue cta {cta} warp {warp}: expected tile {t}, accumulator holds {p.acc[cta][acc]}")
        p.acc_unread[cta][acc] -= 1
        p.later(lambda a=acc: p.tmem_empty[a].arrive())      # (remote) arrive on the LEADER's barrier
        acc += 1
        if acc == ACC:
            acc, acc_phase = 0, acc_phase ^ 1
        yield


def run(seed, tiles=5, kblocks=4):
    rng = random.Random(seed)
    p = Pair(rng, tiles, kblocks)
    actors = [producer(p, 0), producer(p, 1), mma(p)] + [epilogue(p, c, w) for c in range(2) for w in range(4)]
    live = list(actors)
    idle_rounds = 0
    # in-order completion of one thread's MMAs (and of each producer's TMA queue) is NOT assumed for barrier arrivals of
    # different stages, only that an operation completes some time after it was issued
    while live or p.async_ops:
        progressed = False
        if p.async_ops and (not live or rng.random() < 0.4):
            op = rng.choice(p.async_ops)
            if op[0] > 0:
                op[0] -= 1
            else:
                p.async_ops.remove(op)
                op[1]()
            progressed = True
        elif live:
            a = rng.choice(live)
            before = _snapshot(p)
            try:
                next(a)
            except StopIteration:
                live.remove(a)
                progressed = True
            progressed = progressed or _snapshot(p) != before
        idle_rounds = 0 if progressed else idle_rounds + 1
        if idle_rounds > 20000:
            p.errors.append("deadlock")
            break
    return p.errors


def _snapshot(p):
    return (tuple(b.completed for b in p.full), tuple(b.completed for c in p.empty for b in c),
            tuple(b.completed for c in p.tmem_full for b in c), tuple(b.completed for b in p.tmem_empty),
            tuple(b.pending for b in p.full), tuple(b.pending for b in p.tmem_empty), len(p.async_ops))


@pytest.mark.parametrize("seed", range(40))
def test_pair_gemm_pipeline_is_safe_under_random_interleavings(seed):
    assert run(seed) == []


def test_the_model_catches_a_missing_accumulator_handshake():
    """Sanity of the checker itself: if the MMA thread did not wait for the epilogues (tmem_empty), overwrites are found."""
    found = 0
    for seed in range(40):
        rng = random.Random(seed)
        p = Pair(rng, 5, 4)
        p.tmem_empty = [_AlwaysOpen() for _ in range(ACC)]
        actors = [producer(p, 0), producer(p, 1), mma(p)] + [epilogue(p, c, w) for c in range(2) for w in range(4)]
        live = list(actors)
        steps = 0
        while (live or p.async_ops) and steps < 200000:
            steps += 1
            if p.async_ops and (not live or rng.random() < 0.4):
                op = rng.choice(p.async_ops)
                if op[0] > 0:
                    op[0] -= 1
                else:
                    p.async_ops.remove(op)
                    op[1]()
            elif live:
                a = rng.choice(live)
                try:
                    next(a)
                except StopIteration:
                    live.remove(a)
        found += bool(p.errors)
    assert found > 0


class _AlwaysOpen:
    completed = pending = 0

    def passed(self, parity):
        return True

    def arrive(self, expect_tx=0):
        pass
