"""Randomised discrete-event simulation (CPU) of the mbarrier protocol of the DUAL `gemm_tc_kernel` (row-tile pairing with
optional k-skew, neurosis_b200/csrc/gemm_tc.cu): the producer warp, the MMA-issuing thread and the 16 epilogue warps of
a CTA pair are generators that follow the kernel's loops literally — same barriers, same counts, same PARITY expressions
— while TMA loads and tcgen05 commits complete asynchronously after random delays (MMAs retire in issue order, a commit
fires when everything issued before it has retired).  Checked under many random interleavings:
  * no deadlock (the kernel would trap after 2 s);
  * every parity wait is unambiguous: when it is evaluated, the barrier has completed either exactly the phase waited for
    or the one before it (a phase overrun would let a wait pass on the wrong phase);
  * data hazards: an MMA reads a ring slot only while it holds the k-iteration it expects and before the producer
    overwrites it; an accumulator is not written for scheduler tile n + 1 before every epilogue warp has read tile n;
    an epilogue warp reads an accumulator only after all of its MMAs have retired.
The unpaired kernel's protocol (the one measured on the GPU) runs through the same simulator as a sanity check of the
simulator itself."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count


class Sim:
    def __init__(self, seed, stages, tiles, iters, dual, skew):
        self.rng = random.Random(seed)
        self.S, self.T, self.K, self.dual, self.skew = stages, tiles, iters, dual, skew
        self.full = [Bar(1) for _ in range(stages)]      # arrive.expect_tx + bytes: modelled as one arrival when the loads land
        self.empty = [Bar(1) for _ in range(stages)]
        self.tfull = [Bar(1), Bar(1)]
        self.tempty = [Bar(16), Bar(16)]                 # 8 epilogue warps x 2 CTAs
        self.slot = [None] * stages                      # (tile, k) currently in the ring slot
        self.acc_written = [None, None]                  # scheduler tile whose MMAs last wrote the accumulator
        self.acc_reads = [dict(), dict()]                # accumulator -> {tile: number of epilogue warps that read it}
        self.inflight = []                               # async events: ("load", slot, payload) | ("mma", ...) | ("commit", bar)
        self.mma_queue = []                              # issued, not yet retired MMAs / commits, in order
        self.retired = set()

    # -- parity wait with the unambiguity check
    def wait(self, bar, parity, intended):
        while True:
            assert bar.completed in (intended, intended + 1), ("phase overrun", bar.completed, intended)
            if (bar.completed & 1) != parity:
                assert bar.completed == intended + 1
                return
            yield

    def producer(self):
        stage, phase, n = 0, 0, 0
        for t in range(self.T):
            for k in range(self.K):
                yield from self.wait(self.empty[stage], phase ^ 1, n // self.S - 1)
                self.inflight.append(("load", stage, (t, k)))
                n += 1
                stage += 1
                if stage == self.S:
                    stage, phase = 0, phase ^ 1
                yield

    def issue_mma(self, stage, tile, k, acc):
        assert self.slot[stage] == (tile, k), ("MMA reads a slot that does not hold its k-iteration", self.slot[stage], (tile, k))
        self.mma_queue.append(("mma", stage, tile, k, acc))

    def commit(self, bar, tag):
        self.mma_queue.append(("commit", bar, tag))

    def mma_thread(self):
        stage, phase, nfull = 0, 0, 0
        if self.dual:
            for local in range(self.T):
                sk = min(self.skew, self.S - 1, self.K)
                par = (local & 1) ^ 1
                st1 = stage
                for j in range(self.K + sk):
                    if j < self.K:
                        yield from self.wait(self.full[stage], phase, nfull // self.S)
                        nfull += 1
                        if j == 0:
                            yield from self.wait(self.tempty[0], par, local - 1)
                            self.begin_acc(0, local)
                        self.issue_mma(stage, local, j, 0)
                        if j == self.K - 1:
                            self.commit(self.tfull[0], ("tfull", 0, local))
                        stage += 1
                        if stage == self.S:
                            stage, phase = 0, phase ^ 1
                        yield
                    if j >= sk:
                        j1 = j - sk
                        if j1 == 0:
                            yield from self.wait(self.tempty[1], par, local - 1)
                            self.begin_acc(1, local)
                        self.issue_mma(st1, local, j1, 1)
                        if j1 == self.K - 1:
                            self.commit(self.tfull[1], ("tfull", 1, local))
                        self.commit(self.empty[st1], ("empty", st1))
                        st1 = (st1 + 1) % self.S
                        yield
        else:
            for local in range(self.T):
                acc = local & 1
                yield from self.wait(self.tempty[acc], ((local >> 1) & 1) ^ 1, (local >> 1) - 1)
                self.begin_acc(acc, local)
                for k in range(self.K):
                    yield from self.wait(self.full[stage], phase, nfull // self.S)
                    nfull += 1
                    self.issue_mma(stage, local, k, acc)
                    self.commit(self.empty[stage], ("empty", stage))
                    stage += 1
                    if stage == self.S:
                        stage, phase = 0, phase ^ 1
                    yield
                self.commit(self.tfull[acc], ("tfull", acc, local))

    def begin_acc(self, acc, tile):
        prev = self.acc_written[acc]
        if prev is not None:  # every epilogue warp of the pair has read the previous contents
            assert self.acc_reads[acc].get(prev, 0) == 16, ("accumulator overwritten before it was drained", acc, prev)
        self.acc_written[acc] = tile

    def epilogue_warp(self):
        ntiles = self.T * (2 if self.dual else 1)  # 128-row tiles seen by the epilogue
        for local in range(ntiles):
            acc = local & 1
            sched = (local >> 1) if self.dual else local
            yield from self.wait(self.tfull[acc], (local >> 1) & 1, local >> 1)
            assert self.acc_written[acc] == sched, ("epilogue reads an accumulator of another tile", acc, self.acc_written[acc], sched)
            assert all(("mma", sched, k, acc) in self.retired for k in range(self.K)), "accumulator read before its MMAs retired"
            for _ in range(self.rng.randint(0, 3)):
                yield
            self.acc_reads[acc][sched] = self.acc_reads[acc].get(sched, 0) + 1
            self.tempty[acc].arrive()
            yield

    def step_async(self):
        """fire at most one asynchronous completion: a landed TMA load, or the oldest MMA / commit in the tensor pipe."""
        choices = []
        if self.inflight:
            choices.append("load")
        if self.mma_queue:
            choices.append("mma")
        if not choices:
            return False
        if self.rng.choice(choices) == "load":
            _, stage, payload = self.inflight.pop(self.rng.randrange(len(self.inflight)))
            # the slot must be free: every MMA that read its previous contents has retired (that is what empty_bar says)
            prev = self.slot[stage]
            if prev is not None:
                accs = (0, 1) if self.dual else (prev[0] & 1,)
                assert all(("mma", prev[0], prev[1], a) in self.retired for a in accs), ("ring slot overwritten while in use", prev)
            self.slot[stage] = payload
            self.full[stage].arrive()
        else:
            ev = self.mma_queue.pop(0)
            if ev[0] == "mma":
                _, stage, tile, k, acc = ev
                assert self.slot[stage] == (tile, k), ("slot changed under a pending MMA", self.slot[stage], (tile, k))
                self.retired.add(("mma", tile, k, acc))
            else:
                ev[1].arrive()
        return True

    def run(self):
        threads = [self.producer(), self.mma_thread()] + [self.epilogue_warp() for _ in range(16)]
        alive = list(threads)
        idle_rounds = 0
        while alive:
            progressed = False
            self.rng.shuffle(alive)
            for th in list(alive):
                if self.rng.random() < 0.35:
                    continue
                before = (tuple(b.completed for b in self.full + self.empty + self.tfull + self.tempty), len(self.inflight),
                          len(self.mma_queue), len(self.retired))
                try:
                    next(th)
                except StopIteration:
                    alive.remove(th)
                    progressed = True
                    continue
                after = (tuple(b.completed for b in self.full + self.empty + self.tfull + self.tempty), len(self.inflight),
                         len(self.mma_queue), len(self.retired))
                progressed = progressed or before != after
            if self.rng.random() < 0.7:
                progressed = self.step_async() or progressed
            idle_rounds = 0 if progressed else idle_rounds + 1
            if idle_rounds > 400:
                while self.step_async():
                    idle_rounds = 0
                if idle_rounds > 400:
                    raise AssertionError("deadlock: no thread can make progress")
        while self.step_async():
            pass
        want = self.T * self.K * (2 if self.dual else 1)
        assert len(self.retired) == want


@pytest.mark.parametrize("stages,iters,skew", [(3, 20, 0), (3, 20, 2), (3, 20, 3), (4, 20, 3), (3, 1, 3), (4, 2, 3), (3, 5, 1), (2, 7, 1),
                                               (3, 3, 2), (8, 20, 7)])
def test_dual_protocol_has_no_deadlock_overrun_or_hazard(stages, iters, skew):
    for seed in range(12):
        Sim(seed, stages, tiles=5, iters=iters, dual=True, skew=skew).run()


@pytest.mark.parametrize("stages,iters", [(7, 20), (3, 1), (8, 10), (2, 5)])
def test_unpaired_protocol_passes_the_same_simulator(stages, iters):
    for seed in range(8):
        Sim(seed, stages, tiles=6, iters=iters, dual=False, skew=0).run()


# ---- the simulator can tell: protocol mistakes one could plausibly make are caught -------------------------------------
class _EarlyRelease(Sim):
    """frees the ring slot after row tile 0 has read it instead of after row tile 1"""

    def mma_thread(self):
        stage, phase, nfull = 0, 0, 0
        for local in range(self.T):
            par = (local & 1) ^ 1
            for j in range(self.K):
                yield from self.wait(self.full[stage], phase, nfull // self.S)
                nfull += 1
                if j == 0:
                    yield from self.wait(self.tempty[0], par, local - 1)
                    self.begin_acc(0, local)
                self.issue_mma(stage, local, j, 0)
                self.commit(self.empty[stage], ("empty", stage))
                if j == self.K - 1:
                    self.commit(self.tfull[0], ("tfull", 0, local))
                if j == 0:
                    yield from self.wait(self.tempty[1], par, local - 1)
                    self.begin_acc(1, local)
                yield
                self.issue_mma(stage, local, j, 1)
                if j == self.K - 1:
                    self.commit(self.tfull[1], ("tfull", 1, local))
                stage += 1
                if stage == self.S:
                    stage, phase = 0, phase ^ 1
                yield


class _HalfCount(Sim):
    """tmem_empty initialised for one CTA's epilogue warps only (8 instead of 8 * CG)"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.tempty = [Bar(8), Bar(8)]


class _UnpairedParity(Sim):
    """epilogue waits tmem_full with parity local & 1 (accumulator index) instead of (local >> 1) & 1"""

    def epilogue_warp(self):
        for local in range(self.T * 2):
            acc = local & 1
            yield from self.wait(self.tfull[acc], local & 1, local >> 1)
            self.acc_reads[acc][local >> 1] = self.acc_reads[acc].get(local >> 1, 0) + 1
            self.tempty[acc].arrive()
            yield


@pytest.mark.parametrize("cls", [_EarlyRelease, _HalfCount, _UnpairedParity])
def test_simulator_detects_protocol_mistakes(cls):
    caught = 0
    for seed in range(6):
        try:
            cls(seed, 3, tiles=4, iters=8, dual=True, skew=0).run()
        except AssertionError:
            caught += 1
    assert caught == 6
