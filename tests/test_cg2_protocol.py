"""Model check of the barrier protocol of ganslate_b200/csrc/igemm_cg2.cu (CPU, no GPU needed).

The kernel's three roles (TMA producer of each CTA, MMA issuer of the leader CTA, epilogue warps of both CTAs) are
restated as coroutines that use the SAME phase / parity arithmetic as the CUDA code, over a small model of mbarrier
semantics (pending arrivals + transaction bytes per phase; try_wait.parity(P) succeeds once the phase of parity P
has completed).  Asynchronous agents (TMA loads in flight, the in-order tensor pipe with its commits) complete at
random later times.  Thousands of random interleavings are run per configuration; the run fails on

  * dead-lock (no agent can make progress before all items are done),
  * a TMA load landing in a ring stage that an unfinished MMA still reads,
  * an MMA consuming a stage whose loads (of BOTH CTAs) have not landed or belong to another K block,
  * the first MMA of an item overwriting an accumulator that some epilogue warp has not finished reading,
  * an epilogue warp reading an accumulator before the item's last MMA has completed.

This pins the protocol logic (not the hardware semantics of cta_group::2, which only a B200 run can).
"""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.tx += expect_tx
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def passed(self, parity):
        return (self.phase & 1) != parity  # the phase with this parity has completed


def simulate(seed, pair, stages, items, kb_per_item, epi_warps=8, a_bytes=16384, b_bytes=16384, fault=None):
    rng = random.Random(seed)
    ncta = 2 if pair else 1
    # barriers of every CTA (the peer's full / tempty copies exist but are never used)
    full = [[MBar(1) for _ in range(stages)] for _ in range(ncta)]
    empty = [[MBar(1) for _ in range(stages)] for _ in range(ncta)]
    tfull = [[MBar(1) for _ in range(2)] for _ in range(ncta)]
    # fault injection (the checker must notice): "tempty_count" = barrier initialised for one CTA's warps only
    tempty_count = epi_warps if fault == "tempty_count" else epi_warps * ncta
    tempty = [[MBar(tempty_count) for _ in range(2)] for _ in range(ncta)]
    LEADER = 0
    stage_data = [[None] * stages for _ in range(ncta)]     # (item, kb) whose A / B-part landed in the stage
    stage_landed = [[0] * stages for _ in range(ncta)]      # loads landed for the current contents (2 per fill)
    mma_queue = []           # in-order tensor pipe: ("mma", item, kb, stage, acc) / ("commit", [barriers])
    inflight_tma = []        # (cta, stage, item, kb, nbytes)
    acc_written_by = [None, None]   # item whose MMAs are (being) accumulated in accumulator a
    acc_complete = [None, None]     # item whose last MMA has completed
    acc_reads_left = [0, 0]         # epilogue warps that still have to read the current contents
    done_epi = [0]

    def producer(cta):
        s, rnd = 0, 0
        tx = ncta * (a_bytes + b_bytes)
        for item in range(items):
            for kb in range(kb_per_item):
                if rnd > 0:
                    while not empty[cta][s].passed((rnd - 1) & 1):
                        yield
                if cta == LEADER:
                    full[LEADER][s].arrive(expect_tx=tx // ncta if fault == "tx_bytes" else tx)
                # the stage is about to be overwritten: nothing may still read it
                assert not any(e[0] == "mma" and e[3] == s for e in mma_queue), "TMA overwrites a stage an MMA still reads"
                stage_data[cta][s], stage_landed[cta][s] = (item, kb), 0
                inflight_tma.append((cta, s, item, kb, a_bytes))
                inflight_tma.append((cta, s, item, kb, b_bytes))
                s += 1
                if s == stages:
                    s, rnd = 0, rnd + 1
                yield

    def mma():
        s, rnd, acc_it = 0, 0, 0
        for item in range(items):
            a, use = acc_it & 1, acc_it >> 1
            if use > 0:
                while not tempty[LEADER][a].passed((use - 1) & 1):
                    yield
            for kb in range(kb_per_item):
                while not full[LEADER][s].passed((rnd + 1) & 1 if fault == "full_parity" else rnd & 1):
                    yield
                for cta in range(ncta):
                    assert stage_data[cta][s] == (item, kb) and stage_landed[cta][s] == 2, \
                        f"MMA consumes stage {s} of CTA {cta} holding {stage_data[cta][s]} ({stage_landed[cta][s]} loads landed), wants {(item, kb)}"
                if kb == 0:
                    assert acc_reads_left[a] == 0, "MMA overwrites an accumulator an epilogue warp has not read"
                    acc_written_by[a] = item
                mma_queue.append(("mma", item, kb, s, a))
                mma_queue.append(("commit", [empty[c][s] for c in range(ncta)]))
                s += 1
                if s == stages:
                    s, rnd = 0, rnd + 1
                yield
            mma_queue.append(("acc_done", item, a))
            mma_queue.append(("commit", [tfull[c][a] for c in range(ncta)]))
            acc_it += 1
            yield

    def epilogue(cta, w):
        acc_it = 0
        for item in range(items):
            a, use = acc_it & 1, acc_it >> 1
            acc_it += 1
            while not tfull[cta][a].passed(use & 1):
                yield
            assert acc_complete[a] == item, f"epilogue reads accumulator {a} holding item {acc_complete[a]}, wants {item}"
            yield  # tcgen05.ld in flight
            acc_reads_left[a] -= 1
            assert acc_reads_left[a] >= 0
            tempty[LEADER][a].arrive()
            yield  # stores, statistics
        done_epi[0] += 1

    agents = [producer(c) for c in range(ncta)] + [mma()] + [epilogue(c, w) for c in range(ncta) for w in range(epi_warps)]
    alive = list(range(len(agents)))
    idle_rounds = 0
    while alive or mma_queue or inflight_tma:
        progressed = False
        choices = [("agent", i) for i in alive]
        if inflight_tma:
            choices.append(("tma", None))
        if mma_queue:
            choices.append(("pipe", None))
        kind, i = rng.choice(choices)
        if kind == "tma":
            cta, s, item, kb, nbytes = inflight_tma.pop(rng.randrange(len(inflight_tma)))
            assert stage_data[cta][s] == (item, kb)
            stage_landed[cta][s] += 1
            full[LEADER][s].complete_tx(nbytes)
            progressed = True
        elif kind == "pipe":
            e = mma_queue.pop(0)  # in order
            if e[0] == "commit":
                for b in e[1]:
                    b.arrive()
            elif e[0] == "acc_done":
                _, item, a = e
                acc_complete[a] = item
                acc_reads_left[a] = epi_warps * ncta
            progressed = True
        else:
            before = _snapshot(full, empty, tfull, tempty)
            try:
                next(agents[i])
            except StopIteration:
                alive.remove(i)
                progressed = True
            if _snapshot(full, empty, tfull, tempty) != before or len(mma_queue) or len(inflight_tma):
                progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 20000, f"dead-lock: agents {alive} wait forever (seed {seed})"
    assert done_epi[0] == epi_warps * ncta


def _snapshot(*groups):
    return tuple((b.phase, b.pending, b.tx) for g in groups for row in g for b in row)


@pytest.mark.parametrize("pair", [True, False], ids=["cta-pair", "single-cta"])
@pytest.mark.parametrize("stages,items,kb", [(1, 3, 2), (2, 5, 3), (4, 4, 9), (6, 3, 36), (3, 7, 1), (6, 1, 4), (5, 6, 5)])
def test_cg2_barrier_protocol(pair, stages, items, kb):
    for seed in range(60):
        simulate(seed, pair, stages, items, kb)


@pytest.mark.parametrize("fault", ["tempty_count", "full_parity", "tx_bytes"])
def test_model_detects_a_broken_protocol(fault):
    """Sanity of the checker itself: a tempty barrier initialised for one CTA's warps, a full-barrier wait on the
    wrong parity, or an expect_tx that forgets the peer's bytes must all be reported."""
    with pytest.raises(AssertionError):
        for seed in range(40):
            simulate(seed, True, 3, 6, 4, fault=fault)
