"""Model check of the two-ring barrier protocol of ganslate_b200/csrc/igemm_halo_narrow.cu (CPU, no GPU needed).

The kernel walks the taps of a convolution in order.  Activations arrive as one halo box per depth group (ring of
A_STAGES buffers), weights as stages of TG boxes of 64 / C taps each (ring of B_STAGES buffers); boxes straddle group
boundaries, so the two rings are decoupled and ONE producer thread feeds both.  Its rule: issue weight stage s, then
request every halo box whose buffer's previous user (group g - A_STAGES) has its last tap inside the stages issued so
far.  The danger of any such rule is a cycle -- the producer blocked on a halo buffer whose release needs a weight stage
it has not issued yet -- which is exactly what the single-buffered 32-channel configuration (A_STAGES = 1) hits with the
naive "request the next box as soon as the current group starts" order (the fault the last test injects).

Producer and MMA issuer are restated as coroutines with the SAME index / phase arithmetic as the CUDA code over the
mbarrier model of tests/test_cg2_protocol.py; TMA loads and the in-order tensor pipe complete at random later times.
Every configuration is run under many random interleavings and fails on dead-lock, on a load landing in a buffer an
unfinished MMA still reads, or on an MMA reading a buffer whose contents are not the ones its tap needs."""
import random

import pytest

from test_cg2_protocol import MBar


def simulate(seed, group_sizes, tpb, tg, a_stages, b_stages, naive=False):
    rng = random.Random(seed)
    ngroups = len(group_sizes)
    group_begin = [0]
    for n in group_sizes:
        group_begin.append(group_begin[-1] + n)
    ntaps = group_begin[-1]
    taps_per_stage = tg * tpb
    nblk = -(-ntaps // tpb)
    nbst = -(-nblk // tg)
    a_full = [MBar(1) for _ in range(a_stages)]
    a_empty = [MBar(1) for _ in range(a_stages)]
    b_full = [MBar(1) for _ in range(b_stages)]
    b_empty = [MBar(1) for _ in range(b_stages)]
    a_buf = [None] * a_stages      # group whose box has LANDED in the buffer
    b_buf = [None] * b_stages      # weight stage that has landed
    a_readers = [0] * a_stages     # issued, not yet completed MMAs reading the buffer
    b_readers = [0] * b_stages
    inflight = []                  # ("a" | "b", buffer, contents)
    pipe = []                      # in-order tensor pipe: ("mma", as, bs) / ("commit", barrier)

    def producer():
        def load_a(g):
            as_, it = g % a_stages, g // a_stages
            if it > 0:
                while not a_empty[as_].passed((it - 1) & 1):
                    yield
            a_full[as_].arrive(expect_tx=1)
            inflight.append(("a", as_, g))

        next_a = 0
        while next_a < ngroups and next_a < a_stages:
            yield from load_a(next_a)
            next_a += 1
        g_cur = 0
        for s in range(nbst):
            bs, it = s % b_stages, s // b_stages
            if it > 0:
                while not b_empty[bs].passed((it - 1) & 1):
                    yield
            b_full[bs].arrive(expect_tx=1)
            inflight.append(("b", bs, s))
            last_tap = min(ntaps, (s + 1) * taps_per_stage) - 1
            if naive:   # fault injection: "request box g + 1 as soon as the first weight stage of group g is out"
                while g_cur < ngroups and group_begin[g_cur] <= last_tap:
                    g_cur += 1
                    if next_a < ngroups and next_a <= g_cur:
                        yield from load_a(next_a)
                        next_a += 1
            else:       # the kernel's rule
                while next_a < ngroups and group_begin[next_a - a_stages + 1] - 1 <= last_tap:
                    yield from load_a(next_a)
                    next_a += 1
            yield

    def consumer():
        g, g_end = 0, group_begin[1]
        while not a_full[0].passed(0):
            yield
        tp = 0
        for s in range(nbst):
            bs = s % b_stages
            while not b_full[bs].passed((s // b_stages) & 1):
                yield
            tap_end = min(ntaps, (s + 1) * taps_per_stage)
            while tp < tap_end:
                if tp == g_end:
                    g += 1
                    g_end = group_begin[g + 1]
                    while not a_full[g % a_stages].passed((g // a_stages) & 1):
                        yield
                as_ = g % a_stages
                assert a_buf[as_] == g, f"tap {tp}: halo buffer holds group {a_buf[as_]}, needs {g}"
                assert b_buf[bs] == s, f"tap {tp}: weight buffer holds stage {b_buf[bs]}, needs {s}"
                a_readers[as_] += 1
                b_readers[bs] += 1
                pipe.append(("mma", as_, bs))
                if tp + 1 == g_end:
                    pipe.append(("commit", a_empty[as_]))
                tp += 1
                yield
            pipe.append(("commit", b_empty[bs]))
        pipe.append(("done", None))

    agents = [producer(), consumer()]
    alive = [True, True]
    finished = False
    idle = 0
    while not finished:
        choices = [i for i in range(2) if alive[i]] + (["tma"] if inflight else []) + (["pipe"] if pipe else [])
        assert choices, "dead-lock: nothing can run"
        c = rng.choice(choices)
        before = (a_full[0].phase, sum(b.phase for b in a_full + a_empty + b_full + b_empty), len(inflight), len(pipe))
        if c == "tma":
            kind, buf, what = inflight.pop(rng.randrange(len(inflight)))
            if kind == "a":
                assert a_readers[buf] == 0, "a halo box landed in a buffer that issued MMAs still read"
                a_buf[buf] = what
                a_full[buf].complete_tx(1)
            else:
                assert b_readers[buf] == 0, "a weight stage landed in a buffer that issued MMAs still read"
                b_buf[buf] = what
                b_full[buf].complete_tx(1)
        elif c == "pipe":
            op, x, y = (pipe.pop(0) + (None,))[:3]
            if op == "mma":
                a_readers[x] -= 1
                b_readers[y] -= 1
            elif op == "commit":
                x.arrive()
            else:
                finished = True
        else:
            try:
                next(agents[c])
            except StopIteration:
                alive[c] = False
        after = (a_full[0].phase, sum(b.phase for b in a_full + a_empty + b_full + b_empty), len(inflight), len(pipe))
        idle = idle + 1 if (before == after and c not in ("tma", "pipe")) else 0
        assert idle < 20000, "dead-lock: the agents spin without progress"


# (taps per depth group, taps per weight box, boxes per stage, halo stages, weight stages): the kernel's configurations
CONFIGS = [
    ([25] * 5, 2, 4, 2, 3),    # 32 channels, one patch per CTA: 5x5x5, 8 taps per stage
    ([25] * 5, 2, 4, 1, 2),    # 32 channels, four patches: single-buffered halo, two weight stages
    ([25] * 5, 4, 8, 2, 3),    # 16 channels, one patch: 32 taps per stage (stages straddle groups)
    ([25] * 5, 4, 4, 2, 4),    # 16 channels, four patches
    ([9] * 3, 2, 4, 1, 2),     # 3x3x3
    ([49], 4, 8, 2, 3),        # 2-D 7x7: one group
    ([25], 2, 2, 1, 2),        # 2-D 5x5, 64 output columns (two boxes per stage)
    ([1] * 7, 2, 4, 1, 2),     # degenerate: several groups inside one weight stage
    ([3, 1, 30, 2], 4, 1, 1, 2),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[f"{c[0][0]}x{len(c[0])}-tpb{c[1]}-tg{c[2]}-a{c[3]}-b{c[4]}" for c in CONFIGS])
def test_two_ring_protocol_never_deadlocks_or_overwrites(cfg):
    for seed in range(150):
        simulate(seed, *cfg)


def test_random_configurations():
    rng = random.Random(7)
    for _ in range(120):
        groups = [rng.randint(1, 30) for _ in range(rng.randint(1, 6))]
        cfg = (groups, rng.choice([2, 4]), rng.randint(1, 8), rng.choice([1, 2]), rng.randint(2, 4))
        for seed in range(12):
            simulate(seed, *cfg)


def test_the_checker_catches_the_naive_request_order():
    """With a single halo buffer, requesting box g + 1 as soon as the first weight stage of group g is out blocks the
    producer on a buffer whose release needs weight stages it has not issued: the model must report that dead-lock."""
    with pytest.raises(AssertionError, match="dead-lock"):
        for seed in range(40):
            simulate(seed, [25] * 5, 2, 4, 1, 2, naive=True)
