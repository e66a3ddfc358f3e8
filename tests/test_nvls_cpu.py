"""NVLS host-side logic that needs no GPU: the arrival-counter bookkeeping of parallel/nvls.py and a discrete-event model of
its barrier protocol (every rank multicasts +1 to all ranks' counters, then spins on its local copy)."""
import heapq
import random

import pytest

from neuronx_distributed_training_b200.parallel.nvls import NvlsWorkspace


def _bare_workspace(world=4):
    ws = NvlsWorkspace.__new__(NvlsWorkspace)
    ws.world, ws.mc_ptr, ws.local_ptr = world, 0x40000000, 0x7f0000000000
    ws.event, ws.count = 0, [0, 0]
    return ws


def test_events_alternate_between_two_counters_and_accumulate_targets():
    ws = _bare_workspace(world=4)
    seen = []
    for arrivals in (128, 8, 8, 128, 1):
        mc, loc, before = ws._next_event(arrivals)
        seen.append((mc - ws.mc_ptr, loc - ws.local_ptr, before))
    # counter 0: events 0, 2, 4; counter 1: events 1, 3; "before" is the value the counter has when the event starts
    assert seen == [(0, 0, 0), (4, 4, 0), (0, 0, 4 * 128), (4, 4, 4 * 8), (0, 0, 4 * 128 + 4 * 8)]
    assert ws.count == [4 * 128 + 4 * 8 + 4 * 1, 4 * 8 + 4 * 128]


def test_counters_wrap_at_32_bits():
    ws = _bare_workspace(world=8)
    ws.count = [0xFFFFFFF0, 0]
    _, _, before = ws._next_event(4)
    assert before == 0xFFFFFFF0 and ws.count[0] == (0xFFFFFFF0 + 32) & 0xFFFFFFFF == 0x10
    # the kernel compares (int32)(have - target) < 0: still ordered across the wrap
    have, target = 0xFFFFFFF8, 0x10
    assert ((have - target) & 0xFFFFFFFF) >= 0x80000000          # "have" is behind the target


def _simulate(ranks, events, counters, seed):
    """Each rank runs `events` barrier events in order: multicast one arrival per rank (delivered to every destination after
    an independent random delay), then wait until the local copy of the event's counter reaches its cumulative target.
    Returns the number of times a rank left an event before every rank's arrival FOR THAT EVENT had been delivered to it."""
    rng = random.Random(seed)
    per_counter = [0] * counters
    target = []
    for e in range(events):
        per_counter[e % counters] += ranks
        target.append(per_counter[e % counters])
    value = [[0] * counters for _ in range(ranks)]                # local counter copies
    delivered = [[0] * events for _ in range(ranks)]              # arrivals of event e delivered at rank r
    at = [0] * ranks                                              # event each rank is currently in
    queue, tick = [], 0
    def send(src, e, now):
        for dst in range(ranks):
            # occasionally a very slow path to one destination: the case the protocol has to survive
            delay = rng.choice((1, 2, 3, 50)) + rng.random()
            heapq.heappush(queue, (now + delay, next_id(), dst, e))
    ids = iter(range(10 ** 9))
    next_id = lambda: next(ids)
    for r in range(ranks):
        send(r, 0, rng.random())
    premature = 0
    while queue:
        now, _, dst, e = heapq.heappop(queue)
        value[dst][e % counters] += 1
        delivered[dst][e] += 1
        # ranks whose wait is now satisfied move on (possibly through several events)
        progressed = True
        while progressed:
            progressed = False
            for r in range(ranks):
                ev = at[r]
                if ev < events and value[r][ev % counters] >= target[ev]:
                    if delivered[r][ev] < ranks:
                        premature += 1
                    at[r] += 1
                    if at[r] < events:
                        send(r, at[r], now + rng.random() * 0.1)
                    progressed = True
    assert all(a == events for a in at), "the model deadlocked"
    return premature


def test_two_alternating_counters_never_release_a_rank_early():
    for seed in range(300):
        assert _simulate(ranks=4, events=12, counters=2, seed=seed) == 0


def test_a_single_counter_would_release_ranks_early():
    """The hazard the second counter removes: with one counter, arrivals of a fast rank for event i+1 complete a slow rank's
    wait for event i while an arrival for event i is still in flight."""
    assert sum(_simulate(ranks=4, events=12, counters=1, seed=seed) for seed in range(300)) > 0
