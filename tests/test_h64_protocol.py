"""Discrete-event model of the synchronisation protocol of csrc/attn_fwd_h64_sm100.cu.

The kernel's four actors (TMA producer, MMA warp, two softmax warpgroups) are transcribed as coroutines over
mbarriers (phase / parity semantics of ``mbarrier.try_wait.parity``), an in-order tensor pipe with
``tcgen05.commit`` arrivals, and the named turn barriers.  Random latencies explore interleavings; every run
checks
  * progress (no deadlock) and that every actor finishes,
  * tensor-memory hazards: a Q K^T may only start into an S half that is EMPTY, the softmax only reads a FULL half,
    a P V only starts on a half whose P has been written, O is only rescaled while no P V of that q-tile is in
    flight,
  * shared-memory hazards: TMA only overwrites a K/V slot after every MMA that read the previous tile completed,
    and an MMA only reads a slot that holds the tile it expects.
It is a model of the protocol, not of the hardware - what it buys is that the first GPU run of the kernel is not
also the first time its barrier choreography is exercised.
"""
import random

import pytest

QK_CYCLES, PV_CYCLES = 384, 256


class Deadlock(Exception):
    pass


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.completed = count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.completed += 1

    def passed(self, parity):  # mbarrier.try_wait.parity
        return parity != (self.completed & 1)


class NamedBar:  # bar.sync / bar.arrive with two warpgroup-level participants
    def __init__(self):
        self.arrived, self.generation = 0, 0

    def arrive(self):
        self.arrived += 1
        if self.arrived == 2:
            self.arrived = 0
            self.generation += 1


class Sim:
    def __init__(self, tiles, has_t1, seed, rescale_prob):
        self.rng = random.Random(seed)
        self.tiles, self.has_t1, self.rescale_prob = tiles, has_t1, rescale_prob
        self.now = 0
        self.events = []  # (time, seq, fn)
        self.seq = 0
        self.pipe_end = 0  # completion time of the last issued tensor op
        self.bars = {}
        for i in range(4):
            self.bars[("kv_full", i)] = MBar(1)
            self.bars[("kv_empty", i)] = MBar(1)
        for t in range(2):
            self.bars[("pv_done", t)] = MBar(1)
            self.bars[("o_done", t)] = MBar(1)
            for h in range(2):
                self.bars[("s_full", t, h)] = MBar(1)
                self.bars[("p_ready", t, h)] = MBar(1)  # one arrival per warpgroup in the model
        self.named = {1: NamedBar(), 2: NamedBar()}
        # resource states
        self.s_state = {(t, h): "EMPTY" for t in range(2) for h in range(2)}
        self.slot_state = [("EMPTY", None)] * 4
        self.slot_readers = [0] * 4  # MMAs issued on the slot's current content and not yet complete
        self.pv_inflight = [0, 0]
        self.done = set()

    # ---- infrastructure
    def at(self, time, fn):
        self.seq += 1
        self.events.append((time, self.seq, fn))

    def issue(self, dur, on_start, on_end):
        start = max(self.now, self.pipe_end)
        end = start + dur
        self.pipe_end = end
        self.at(start, on_start)
        self.at(end, on_end)

    def commit(self, key):
        bar = self.bars[key]
        self.at(max(self.now, self.pipe_end), bar.arrive)

    def run(self, actors):
        waiting = {name: None for name in actors}  # name -> predicate or wake time
        gens = dict(actors)
        wake = {name: 0 for name in actors}
        while len(self.done) < len(actors):
            progressed = False
            # fire due events
            self.events.sort()
            while self.events and self.events[0][0] <= self.now:
                _, _, fn = self.events.pop(0)
                fn()
                progressed = True
            for name, gen in gens.items():
                if name in self.done or wake[name] > self.now:
                    continue
                pred = waiting[name]
                if pred is not None and not pred():
                    continue
                waiting[name] = None
                try:
                    req = next(gen)
                except StopIteration:
                    self.done.add(name)
                    progressed = True
                    continue
                progressed = True
                if req[0] == "sleep":
                    wake[name] = self.now + req[1]
                elif req[0] == "wait":
                    waiting[name] = req[1]
            if not progressed:
                nxt = [t for t, _, _ in self.events] + [w for n, w in wake.items() if n not in self.done and w > self.now]
                if not nxt:
                    raise Deadlock({n: waiting[n] is not None for n in gens if n not in self.done})
                self.now = min(nxt)

    def jitter(self, lo, hi):
        return self.rng.randint(lo, hi)

    # ---- actors (transcribed from the kernel)
    def producer(self):
        slot, phase = 0, 0
        for j, _ in enumerate(self.tiles):
            for kv in "KV":
                bar = self.bars[("kv_empty", slot)]
                yield ("wait", lambda b=bar, p=phase ^ 1: b.passed(p))
                assert self.slot_state[slot][0] == "EMPTY" and self.slot_readers[slot] == 0, \
                    f"TMA overwrites slot {slot} ({self.slot_state[slot]}) while it is in use"
                self.slot_state[slot] = ("LOADING", (kv, j))

                def landed(s=slot, c=(kv, j)):
                    self.slot_state[s] = ("FULL", c)
                    self.bars[("kv_full", s)].arrive()
                self.at(self.now + self.jitter(200, 2500), landed)
                slot += 1
                if slot == 4:
                    slot, phase = 0, phase ^ 1
                yield ("sleep", self.jitter(5, 40))

    def mma(self):
        slot, phase = 0, 0
        p_phase = {(t, h): 0 for t in range(2) for h in range(2)}
        o_started = [False, False]
        pend = {(t, h): False for t in range(2) for h in range(2)}
        pend_slot, pend_live, pend_tile = 0, False, None

        def read_slot(s, content):
            def start():
                assert self.slot_state[s] == ("FULL", content), f"MMA expects {content} in slot {s}: {self.slot_state[s]}"
            self.slot_readers[s] += 1
            return start

        def release_reader(s):
            def end():
                self.slot_readers[s] -= 1
            return end

        def flush_pv(t, h):
            nonlocal pend
            if not pend[(t, h)]:
                return
            bar = self.bars[("p_ready", t, h)]
            yield ("wait", lambda b=bar, p=p_phase[(t, h)]: b.passed(p))
            p_phase[(t, h)] ^= 1
            s, content = pend_slot, ("V", pend_tile)
            chk = read_slot(s, content)
            rel = release_reader(s)

            def start(t=t, h=h):
                chk()
                assert self.s_state[(t, h)] == "P", f"PV({t},{h}) starts on a half in state {self.s_state[(t, h)]}"
                self.s_state[(t, h)] = "PV"

            def end(t=t, h=h):
                rel()
                self.s_state[(t, h)] = "EMPTY"
                self.pv_inflight[t] -= 1
            self.pv_inflight[t] += 1
            self.issue(PV_CYCLES + self.jitter(0, 60), start, end)
            self.commit(("pv_done", t))
            o_started[t] = True
            pend[(t, h)] = False
            yield ("sleep", self.jitter(5, 60))

        for j, act in enumerate(self.tiles):
            k_slot, k_phase = slot, phase
            slot += 1
            if slot == 4:
                slot, phase = 0, phase ^ 1
            v_slot, v_phase = slot, phase
            slot += 1
            if slot == 4:
                slot, phase = 0, phase ^ 1
            bar = self.bars[("kv_full", k_slot)]
            yield ("wait", lambda b=bar, p=k_phase: b.passed(p))
            for h in range(2):
                for t in range(2):
                    yield from flush_pv(t, h)
                    if act[t]:
                        chk = read_slot(k_slot, ("K", j))
                        rel = release_reader(k_slot)

                        def start(t=t, h=h, chk=chk):
                            chk()
                            assert self.s_state[(t, h)] == "EMPTY", f"QK({t},{h}) overwrites a half in state {self.s_state[(t, h)]}"
                            self.s_state[(t, h)] = "QK"

                        def end(t=t, h=h, rel=rel):
                            rel()
                            self.s_state[(t, h)] = "FULL"
                        self.issue(QK_CYCLES + self.jitter(0, 60), start, end)
                        self.commit(("s_full", t, h))
                        yield ("sleep", self.jitter(5, 80))
            # commits: the slot becomes EMPTY for the producer when every MMA issued so far has completed
            def free(s):
                def fn():
                    assert self.slot_readers[s] == 0
                    self.slot_state[s] = ("EMPTY", None)
                    self.bars[("kv_empty", s)].arrive()
                return fn
            self.at(max(self.now, self.pipe_end), free(k_slot))
            if pend_live:
                self.at(max(self.now, self.pipe_end), free(pend_slot))
            pend_live = False
            bar = self.bars[("kv_full", v_slot)]
            yield ("wait", lambda b=bar, p=v_phase: b.passed(p))
            if act[0] or act[1]:
                for h in range(2):
                    pend[(0, h)] = act[0]
                    pend[(1, h)] = act[1]
                pend_slot, pend_live, pend_tile = v_slot, True, j
            else:
                self.at(max(self.now, self.pipe_end), free(v_slot))
            yield ("sleep", self.jitter(20, 400))  # loop bookkeeping
        for h in range(2):
            for t in range(2):
                yield from flush_pv(t, h)
        if pend_live:
            s = pend_slot

            def fn():
                assert self.slot_readers[s] == 0
                self.slot_state[s] = ("EMPTY", None)
                self.bars[("kv_empty", s)].arrive()
            self.at(max(self.now, self.pipe_end), fn)
        self.commit(("o_done", 0))
        self.commit(("o_done", 1))

    def softmax(self, t):
        if t == 1 and not self.has_t1:
            return
        s_phase = [0, 0]
        n_pv = 0
        handoffs_left = 2 * len(self.tiles) if self.has_t1 else 0
        if self.has_t1 and t == 1 and handoffs_left > 0:
            self.named[1].arrive()

        def turn_wait():
            if self.has_t1:
                nb = self.named[1 + t]
                gen = nb.generation
                nb.arrive()
                yield ("wait", lambda nb=nb, g=gen: nb.generation > g)

        def turn_pass():
            nonlocal handoffs_left
            if self.has_t1:
                handoffs_left -= 1
                if not (t == 1 and handoffs_left == 0):
                    self.named[1 + (1 - t)].arrive()

        for j, act in enumerate(self.tiles):
            if not act[t]:
                for _ in range(2):
                    yield from turn_wait()
                    turn_pass()
                continue
            for h in range(2):
                bar = self.bars[("s_full", t, h)]
                yield ("wait", lambda b=bar, p=s_phase[h]: b.passed(p))
                s_phase[h] ^= 1
                assert self.s_state[(t, h)] == "FULL", f"softmax {t} reads half {h} in state {self.s_state[(t, h)]}"
                self.s_state[(t, h)] = "SM"
                yield ("sleep", self.jitter(40, 120))   # tmem.ld
                yield ("sleep", self.jitter(100, 400))  # mask, row max
                if n_pv > 0 and self.rng.random() < self.rescale_prob:
                    bar = self.bars[("pv_done", t)]
                    yield ("wait", lambda b=bar, p=(n_pv - 1) & 1: b.passed(p))
                    assert self.pv_inflight[t] == 0, f"O{t} rescaled with {self.pv_inflight[t]} P V in flight"
                    assert bar.completed == n_pv, f"parity wait aliased: completed {bar.completed}, expected {n_pv}"
                    yield ("sleep", self.jitter(100, 300))
                    assert self.pv_inflight[t] == 0
                yield from turn_wait()
                yield ("sleep", self.jitter(500, 900))  # exp section, P store
                turn_pass()
                yield ("sleep", self.jitter(20, 80))
                self.s_state[(t, h)] = "P"
                self.bars[("p_ready", t, h)].arrive()
                n_pv += 1
        if n_pv > 0:
            bar = self.bars[("o_done", t)]
            yield ("wait", lambda b=bar: b.passed(0))
            assert self.pv_inflight[t] == 0


def _patterns():
    yield [(True, True)] * 6, True
    yield [(True, True)] * 1, True
    yield [(True, False)] * 4, False           # single q-tile work item
    yield [(True, True)] * 3 + [(False, True)] * 2, True       # causal: the upper q-tile stops earlier
    yield [(True, True), (False, True), (True, True), (True, True), (False, True)], True  # several segments
    yield [(False, True)] * 3 + [(True, True)] * 2, True


@pytest.mark.parametrize("case", range(6))
def test_h64_protocol_has_no_deadlock_or_hazard(case):
    tiles, has_t1 = list(_patterns())[case]
    for seed in range(60):
        for rescale_prob in (0.0, 0.3, 1.0):
            sim = Sim(tiles, has_t1, seed, rescale_prob)
            actors = {"producer": sim.producer(), "mma": sim.mma(), "sm0": sim.softmax(0), "sm1": sim.softmax(1)}
            sim.run(actors)
            assert all(v == "EMPTY" for v in sim.s_state.values())


def test_model_detects_a_broken_protocol():
    """Sanity check of the model itself: releasing the previous V slot BEFORE its P V GEMMs are issued must trip
    the shared-memory hazard check (or deadlock)."""
    class Broken(Sim):
        def mma(self):
            gen = super().mma()
            for req in gen:
                yield req

    tiles = [(True, True)] * 4
    sim = Broken(tiles, True, 0, 0.0)
    # sabotage: producer ignores kv_empty
    def bad_producer():
        slot = 0
        for j, _ in enumerate(tiles):
            for kv in "KV":
                assert sim.slot_state[slot][0] == "EMPTY" and sim.slot_readers[slot] == 0, "hazard"
                sim.slot_state[slot] = ("LOADING", (kv, j))

                def landed(s=slot, c=(kv, j)):
                    sim.slot_state[s] = ("FULL", c)
                    sim.bars[("kv_full", s)].arrive()
                sim.at(sim.now + 50, landed)
                slot = (slot + 1) % 4
                yield ("sleep", 10)
    with pytest.raises((AssertionError, Deadlock)):
        sim.run({"producer": bad_producer(), "mma": sim.mma(), "sm0": sim.softmax(0), "sm1": sim.softmax(1)})
