"""Discrete simulation of the A-ring / accumulator hand-over protocol of query_col_kernel<P = 1> with NG groups of
epilogue warps: every actor is a sequential program of mbarrier waits and arrivals; random interleavings; reports
deadlocks and waits that pass on the wrong phase (parity aliasing)."""
import random
import sys

NG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NSLOT = int(sys.argv[2]) if len(sys.argv) > 2 else 3
TILES = int(sys.argv[3]) if len(sys.argv) > 3 else 40
SEEDS = int(sys.argv[4]) if len(sys.argv) > 4 else 5
WPG = 8


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "over-arrival on " + self.name
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity):
        return (self.phase & 1) != parity


a_ready = [Bar("a_ready%d" % i, WPG) for i in range(NSLOT)]
a_ready_b = [Bar("a_ready_b%d" % i, WPG) for i in range(NSLOT)]
a_free = [Bar("a_free%d" % i, 2) for i in range(NSLOT)]
acc_full = [Bar("acc_full%d" % i, 1) for i in range(2)]
acc_free = [Bar("acc_free%d" % i, WPG * NG) for i in range(2)]
t1_free_b = Bar("t1_free_b", WPG * NG)
sync1 = Bar("bar1", WPG * NG)


def wait(bar, parity, want_phase=None):
    # want_phase: the phase count the barrier must have reached for this wait to be legitimate
    while not bar.passes(parity):
        yield
    if want_phase is not None and bar.phase < want_phase:
        raise RuntimeError("premature pass on %s: phase %d < %d" % (bar.name, bar.phase, want_phase))


def epi_warp(group, w):
    g = group
    acc0 = acc1 = 0
    nsync = 0

    def acquire():
        nonlocal g
        slot, rnd = g % NSLOT, g // NSLOT
        yield from wait(a_free[slot], (rnd & 1) ^ 1, rnd)
        return slot

    def drain(acc_id):
        nonlocal g
        nb = 4 // NG
        for i in range(nb):
            if i == nb - 1:
                acc_free[acc_id].arrive()
            slot = yield from acquire()
            yield
            a_ready[slot].arrive()
            g += 1 + (NG - 1)

    for tile in range(TILES):
        for m in range(2):
            for kb in range(group, 16, NG):
                slot = yield from acquire()
                yield
                a_ready[slot].arrive()
                a_ready_b[slot].arrive()
                g += 1 + (NG - 1)
            yield from wait(acc_full[0], acc0 & 1); acc0 += 1
            yield from drain(0)
            yield from wait(acc_full[1], acc1 & 1); acc1 += 1
            yield from drain(1)
            yield from wait(acc_full[0], acc0 & 1); acc0 += 1
            yield from drain(0)
            yield from wait(acc_full[1], acc1 & 1); acc1 += 1
            yield
            acc_free[1].arrive()
            t1_free_b.arrive()
            if m == 0:
                sync1.arrive()
                yield from wait(sync1, nsync & 1); nsync += 1


def issuer_a():
    ablk = acc0 = acc1 = 0
    for tile in range(TILES):
        for m in range(2):
            yield from wait(acc_free[0], (acc0 & 1) ^ 1)
            yield from wait(acc_free[1], (acc1 & 1) ^ 1)
            for kb in range(16):
                slot = ablk % NSLOT
                yield from wait(a_ready[slot], (ablk // NSLOT) & 1, ablk // NSLOT + 1)
                yield
                a_free[slot].arrive(); ablk += 1
            acc_full[0].arrive(); acc0 += 1; acc1 += 1
            yield from wait(acc_free[0], (acc0 & 1) ^ 1)
            for kb in range(8):
                slot = ablk % NSLOT
                yield from wait(a_ready[slot], (ablk // NSLOT) & 1, ablk // NSLOT + 1)
                yield
                a_free[slot].arrive(); a_free[slot].arrive(); ablk += 1
            acc_full[0].arrive(); acc0 += 1
            yield from wait(acc_free[1], (acc1 & 1) ^ 1)
            for kb in range(4):
                slot = ablk % NSLOT
                yield from wait(a_ready[slot], (ablk // NSLOT) & 1, ablk // NSLOT + 1)
                yield
                a_free[slot].arrive(); a_free[slot].arrive(); ablk += 1
            acc_full[1].arrive(); acc1 += 1


def issuer_b():
    ablk = 0
    aph = [0] * NSLOT
    tph = 0
    for tile in range(TILES):
        for m in range(2):
            yield from wait(t1_free_b, tph ^ 1); tph ^= 1
            for kb in range(16):
                slot = ablk % NSLOT
                yield from wait(a_ready_b[slot], aph[slot]); aph[slot] ^= 1
                yield
                a_free[slot].arrive(); ablk += 1
            acc_full[1].arrive()
            ablk += 12


def run(seed):
    random.seed(seed)
    actors = [epi_warp(g, w) for g in range(NG) for w in range(WPG)] + [issuer_a(), issuer_b()]
    names = ["g%dw%d" % (g, w) for g in range(NG) for w in range(WPG)] + ["A", "B"]
    alive = list(range(len(actors)))
    idle = 0
    state = lambda: tuple(b.phase * 100 + b.pending for b in a_ready + a_ready_b + a_free + acc_full + acc_free + [t1_free_b, sync1])
    last = state()
    steps = 0
    while alive:
        i = random.choice(alive)
        try:
            next(actors[i])
        except StopIteration:
            alive.remove(i)
        steps += 1
        if steps % 2000 == 0:
            s = state()
            if s == last:
                idle += 1
                if idle > 20:
                    print("DEADLOCK seed", seed, "alive:", [names[k] for k in alive])
                    for b in a_ready + a_ready_b + a_free + acc_full + acc_free + [t1_free_b, sync1]:
                        print("   ", b.name, "phase", b.phase, "pending", b.pending)
                    return False
            else:
                idle = 0
            last = s
    return True


ok = True
for seed in range(SEEDS):
    for b in a_ready + a_ready_b + a_free + acc_full + acc_free + [t1_free_b, sync1]:
        b.phase, b.pending = 0, b.count
    ok = run(seed) and ok
print("NG=%d NSLOT=%d:" % (NG, NSLOT), "all runs completed" if ok else "FAILED")
sys.exit(0 if ok else 1)
