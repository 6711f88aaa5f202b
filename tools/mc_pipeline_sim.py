"""Discrete-event check of the multicast pipeline protocol of gemm_tc_mc.cuh (deadlock freedom + no stage overwritten
while a consumer may still read it + every consumer sees exactly the bytes of its round)."""
import heapq, random, sys

class MBar:
    def __init__(self, count):
        self.count0 = count; self.pending = count; self.tx = 0; self.phase = 0; self.waiters = []
    def _maybe_complete(self, sim):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1; self.pending = self.count0
            w, self.waiters = self.waiters, []
            for (parity, cb) in w:
                sim.try_wait(self, parity, cb)
    def arrive(self, sim, tx=0):
        self.tx += tx; self.pending -= 1; assert self.pending >= 0, "too many arrivals"
        self._maybe_complete(sim)
    def complete_tx(self, sim, nbytes):
        self.tx -= nbytes
        self._maybe_complete(sim)

class Sim:
    def __init__(self, seed):
        self.t = 0.0; self.q = []; self.n = 0; self.rng = random.Random(seed)
    def at(self, dt, fn):
        self.n += 1; heapq.heappush(self.q, (self.t + dt, self.n, fn))
    def try_wait(self, bar, parity, cb):
        # mbarrier.try_wait.parity: succeeds when the phase with that parity has completed, i.e. current phase != parity
        if bar.phase != parity: self.at(0.0, cb)
        else: bar.waiters.append((parity, cb))
    def run(self):
        while self.q:
            self.t, _, fn = heapq.heappop(self.q); fn()

def simulate(CL, STAGES, tiles, kblocks, seed):
    sim = Sim(seed); rnd = sim.rng
    A, Bp = 16, 32 // CL; STAGE = A + 32
    full = [[MBar(1) for _ in range(STAGES)] for _ in range(CL)]
    empty = [[MBar(CL) for _ in range(STAGES)] for _ in range(CL)]
    tfull = [[MBar(1) for _ in range(2)] for _ in range(CL)]
    tempty = [[MBar(1) for _ in range(2)] for _ in range(CL)]
    # per (cta, stage): round currently stored (set of contributions), and whether the MMA is reading it
    content = [[dict(round=-1, bytes=0) for _ in range(STAGES)] for _ in range(CL)]
    reading = [[None for _ in range(STAGES)] for _ in range(CL)]
    consumed = [[-1 for _ in range(STAGES)] for _ in range(CL)]   # last round fully consumed per stage
    done = dict(prod=0, mma=0, epi=0)
    mma_clock = [0.0] * CL
    total_k = tiles * kblocks

    def land(dst, s, rnd_no, nbytes):
        # a write into CTA dst's stage s for round rnd_no arrives
        assert reading[dst][s] is None, f"write into cta{dst} stage{s} while its MMA reads round {reading[dst][s]}"
        c = content[dst][s]
        if c["round"] != rnd_no:
            assert consumed[dst][s] == rnd_no - 1 or rnd_no == 0 and consumed[dst][s] == -1, \
                f"round {rnd_no} overwrites unconsumed round {c['round']} in cta{dst} stage{s}"
            c["round"], c["bytes"] = rnd_no, 0
        c["bytes"] += nbytes
        full[dst][s].complete_tx(sim, nbytes)

    def producer(cta):
        st = dict(i=0)
        def step():
            if st["i"] == total_k: done["prod"] += 1; return
            i = st["i"]; s = i % STAGES; ph = (i // STAGES) & 1
            def go():
                full[cta][s].arrive(sim, tx=STAGE)
                r = i // STAGES
                sim.at(rnd.uniform(0.5, 3.0), lambda: land(cta, s, r, A))
                for dst in range(CL):
                    sim.at(rnd.uniform(0.5, 3.0), (lambda d: lambda: land(d, s, r, Bp))(dst))
                st["i"] += 1
                sim.at(rnd.uniform(0.01, 0.2), step)
            sim.try_wait(empty[cta][s], ph ^ 1, go)
        sim.at(rnd.uniform(0, 1), step)

    def mma(cta):
        st = dict(tile=0, kb=0, i=0)
        def tile_start():
            if st["tile"] == tiles: done["mma"] += 1; return
            acc = st["tile"] & 1; aph = (st["tile"] >> 1) & 1
            sim.try_wait(tempty[cta][acc], aph ^ 1, kstep)
        def kstep():
            i = st["i"]; s = i % STAGES; ph = (i // STAGES) & 1
            def go():
                r = i // STAGES
                c = content[cta][s]
                assert c["round"] == r and c["bytes"] == STAGE, f"cta{cta} stage{s} round {r}: saw {c}"
                reading[cta][s] = r
                is_last = st["kb"] + 1 == kblocks
                acc_now = st["tile"] & 1
                def finished():
                    reading[cta][s] = None; consumed[cta][s] = r
                    for dst in range(CL):
                        sim.at(rnd.uniform(0.05, 0.5), (lambda d: lambda: empty[d][s].arrive(sim))(dst))
                    if is_last:   # commit after the tile's last k-block: accumulator complete
                        tfull[cta][acc_now].arrive(sim)
                # MMAs of one CTA complete in issue order: model with a per-CTA completion clock
                t_done = max(sim.t + rnd.uniform(0.2, 1.0), mma_clock[cta] + 0.01)
                mma_clock[cta] = t_done
                sim.at(t_done - sim.t, finished)
                st["i"] += 1; st["kb"] += 1
                if st["kb"] == kblocks:
                    st["kb"] = 0; st["tile"] += 1
                    sim.at(0.01, tile_start)
                else:
                    sim.at(0.01, kstep)
            sim.try_wait(full[cta][s], ph, go)
        sim.at(rnd.uniform(0, 1), tile_start)

    def epilogue(cta):
        st = dict(tile=0)
        def step():
            if st["tile"] == tiles: done["epi"] += 1; return
            acc = st["tile"] & 1; aph = (st["tile"] >> 1) & 1
            def go():
                def fin():
                    tempty[cta][acc].arrive(sim); st["tile"] += 1; step()
                sim.at(rnd.uniform(0.5, 6.0), fin)
            sim.try_wait(tfull[cta][acc], aph, go)
        sim.at(rnd.uniform(0, 1), step)

    for c in range(CL):
        producer(c); mma(c); epilogue(c)
    sim.run()
    assert done == dict(prod=CL, mma=CL, epi=CL), f"deadlock: {done}"

if __name__ == "__main__":
    n = 0
    for CL in (2, 4):
        for STAGES in (3, 4, 6):
            for tiles, kb in ((1, 1), (1, 8), (3, 8), (5, 2), (7, 13)):
                for seed in range(40):
                    simulate(CL, STAGES, tiles, kb, seed); n += 1
    print("ok", n, "simulations")
