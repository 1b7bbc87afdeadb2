#!/usr/bin/env python3
"""Discrete-event model of the barrier protocol of attn_umma_kernel_v8 (csrc/attn_umma_bf16.cu).

The persistent kernel carries its mbarrier phases, TMEM buffers and shared-memory slots across SEGMENTS of a
cluster's work range; a wrong parity or a missing hand-over deadlocks or corrupts silently on the GPU.  This model
replays the kernel's control flow -- the same loops, the same parity formulas, the same arrival counts -- for the
roles of one cluster (2 TMA producers, the MMA issuer, 32 softmax warps) under a randomised scheduler, with
asynchronous completions (TMA bytes, tcgen05.commit) delivered at random later times, and checks that

  * every role terminates (no deadlock),
  * a wait never passes on an OLDER phase of the same parity (each wait also states which completion it means),
  * no barrier phase receives more arrivals than its count,
  * every buffer hand-over holds: S / P double buffers, the O accumulator, the Q slots, the K / V rings.

Run: python tools/v8_protocol_sim.py [trials]
"""
import random
import sys

KST, VST = 4, 3


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done = name, count, 0, 0

    def arrive(self, n=1):
        self.pending += n
        assert self.pending <= self.count, f"over-arrival on {self.name}: {self.pending}/{self.count} in phase {self.done}"
        if self.pending == self.count:
            self.pending = 0
            self.done += 1

    def hw_pass(self, parity):
        return parity != (self.done & 1)


class Named:
    """bar.sync id, n: generation barrier of n warps."""

    def __init__(self, n):
        self.n, self.waiting, self.gen = n, 0, 0


def segments(begin, end, T):
    u, seg = begin, 0
    out = []
    while u < end:
        rb, t0 = divmod(u, T)
        nt = min(T - t0, end - u)
        out.append((seg, rb, t0, nt))
        u += nt
        seg += 1
    return out


class Sim:
    def __init__(self, begin, end, T, rng):
        self.rng = rng
        self.segs = segments(begin, end, T)
        self.G = end - begin
        L = 0  # leader
        self.bars = {}

        def mk(name, count, per_cta=False):
            if per_cta:
                for r in (0, 1):
                    self.bars[(name, r)] = Bar(f"{name}@{r}", count)
            else:
                self.bars[(name, L)] = Bar(f"{name}@leader", count)

        for i in range(2):
            mk(f"q_full{i}", 2 + 2)          # 2 producer arrivals + the bytes of both CTAs
            mk(f"q_free{i}", 1, True)
            mk(f"s_full{i}", 1, True)
            mk(f"p_ready{i}", 16)
            mk(f"s_free{i}", 16)
            mk(f"pv_done{i}", 1, True)
        for i in range(KST):
            mk(f"k_full{i}", 2 + 2)
            mk(f"k_empty{i}", 1, True)
        for i in range(VST):
            mk(f"v_full{i}", 2 + 2)
            mk(f"v_empty{i}", 1, True)
        mk("o_done", 1, True)
        mk("o_free", 32)
        self.async_events = []   # (kind, payload): TMA completions (any order), commits (FIFO)
        self.commit_fifo = []
        # resource state for hand-over checks
        self.S = [dict(tile=None, readers=16) for _ in range(2)]       # readers: warps that have read the tile
        self.P = [dict(tile=None, writers=0, consumed=True) for _ in range(2)]
        self.O = dict(seg=None, readers=32)
        self.Q = [[dict(seg=None, s_left=0) for _ in range(2)] for _ in range(2)]   # [cta][slot]
        self.K = [[dict(g=None, used=True) for _ in range(KST)] for _ in range(2)]
        self.V = [[dict(g=None, used=True) for _ in range(VST)] for _ in range(2)]
        self.named = {}
        self.busy = 0   # bumped by roles that model a long stretch of work between two protocol steps

    # ---- primitives used by the role generators
    def wait(self, name, cta, parity, want):
        b = self.bars[(name, cta)]
        while not b.hw_pass(parity):
            yield
        assert b.done >= want, f"{b.name}: wait(parity={parity}) passed at completion {b.done}, meant {want}"

    def named_sync(self, key, n):
        nb = self.named.setdefault(key, Named(n))
        gen = nb.gen
        nb.waiting += 1
        if nb.waiting == nb.n:
            nb.waiting = 0
            nb.gen += 1
        while nb.gen == gen:
            yield

    def tma(self, barname, fn):
        self.async_events.append((barname, fn))

    def commit(self, actions):
        self.commit_fifo.append(actions)

    # ---- roles
    def producer(self, r):
        g = 0
        for seg, rb, t0, nt in self.segs:
            qs = seg & 1
            if seg >= 2:
                yield from self.wait(f"q_free{qs}", r, ((seg >> 1) - 1) & 1, (seg >> 1))
            q = self.Q[r][qs]
            assert q["s_left"] == 0, f"Q slot {qs} of CTA {r} overwritten while S MMAs of segment {q['seg']} are pending"
            self.bars[(f"q_full{qs}", 0)].arrive()

            def q_landed(r=r, qs=qs, seg=seg, nt=nt):
                self.Q[r][qs].update(seg=seg, s_left=nt)
                self.bars[(f"q_full{qs}", 0)].arrive()
            self.tma(f"q_full{qs}", q_landed)
            for j in range(nt):
                ks, vs = g % KST, g % VST
                yield from self.wait(f"k_empty{ks}", r, ((g // KST) & 1) ^ 1, g // KST)
                assert self.K[r][ks]["used"], f"K stage {ks} of CTA {r} overwritten before use"
                self.bars[(f"k_full{ks}", 0)].arrive()

                def k_landed(r=r, ks=ks, g=g):
                    self.K[r][ks].update(g=g, used=False)
                    self.bars[(f"k_full{ks}", 0)].arrive()
                self.tma(f"k_full{ks}", k_landed)
                yield from self.wait(f"v_empty{vs}", r, ((g // VST) & 1) ^ 1, g // VST)
                assert self.V[r][vs]["used"], f"V stage {vs} of CTA {r} overwritten before use"
                self.bars[(f"v_full{vs}", 0)].arrive()

                def v_landed(r=r, vs=vs, g=g):
                    self.V[r][vs].update(g=g, used=False)
                    self.bars[(f"v_full{vs}", 0)].arrive()
                self.tma(f"v_full{vs}", v_landed)
                g += 1
                yield

    def mma(self):
        segs = self.segs
        cs = dict(i=0, j=0)
        cp = dict(i=0, j=0)

        def issue_s(g):
            seg, rb, t0, nt = segs[cs["i"]]
            sb, ks, qs = g & 1, g % KST, seg & 1
            if cs["j"] == 0:
                yield from self.wait(f"q_full{qs}", 0, (seg >> 1) & 1, (seg >> 1) + 1)
            yield from self.wait(f"k_full{ks}", 0, (g // KST) & 1, g // KST + 1)
            for r in (0, 1):
                assert self.Q[r][qs]["seg"] == seg, f"S({g}) reads Q slot {qs} holding segment {self.Q[r][qs]['seg']}, wants {seg}"
                assert self.K[r][ks]["g"] == g and not self.K[r][ks]["used"], f"S({g}) reads K stage {ks} holding {self.K[r][ks]}"
            assert self.S[sb]["readers"] == 16, f"S({g}) overwrites S buffer {sb} read by only {self.S[sb]['readers']} warps"
            self.S[sb] = dict(tile=None, readers=0, pending=g)
            last = cs["j"] == nt - 1

            def done(g=g, sb=sb, ks=ks, qs=qs, last=last):
                self.S[sb]["tile"] = g
                for r in (0, 1):
                    self.K[r][ks]["used"] = True
                    self.Q[r][qs]["s_left"] -= 1
                    self.bars[(f"s_full{sb}", r)].arrive()
                    self.bars[(f"k_empty{ks}", r)].arrive()
                    if last:
                        assert self.Q[r][qs]["s_left"] == 0
                        self.bars[(f"q_free{qs}", r)].arrive()
            self.commit(done)
            cs["j"] += 1
            if cs["j"] == nt:
                cs["i"] += 1
                cs["j"] = 0

        def issue_pv(g):
            seg, rb, t0, nt = segs[cp["i"]]
            pb, vs = g & 1, g % VST
            first, last = cp["j"] == 0, cp["j"] == nt - 1
            if first and seg >= 1:
                yield from self.wait("o_free", 0, (seg - 1) & 1, seg)
            yield from self.wait(f"v_full{vs}", 0, (g // VST) & 1, g // VST + 1)
            yield from self.wait(f"p_ready{pb}", 0, (g >> 1) & 1, (g >> 1) + 1)
            assert self.P[pb]["tile"] == g and self.P[pb]["writers"] == 16, f"PV({g}) reads P buffer {pb}: {self.P[pb]}"
            for r in (0, 1):
                assert self.V[r][vs]["g"] == g and not self.V[r][vs]["used"], f"PV({g}) reads V stage {vs}: {self.V[r][vs]}"
            if first:
                assert self.O["readers"] == 32, f"PV({g}) overwrites O of segment {self.O['seg']} read by {self.O['readers']} warps"
                self.O = dict(seg=seg, readers=0, complete=False)
            assert self.O["seg"] == seg

            def done(g=g, pb=pb, vs=vs, last=last):
                self.P[pb]["consumed"] = True
                for r in (0, 1):
                    self.V[r][vs]["used"] = True
                    self.bars[(f"pv_done{pb}", r)].arrive()
                    self.bars[(f"v_empty{vs}", r)].arrive()
                    if last:
                        self.bars[("o_done", r)].arrive()
                if last:
                    self.O["complete"] = True
            self.commit(done)
            cp["j"] += 1
            if cp["j"] == nt:
                cp["i"] += 1
                cp["j"] = 0

        yield from issue_s(0)
        if self.G > 1:
            yield from issue_s(1)
        for g in range(self.G):
            if g + 2 < self.G:
                yield from self.wait(f"s_free{g & 1}", 0, (g >> 1) & 1, (g >> 1) + 1)
                yield from issue_s(g + 2)
            yield from issue_pv(g)
            yield

    def softmax(self, r, group, half, quad):
        def tile_step(g, first):
            sb = g & 1
            yield from self.wait(f"s_full{sb}", r, (g >> 1) & 1, (g >> 1) + 1)
            assert self.S[sb]["tile"] == g, f"softmax({g}) reads S buffer {sb} holding {self.S[sb]}"
            self.S[sb]["readers"] += 1
            self.bars[(f"s_free{sb}", 0)].arrive()
            yield
            if first:
                yield from self.named_sync(("id", r, group, quad), 2)
                yield from self.named_sync(("all", r, quad), 4)
            if g >= 2:
                yield from self.wait(f"pv_done{sb}", r, ((g >> 1) - 1) & 1, (g >> 1))
            p = self.P[sb]
            if p["tile"] != g:
                assert p["consumed"], f"softmax({g}) overwrites P buffer {sb} (tile {p['tile']}) before its PV completed"
                self.P[sb] = dict(tile=g, writers=0, consumed=False)
            self.P[sb]["writers"] += 1
            self.bars[(f"p_ready{sb}", 0)].arrive()
            yield

        g0 = 0
        for seg, rb, t0, nt in self.segs:
            if group == (g0 & 1):
                yield from tile_step(g0, True)
                g = g0 + 2
            else:
                yield from self.named_sync(("all", r, quad), 4)
                g = g0 + 1
            while g < g0 + nt:
                yield from tile_step(g, False)
                g += 2
            yield from self.named_sync(("all", r, quad), 4)
            yield from self.wait("o_done", r, seg & 1, seg + 1)
            assert self.O["seg"] == seg and self.O.get("complete"), f"epilogue of segment {seg} reads O: {self.O}"
            self.O["readers"] += 1
            self.bars[("o_free", 0)].arrive()
            yield
            g0 += nt

    # ---- scheduler
    def run(self):
        roles = [self.producer(0), self.producer(1), self.mma()]
        for r in (0, 1):
            for group in (0, 1):
                for half in (0, 1):
                    for quad in range(4):
                        roles.append(self.softmax(r, group, half, quad))
        live = list(range(len(roles)))
        idle_rounds = 0
        steps = 0
        frozen, frozen_until = None, 0
        while live:
            steps += 1
            progressed = False
            # deliver some asynchronous completions: TMA in any order, commits strictly FIFO
            if self.async_events and self.rng.random() < 0.5:
                _, fn = self.async_events.pop(self.rng.randrange(len(self.async_events)))
                fn()
                progressed = True
            if self.commit_fifo and self.rng.random() < 0.4:
                self.commit_fifo.pop(0)()
                progressed = True
            # adversarial scheduling: every so often one role is frozen for a while (a warp that is descheduled between
            # two dependent steps is exactly what exposes a missing hand-over)
            if frozen_until <= steps and self.rng.random() < 0.02:
                frozen, frozen_until = self.rng.choice(live), steps + self.rng.randrange(50, 3000)
            candidates = [x for x in live if x != frozen or frozen_until <= steps] or live
            i = self.rng.choice(candidates)
            snapshot = self.state_hash()
            try:
                next(roles[i])
            except StopIteration:
                live.remove(i)
                progressed = True
            if self.state_hash() != snapshot:
                progressed = True
            if progressed:
                idle_rounds = 0
            else:
                idle_rounds += 1
                if idle_rounds > 20000 and not self.async_events and not self.commit_fifo:
                    raise RuntimeError(f"deadlock: {len(live)} roles blocked; bars: " +
                                       ", ".join(f"{b.name}={b.done}+{b.pending}" for b in self.bars.values() if b.pending))
                if idle_rounds > 20000:
                    # force delivery of outstanding completions
                    if self.commit_fifo:
                        self.commit_fifo.pop(0)()
                    elif self.async_events:
                        self.async_events.pop(0)[1]()
                    idle_rounds = 0
        assert not self.commit_fifo or all(True for _ in self.commit_fifo)
        return steps

    def state_hash(self):
        return (tuple((b.done, b.pending) for b in self.bars.values()), len(self.async_events), len(self.commit_fifo),
                tuple((n.gen, n.waiting) for n in self.named.values()), self.busy)


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(7)
    for t in range(trials):
        T = rng.choice([1, 2, 3, 5, 8, 13, 40])
        RB = rng.randint(1, 6)
        C = rng.choice([1, 2, 3, 5])
        W = RB * T
        if W < C:
            continue
        c = rng.randrange(C)
        begin, end = c * W // C, (c + 1) * W // C
        if end == begin:
            continue
        Sim(begin, end, T, rng).run()
    # the c3 shape of one cluster: 221 tiles crossing a row-block boundary
    Sim(221 * 2, 221 * 3, 512, rng).run()
    print(f"v8 protocol model: {trials} random work ranges + the c3 range passed (no deadlock, no stale phase, no broken hand-over)")


if __name__ == "__main__":
    main()
