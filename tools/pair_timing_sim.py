"""Timing model of the CTA-pair gradient kernel (csrc/tc_pair.cu): two CTAs, each with an in-order TMA producer,
an in-order MMA issuer feeding an in-order tensor pipe, an epilogue (8 warps as one unit) and a W sender, coupled
through mbarriers.  It is a max-plus event simulation, not cycle accurate: it is used to compare SCHEDULES (issue
order, lag, ring size, buffer counts) under measured latencies before spending GPU time.

usage: python tools/pair_timing_sim.py
"""
import heapq
import itertools
from collections import defaultdict


class Sim:
    def __init__(self):
        self.t_done = defaultdict(list)      # barrier key -> completion time of phase 0, 1, 2, ...
        self.pending = defaultdict(int)
        self.count = {}
        self.waiters = defaultdict(list)
        self.heap, self.seq = [], itertools.count()
        self.roles = {}

    def bar(self, key, count):
        self.count[key] = count

    def arrive(self, key, t):
        self.pending[key] += 1
        if self.pending[key] == self.count[key]:
            self.pending[key] = 0
            self.t_done[key].append(t)
            ph = len(self.t_done[key]) - 1
            for (rid, want) in list(self.waiters[key]):
                if want == ph:
                    self.waiters[key].remove((rid, want))
                    r = self.roles[rid]
                    r["t"] = max(r["t"], t)
                    heapq.heappush(self.heap, (r["t"], next(self.seq), rid))

    def add_role(self, rid, gen):
        self.roles[rid] = {"gen": gen, "t": 0.0, "blocked": 0.0, "started": False}
        heapq.heappush(self.heap, (0.0, next(self.seq), rid))

    def run(self):
        while self.heap:
            _, _, rid = heapq.heappop(self.heap)
            r = self.roles[rid]
            try:
                while True:
                    if r["started"]:
                        act = r["gen"].send(r["t"])
                    else:
                        r["started"] = True
                        act = next(r["gen"])
                    kind = act[0]
                    if kind == "wait":          # ("wait", key, phase index)
                        _, key, ph = act
                        if ph < 0:
                            continue
                        if len(self.t_done[key]) > ph:
                            t0 = r["t"]
                            r["t"] = max(r["t"], self.t_done[key][ph]) + 20
                            r["blocked"] += r["t"] - t0 - 20
                            continue
                        r["wait_from"] = r["t"]
                        self.waiters[key].append((rid, ph))
                        break
                    elif kind == "delay":
                        r["t"] += act[1]
                    elif kind == "arrive":      # ("arrive", key, at_time or None)
                        self.arrive(act[1], r["t"] if len(act) < 3 else act[2])
                    else:
                        raise ValueError(kind)
            except StopIteration:
                pass
        stuck = {k: v for k, v in self.waiters.items() if v}
        return stuck


def simulate(nt=64, nslots=4, lag_own=2, lag_peer=2, L_tma=700.0, tma_bw=80.0, E=2500.0, X=2900.0, ack=400.0,
             wbufs=1, chunk_clk=256.0, v_pair_clk=512.0, kch=8, gch=4, issue_cost=40.0, a_stream=0, w_in_ring=False, verbose=False):
    """One item of nt column tiles on a CTA pair.  Returns (cycles per 2 tiles, tensor-pipe busy fraction)."""
    sim = Sim()
    CH = 16384.0
    for c in (0, 1):
        for s in range(nslots):
            sim.bar((c, "full", s), 1)
            sim.bar((c, "empty", s), 1)
        for b in (0, 1):
            sim.bar((c, "s_full", b), 1)
            sim.bar((c, "s_empty", b), 2)      # MMA2 commit + sender's TMEM read
            sim.bar((c, "g_full", b), 1)
        for w in range(wbufs):
            sim.bar((c, "w_full", w), 1)       # data landed in c's Wrecv[w]
            sim.bar((c, "w_empty", w), 1)      # c's peer consumed what c sent into the peer's Wrecv[w]
    pipe_free = [0.0, 0.0]
    pipe_busy = [0.0, 0.0]
    tma_free = [0.0, 0.0]
    end_time = [0.0, 0.0]

    def own(c, t):
        return (t & 1) == c

    def steps(c):
        """issue order: per step s: MMA1(s) if own; MMA2 of own tile s-lag_own; MMA2 of peer tile s-lag_peer."""
        seq = []
        for s in range(nt + max(lag_own, lag_peer) + 1):
            if s < nt and own(c, s):
                seq.append(("m1", s))
            to = s - lag_own
            if 0 <= to < nt and own(c, to):
                seq.append(("m2o", to))
            tp = s - lag_peer
            if 0 <= tp < nt and not own(c, tp):
                seq.append(("m2p", tp))
        return seq

    def producer(c):
        t = yield ("delay", 0)
        slot, uses = 0, defaultdict(int)
        wk = 0
        for kind, tile in steps(c):
            n = kch + a_stream if kind == "m1" else gch      # the last a_stream K-chunks of A are streamed too
            if kind == "m2p" and w_in_ring:                   # W arrives through L2: wait for the peer's "ready", then 2 chunks
                t = yield ("wait", (c, "w_full", wk % wbufs), wk // wbufs)
                wk += 1
                n += 2
            for _ in range(n):
                t = yield ("wait", (c, "empty", slot), uses[slot] - 1)
                uses[slot] += 1
                start = max(t + L_tma, tma_free[c])
                tma_free[c] = start + CH / tma_bw
                t = yield ("arrive", (c, "full", slot), tma_free[c])
                t = yield ("delay", 20)
                slot = (slot + 1) % nslots

    def mma(c):
        t = yield ("delay", 0)
        slot, uses = 0, defaultdict(int)
        k1 = k2 = kp = 0

        def run_op(t, dur):
            start = max(t, pipe_free[c])
            pipe_free[c] = start + dur
            pipe_busy[c] += dur
            return pipe_free[c]

        for kind, tile in steps(c):
            if kind == "m1":
                b = k1 & 1
                t = yield ("wait", (c, "s_empty", b), (k1 >> 1) - 1)
                for kc in range(kch):
                    used = []
                    for _ in range(2 if kc >= kch - a_stream else 1):
                        t = yield ("wait", (c, "full", slot), uses[slot])
                        uses[slot] += 1
                        used.append(slot)
                        slot = (slot + 1) % nslots
                    t = yield ("delay", issue_cost)
                    fin = run_op(t, chunk_clk)
                    for u in used:
                        t = yield ("arrive", (c, "empty", u), fin + 30)
                t = yield ("arrive", (c, "s_full", b), fin + 30)
                k1 += 1
            else:
                if kind == "m2o":
                    b = k2 & 1
                    t = yield ("wait", (c, "g_full", b), k2 >> 1)
                else:
                    w = kp % wbufs
                    if not w_in_ring:
                        t = yield ("wait", (c, "w_full", w), kp // wbufs)
                wslots = []
                if kind == "m2p" and w_in_ring:
                    for _ in range(2):
                        t = yield ("wait", (c, "full", slot), uses[slot])
                        uses[slot] += 1
                        wslots.append(slot)
                        slot = (slot + 1) % nslots
                for g in range(gch // 2):
                    for _ in range(2):
                        t = yield ("wait", (c, "full", slot), uses[slot])
                        uses[slot] += 1
                        slot = (slot + 1) % nslots
                    t = yield ("delay", issue_cost)
                    fin = run_op(t, v_pair_clk)
                    for d in (2, 1):
                        t = yield ("arrive", (c, "empty", (slot - d) % nslots), fin + 30)
                for u in wslots:
                    t = yield ("arrive", (c, "empty", u), fin + 30)
                if kind == "m2o":
                    t = yield ("arrive", (c, "s_empty", k2 & 1), fin + 30)
                    k2 += 1
                else:
                    t = yield ("arrive", (1 - c, "w_empty", kp % wbufs), fin + 30 + ack)
                    kp += 1
                end_time[c] = fin

    def epilogue(c):
        t = yield ("delay", 0)
        k = 0
        for tile in range(nt):
            if not own(c, tile):
                continue
            b = k & 1
            t = yield ("wait", (c, "s_full", b), k >> 1)
            t = yield ("delay", E)
            t = yield ("arrive", (c, "g_full", b))
            k += 1

    def sender(c):
        t = yield ("delay", 0)
        k = 0
        for tile in range(nt):
            if not own(c, tile):
                continue
            b = k & 1
            t = yield ("wait", (c, "g_full", b), k >> 1)
            t = yield ("delay", 150)
            t = yield ("arrive", (c, "s_empty", b))
            t = yield ("wait", (c, "w_empty", k % wbufs), k // wbufs - 1)
            t = yield ("delay", X)
            t = yield ("arrive", (1 - c, "w_full", k % wbufs))
            k += 1

    for c in (0, 1):
        sim.add_role((c, "tma"), producer(c))
        sim.add_role((c, "mma"), mma(c))
        sim.add_role((c, "epi"), epilogue(c))
        sim.add_role((c, "snd"), sender(c))
    stuck = sim.run()
    if stuck:
        return None, stuck
    total = max(end_time)
    return total / (nt / 2.0), sum(pipe_busy) / (2 * total)


if __name__ == "__main__":
    base = dict(nt=128)
    print("ideal: 4096 cycles per 2 tiles")
    for name, kw in [
        ("as built: 4 slots, lag 2/2", dict()),
        ("lag 2/3", dict(lag_peer=3)),
        ("lag 2/4", dict(lag_peer=4)),
        ("lag 2/5", dict(lag_peer=5)),
        ("lag 2/4, 2 W buffers", dict(lag_peer=4, wbufs=2)),
        ("lag 2/6, 2 W buffers", dict(lag_peer=6, wbufs=2)),
        ("6 slots, lag 2/4", dict(nslots=6, lag_peer=4)),
        ("6 slots, lag 2/4, 2 W buffers", dict(nslots=6, lag_peer=4, wbufs=2)),
        ("8 slots, lag 2/4, 2 W buffers", dict(nslots=8, lag_peer=4, wbufs=2)),
        ("4 slots, lag 2/4, X=1600", dict(lag_peer=4, X=1600)),
        ("4 slots, lag 2/4, E=1500", dict(lag_peer=4, E=1500)),
        ("4 slots, lag 2/4, L_tma=400", dict(lag_peer=4, L_tma=400)),
    ]:
        per, busy = simulate(**base, **kw)
        if per is None:
            print(f"{name:40s} DEADLOCK {busy}")
        else:
            print(f"{name:40s} {per:8.0f} cycles per 2 tiles   tensor busy {busy:.2f}")


def simulate_order(order="A", nt=128, nslots=6, lag_own=2, lag_peer=3, L_tma=1500.0, tma_bw=80.0, E=2300.0, X=3200.0,
                   ack=400.0, chunk_clk=330.0, v_pair_clk=600.0, kch=8, gch=4, issue_cost=40.0, a_stream=2, w_ring=False):
    """Like simulate() (single Wrecv, DSMEM hand-over) but the MMA/TMA issue order inside a step is a parameter:
    ops of the step of own tile t: MMA1 chunks c0..c7 of t, MMA2 groups o0,o1 of own tile t-lag_own, p0,p1 of peer tile
    t-lag_peer.  order: 'A' c0-7 o0 o1 p0 p1 (as built) | 'B' c0-3 o0 c4-7 o1 p0 p1 | 'C' c0-1 o0 c2-3 o1 c4-5 p0 c6-7 p1
    | 'D' c0-3 o0 o1 c4-7 p0 p1.
    w_ring: no dedicated landing buffer -- the peer's weight tile lands in two ordinary ring slots that the receiver's
    producer grants (in ring order) instead of filling them by TMA; they are released after the second MMA2-peer group."""
    sim = Sim()
    CH = 16384.0
    for c in (0, 1):
        for s in range(nslots):
            sim.bar((c, "full", s), 1)
            sim.bar((c, "empty", s), 1)
        for b in (0, 1):
            sim.bar((c, "s_full", b), 1)
            sim.bar((c, "s_empty", b), 2)
            sim.bar((c, "g_full", b), 1)
        sim.bar((c, "w_full", 0), 1)
        sim.bar((c, "w_empty", 0), 1)
        sim.bar((c, "w_grant", 0), 2)
    wslots = [dict(), dict()]
    pipe_free, pipe_busy, tma_free, end_time = [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]
    pat = {"A": "c0 c1 c2 c3 c4 c5 c6 c7 o0 o1 p0 p1", "B": "c0 c1 c2 c3 o0 c4 c5 c6 c7 o1 p0 p1",
           "C": "c0 c1 o0 c2 c3 o1 c4 c5 p0 c6 c7 p1", "D": "c0 c1 c2 c3 o0 o1 c4 c5 c6 c7 p0 p1",
           "E": "o0 c0 c1 c2 c3 o1 c4 c5 c6 c7 p0 p1"}[order].split()

    def own(c, t):
        return (t & 1) == c

    def ops(c):
        seq = []
        for s in range(c, nt + max(lag_own, lag_peer) + 2, 2):
            for o in pat:
                if o[0] == "c":
                    if s < nt:
                        seq.append(("c", s, int(o[1])))
                elif o[0] == "o":
                    if 0 <= s - lag_own < nt:
                        seq.append(("o", s - lag_own, int(o[1])))
                else:
                    if 0 <= s - lag_peer < nt:
                        seq.append(("p", s - lag_peer, int(o[1])))
        return seq

    def nslots_of(op):
        if op[0] == "c":
            return 2 if op[2] >= kch - a_stream else 1
        if w_ring and op[0] == "p" and op[2] == 0:
            return 4                                   # the weight tile's two slots, then the V pair
        return 2

    def producer(c):
        t = yield ("delay", 0)
        slot, uses = 0, defaultdict(int)
        for op in ops(c):
            for i in range(nslots_of(op)):
                t = yield ("wait", (c, "empty", slot), uses[slot] - 1)
                uses[slot] += 1
                if w_ring and op[0] == "p" and op[2] == 0 and i < 2:      # grant the slot to the peer's sender
                    wslots[c].setdefault(op[1] // 2, []).append(slot)
                    t = yield ("arrive", (c, "w_grant", 0), t + ack)
                    slot = (slot + 1) % nslots
                    continue
                start = max(t + L_tma, tma_free[c])
                tma_free[c] = start + CH / tma_bw
                t = yield ("arrive", (c, "full", slot), tma_free[c])
                t = yield ("delay", 20)
                slot = (slot + 1) % nslots

    def mma(c):
        t = yield ("delay", 0)
        slot, uses = 0, defaultdict(int)
        idx = {}            # own tile -> own index ; peer tile -> peer index
        no = npeer = 0
        for op in ops(c):
            kind, tile, sub = op
            if kind == "c":
                k1 = tile // 2
                if sub == 0:
                    t = yield ("wait", (c, "s_empty", k1 & 1), (k1 >> 1) - 1)
            elif kind == "o":
                k2 = tile // 2
                if sub == 0:
                    t = yield ("wait", (c, "g_full", k2 & 1), k2 >> 1)
            else:
                kp = tile // 2
                if sub == 0 and not w_ring:
                    t = yield ("wait", (c, "w_full", 0), kp)
            used = []
            for _ in range(nslots_of(op)):
                t = yield ("wait", (c, "full", slot), uses[slot])
                uses[slot] += 1
                used.append(slot)
                slot = (slot + 1) % nslots
            t = yield ("delay", issue_cost)
            start = max(t, pipe_free[c])
            dur = chunk_clk if kind == "c" else v_pair_clk
            pipe_free[c] = start + dur
            pipe_busy[c] += dur
            fin = pipe_free[c]
            if w_ring and kind == "p":
                if sub == 0:
                    wpend, used = used[:2], used[2:]
                else:
                    used = used + wpend
            for u in used:
                t = yield ("arrive", (c, "empty", u), fin + 30)
            if kind == "c" and sub == kch - 1:
                t = yield ("arrive", (c, "s_full", (tile // 2) & 1), fin + 30)
            if kind == "o" and sub == 1:
                t = yield ("arrive", (c, "s_empty", (tile // 2) & 1), fin + 30)
            if kind == "p" and sub == 1 and not w_ring:
                t = yield ("arrive", (1 - c, "w_empty", 0), fin + 30 + ack)
            end_time[c] = max(end_time[c], fin)

    def epilogue(c):
        t = yield ("delay", 0)
        for k, tile in enumerate(range(c, nt, 2)):
            t = yield ("wait", (c, "s_full", k & 1), k >> 1)
            t = yield ("delay", E)
            t = yield ("arrive", (c, "g_full", k & 1))

    def sender(c):
        t = yield ("delay", 0)
        for k, tile in enumerate(range(c, nt, 2)):
            t = yield ("wait", (c, "g_full", k & 1), k >> 1)
            t = yield ("delay", 150)
            t = yield ("arrive", (c, "s_empty", k & 1))
            if w_ring:
                t = yield ("wait", (1 - c, "w_grant", 0), k)
                t = yield ("delay", X)
                for sl in wslots[1 - c][k]:
                    t = yield ("arrive", (1 - c, "full", sl))
                continue
            t = yield ("wait", (c, "w_empty", 0), k - 1)
            t = yield ("delay", X)
            t = yield ("arrive", (1 - c, "w_full", 0))

    for c in (0, 1):
        sim.add_role((c, "tma"), producer(c))
        sim.add_role((c, "mma"), mma(c))
        sim.add_role((c, "epi"), epilogue(c))
        sim.add_role((c, "snd"), sender(c))
    stuck = sim.run()
    if stuck:
        return None, stuck
    total = max(end_time)
    return total / (nt / 2.0), sum(pipe_busy) / (2 * total)
