"""Discrete model of the k_tc_pass mbarrier protocol (producer / MMA1 / MMA2 / epilogue roles), to find
deadlocks on the CPU before spending GPU time.  Each role is a generator that yields ("wait", barrier, parity)
or ("arrive", barrier) / ("commit", barrier) steps; commits complete only after every earlier MMA of the same
issuer has "executed" (the tensor pipe is modelled as a FIFO that retires ops whose operands are ready)."""
import sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):            # try_wait.parity semantics
        return (self.phase & 1) != parity

    def check(self, parity, who, name, problems):
        """A passing parity wait must not be an alias: the waiter may be at most one phase behind/ahead.
        Every waiter of a barrier observes every phase in order in a correct protocol, so the number of passed
        waits (per barrier, all waiters) must equal the number of completed phases it has consumed."""
        self.waits = getattr(self, "waits", 0) + 1
        if self.waits > self.phase + (1 if False else 0) and parity == ((self.waits - 1) & 1) and self.phase < self.waits:
            problems.append(f"ALIAS: {who} passed wait #{self.waits} on {name} with only {self.phase} phases complete")


def simulate(nps, kch, gch, nt_list, split_issuers, a_stat=True, verbose=False, skip_waits=False):
    KG = (kch + 1) // 2
    VG = (gch + 1) // 2
    bars = {("full", s): Bar(1) for s in range(nps)}
    bars.update({("empty", s): Bar(1) for s in range(nps)})
    for b in (0, 1):
        bars[("s_full", b)] = Bar(1)
        bars[("s_empty", b)] = Bar(1)     # TS mode: MMA2 commit
    bars["g_full"] = Bar(1)               # (8 epilogue warps modelled as one)
    bars["out_full"], bars["out_empty"] = Bar(1), Bar(1)
    bars["a_full"], bars["a_empty"] = Bar(1), Bar(1)

    class Ring:
        def __init__(self, bits):
            self.slot, self.bits = 0, bits

        def take(self):
            s = self.slot
            self.slot = 0 if self.slot + 1 == nps else self.slot + 1
            return s

        def par(self, s):
            p = (self.bits >> s) & 1
            self.bits ^= 1 << s
            return p

    def producer():
        ring, a_par = Ring((1 << 32) - 1), 1
        for nt in nt_list:
            yield ("wait", "a_empty", a_par); a_par ^= 1
            yield ("arrive", "a_full")
            def load():
                s = ring.take()
                yield ("wait", ("empty", s), ring.par(s))
                yield ("arrive", ("full", s))
            for t in range(nt):
                for _ in range(KG):
                    yield from load()
                if t >= 1:
                    for _ in range(VG):
                        yield from load()
            for _ in range(VG):
                yield from load()

    def issuer(do1, do2, name, skip_waits=False):
        ring, a_par, oe_par, gt1, gt2 = Ring(0), 0, 1, 0, 0
        uses = [0] * nps
        def need(s):
            uses[s] += 1
            return uses[s]
        def skip(n):
            for _ in range(n):
                s = ring.take(); p = ring.par(s); nd = need(s)
                if skip_waits:
                    yield ("wait", ("full", s), p, nd)
        for nt in nt_list:
            if do1:
                yield ("wait", "a_full", a_par); a_par ^= 1
            def mma1(last):
                nonlocal gt1
                b = gt1 & 1
                yield ("wait", ("s_empty", b), ((gt1 >> 1) & 1) ^ 1)
                for _ in range(KG):
                    s = ring.take()
                    yield ("wait", ("full", s), ring.par(s), need(s))
                    yield ("mma", name)
                    yield ("commit", ("empty", s), name)
                yield ("commit", ("s_full", b), name)
                if last:
                    yield ("commit", "a_empty", name)
                gt1 += 1
            def mma2(first, last):
                nonlocal gt2, oe_par
                b = gt2 & 1
                yield ("wait", "g_full", gt2 & 1)
                if first:
                    yield ("wait", "out_empty", oe_par); oe_par ^= 1
                for g in range(VG):
                    s = ring.take()
                    yield ("wait", ("full", s), ring.par(s), need(s))
                    yield ("mma", name)
                    yield ("commit", ("empty", s), name)
                    if g == VG - 1:
                        yield ("commit", ("s_empty", b), name)
                        if last:
                            yield ("commit", "out_full", name)
                gt2 += 1
            for t in range(nt):
                if do1: yield from mma1(t == nt - 1)
                else: yield from skip(KG)
                if t >= 1:
                    if do2: yield from mma2(t == 1, False)
                    else: yield from skip(VG)
            if do2: yield from mma2(nt == 1, True)
            else: yield from skip(VG)

    def epilogue():
        gt, item = 0, 0
        for nt in nt_list:
            for t in range(nt):
                b = gt & 1
                yield ("wait", ("s_full", b), (gt >> 1) & 1)
                yield ("arrive", "g_full")
                gt += 1
            yield ("wait", "out_full", item & 1)
            yield ("arrive", "out_empty")
            item += 1

    roles = {"producer": producer(), "epilogue": epilogue()}
    if split_issuers:
        roles["mma1"] = issuer(True, False, "mma1", skip_waits)
        roles["mma2"] = issuer(False, True, "mma2", skip_waits)
    else:
        roles["mma"] = issuer(True, True, "mma")
    pending = {k: None for k in roles}          # current blocked step
    problems = []
    # tensor pipe: FIFO of ("mma"|"commit", ...) in issue order; everything retires in order
    progress = True
    steps = 0
    live = set(roles)
    while live and progress:
        progress = False
        for k in list(live):
            for _once in (0,):                 # fair scheduling: ONE step per role per round
                st = pending[k]
                if st is None:
                    try:
                        st = next(roles[k])
                    except StopIteration:
                        live.discard(k)
                        break
                if st[0] == "wait":
                    if bars[st[1]].done(st[2]):
                        if len(st) > 3:      # ("wait", bar, parity, needed_phase_count)
                            if bars[st[1]].phase < st[3]:
                                problems.append(f"ALIAS: {k} passed {st[1]} needing {st[3]} phases, only {bars[st[1]].phase} complete")
                        pending[k] = None; progress = True; steps += 1
                        break
                    pending[k] = st
                    break
                if st[0] in ("arrive", "commit"):
                    bars[st[1]].arrive()        # (async completion modelled as immediate: operands were waited for)
                pending[k] = None; progress = True; steps += 1
    if problems:
        print(f"nps={nps} kch={kch} gch={gch} nt={nt_list} split={split_issuers}:", problems[0], f"(+{len(problems) - 1} more)")
        return False
    if live:
        print(f"DEADLOCK nps={nps} kch={kch} gch={gch} nt={nt_list} split={split_issuers}:",
              {k: pending[k] for k in live})
        return False
    return True


if __name__ == "__main__":
    ok = True
    for skip_waits in (False, True):
        print(f"--- split issuers, skipped slots {'observed (wait)' if skip_waits else 'not observed'}")
        for nps, kch, gch in ((3, 8, 4), (2, 8, 4), (3, 3, 3), (5, 2, 2), (6, 1, 1), (7, 12, 4)):
            for nt_list in ([1], [2], [3, 1], [128] * 3, [5, 4, 7]):
                r = simulate(nps, kch, gch, nt_list, True, skip_waits=skip_waits)
                if skip_waits:
                    ok &= r
    print("protocol", "OK" if ok else "HAS DEADLOCKS")
    sys.exit(0 if ok else 1)
