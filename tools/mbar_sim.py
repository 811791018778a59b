"""Randomised protocol check of the mbarrier pipelines of the staged tcgen05 attention kernels.

The kernels' correctness on hardware has two halves: the arithmetic (mirrors the validated mma.sync kernels and is
covered by the GPU parity tests) and the producer / MMA / consumer PROTOCOL — which barrier is waited on with which
parity, who arrives how often, when a shared-memory or TMEM buffer may be overwritten.  A protocol slip shows up
as a hang (the kernels' bounded waits turn it into a trap) or as silent corruption.  This tool replays the
barrier traffic of

    attention_bwd_tc.cu   (warp 0 producer [+31 staging lanes in the dKV pass], warp 1 MMA issuer, 256 elementwise threads)
    attention_tc_long.cu  (producer, MMA issuer, two softmax groups of 128 threads)

as cooperating coroutines under a random scheduler, with asynchronous completions (TMA transactions and
tcgen05.commit arrivals are delivered after random delays), and checks on every run that (a) nobody deadlocks and
(b) every buffer is written only when free and read only while it holds the block the reader expects.

    python tools/mbar_sim.py            # a few thousand random schedules per configuration
"""
import random
import sys


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done = name, count, count, 0     # done = completed phases

    def arrive(self, n=1):
        assert self.pending >= n, f"{self.name}: more arrivals than the barrier expects in one phase"
        self.pending -= n
        if self.pending == 0:
            self.pending = self.count
            self.done += 1

    def ready(self, parity):
        return (self.done & 1) != parity          # try_wait.parity P succeeds iff the current phase's parity != P


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.bars = {}
        self.async_q = []                          # (due_tick, fn) completions in flight
        self.tick = 0
        self.buf = {}                              # buffer name -> content tag or None (free)

    def bar(self, name, count):
        self.bars[name] = Barrier(name, count)

    def later(self, fn, max_delay=6):
        self.async_q.append((self.tick + self.rng.randint(0, max_delay), fn))

    # buffer discipline
    def write(self, name, tag):
        assert self.buf.get(name) is None, f"{name} overwritten while it still holds {self.buf[name]} (writing {tag})"
        self.buf[name] = tag

    def read(self, name, tag):
        assert self.buf.get(name) == tag, f"{name}: expected {tag}, holds {self.buf.get(name)}"

    def free(self, name, tag):
        self.read(name, tag)
        self.buf[name] = None

    def run(self, agents, max_ticks=200000):
        live = {k: g for k, g in agents.items()}
        waiting = {}
        while live:
            self.tick += 1
            assert self.tick < max_ticks, "tick limit"
            due = [x for x in self.async_q if x[0] <= self.tick]
            self.async_q = [x for x in self.async_q if x[0] > self.tick]
            self.rng.shuffle(due)
            for _, fn in due:
                fn()
            runnable = []
            for k in live:
                w = waiting.get(k)
                if w is None or self.bars[w[0]].ready(w[1]):
                    runnable.append(k)
            if not runnable:
                if self.async_q:
                    continue
                raise AssertionError("DEADLOCK: " + ", ".join(f"{k} waits {waiting[k]} (done={self.bars[waiting[k][0]].done})"
                                                               for k in live))
            k = self.rng.choice(runnable)
            waiting.pop(k, None)
            try:
                op = next(live[k])
            except StopIteration:
                del live[k]
                continue
            if op is not None:
                waiting[k] = op                     # ("barrier name", parity): checked before the agent resumes
        assert not self.async_q or True


# ---------------------------------------------------------------------------------------------------------------
# attention_bwd_tc.cu
# ---------------------------------------------------------------------------------------------------------------
def sim_bwd(seed, items, nblk, dkv, stages=4):
    s = Sim(seed)
    for i in range(2):
        s.bar(f"r_full{i}", 1), s.bar(f"r_empty{i}", 1), s.bar(f"t_full{i}", 1), s.bar(f"t_empty{i}", 256)
        s.bar(f"e_full{i}", 256), s.bar(f"e_empty{i}", 1)
    for i in range(stages):
        s.bar(f"x_full{i}", 33 if dkv else 1), s.bar(f"x_empty{i}", 1)
    s.bar("acc_full", 1), s.bar("acc_empty", 256)
    B = s.bars

    def producer_lane0():
        c = 0
        for il in range(items):
            rs = il & 1
            yield (f"r_empty{rs}", ((il >> 1) & 1) ^ 1)
            s.write(f"R{rs}", il)                                   # TMA may start writing now
            s.later(lambda rs=rs: B[f"r_full{rs}"].arrive())        # expect_tx arrival + bytes landed
            for j in range(nblk):
                st = c % stages
                yield (f"x_empty{st}", ((c // stages) & 1) ^ 1)
                s.write(f"X{st}", c)
                s.later(lambda st=st: B[f"x_full{st}"].arrive())
                if dkv:
                    s.write(f"V{st}_l0", c)
                    B[f"x_full{st}"].arrive()                       # lane 0 also stages its two vector entries
                c += 1
                yield None

    def producer_lanes():                                           # lanes 1..31 of warp 0 (dKV pass only)
        c = 0
        for il in range(items):
            for j in range(nblk):
                st = c % stages
                yield (f"x_empty{st}", ((c // stages) & 1) ^ 1)
                s.write(f"V{st}_l", c)
                B[f"x_full{st}"].arrive(31)
                c += 1
                yield None

    def mma():
        def issue_t(c, rs):
            buf, st = c & 1, c % stages
            yield (f"x_full{st}", (c // stages) & 1)
            yield (f"t_empty{buf}", ((c >> 1) & 1) ^ 1)
            s.read(f"X{st}", c)
            s.read(f"R{rs}", c_item[c])
            s.write(f"T{buf}", c)
            s.later(lambda buf=buf: B[f"t_full{buf}"].arrive())
        c0 = 0
        for il in range(items):
            rs = il & 1
            yield (f"r_full{rs}", (il >> 1) & 1)
            yield from issue_t(c0, rs)
            for j in range(nblk):
                c = c0 + j
                buf, st = c & 1, c % stages
                if j + 1 < nblk:
                    yield from issue_t(c + 1, rs)
                yield (f"e_full{buf}", (c >> 1) & 1)
                if j == 0:
                    yield ("acc_empty", (il & 1) ^ 1)
                    s.write("ACC", il)
                s.read(f"E{buf}", c)
                s.read(f"X{st}", c)
                s.read("ACC", il)

                def done(buf=buf, st=st, c=c):                       # commits arrive once the MMAs have completed
                    s.free(f"E{buf}", c)
                    B[f"e_empty{buf}"].arrive()
                    s.free(f"X{st}", c)
                    if dkv:
                        s.free(f"V{st}_l0", c), s.free(f"V{st}_l", c)
                    B[f"x_empty{st}"].arrive()
                s.later(done)
                yield None
            # tcgen05.commit tracks every earlier MMA: deliver these after the block commits above
            def item_done(rs=rs, il=il):
                B["acc_full"].arrive()
                s.free(f"R{rs}", il)
                B[f"r_empty{rs}"].arrive()
            s.later(item_done, max_delay=20)
            # (commit ordering: hardware delivers commits in issue order; model that by draining the queue first)
            while any(fn.__name__ == "done" for _, fn in s.async_q):
                yield None
            c0 += nblk

    c_item = {}
    cc = 0
    for il in range(items):
        for j in range(nblk):
            c_item[cc] = il
            cc += 1

    def elementwise():
        c0 = 0
        for il in range(items):
            for j in range(nblk):
                c = c0 + j
                buf, st = c & 1, c % stages
                yield (f"t_full{buf}", (c >> 1) & 1)
                s.free(f"T{buf}", c)
                B[f"t_empty{buf}"].arrive(256)
                if dkv:
                    s.read(f"V{st}_l0", c), s.read(f"V{st}_l", c)
                yield (f"e_empty{buf}", ((c >> 1) & 1) ^ 1)
                s.write(f"E{buf}", c)
                B[f"e_full{buf}"].arrive(256)
                yield None
            yield ("acc_full", il & 1)
            s.free("ACC", il)
            B["acc_empty"].arrive(256)
            c0 += nblk

    agents = {"producer": producer_lane0(), "mma": mma(), "elementwise": elementwise()}
    if dkv:
        agents["lanes"] = producer_lanes()
    s.run(agents)


# ---------------------------------------------------------------------------------------------------------------
# attention_tc_long.cu
# ---------------------------------------------------------------------------------------------------------------
def sim_fwd_long(seed, items, nkb, stages=3):
    s = Sim(seed)
    for g in range(2):
        for q in range(2):
            s.bar(f"q_full{q}{g}", 1), s.bar(f"q_empty{q}{g}", 1)
        s.bar(f"s_full{g}", 1), s.bar(f"s_empty{g}", 128), s.bar(f"p_full{g}", 128), s.bar(f"p_empty{g}", 1)
        s.bar(f"o_full{g}", 1), s.bar(f"o_empty{g}", 128)
    for i in range(stages):
        s.bar(f"kv_full{i}", 1), s.bar(f"kv_empty{i}", 1)
    B = s.bars

    def producer():
        c = 0
        for il in range(items):
            qs = il & 1
            for g in range(2):
                yield (f"q_empty{qs}{g}", ((il >> 1) & 1) ^ 1)
                s.write(f"Q{qs}{g}", il)
                s.later(lambda qs=qs, g=g: B[f"q_full{qs}{g}"].arrive())
            for j in range(nkb):
                st = c % stages
                yield (f"kv_empty{st}", ((c // stages) & 1) ^ 1)
                s.write(f"KV{st}", c)
                s.later(lambda st=st: B[f"kv_full{st}"].arrive())
                c += 1
                yield None

    def mma():
        def issue_s(g, c, qs, il):
            yield (f"s_empty{g}", (c & 1) ^ 1)
            s.read(f"Q{qs}{g}", il)
            s.read(f"KV{c % stages}", c)
            s.write(f"S{g}", c)
            s.later(lambda g=g: B[f"s_full{g}"].arrive())
        c0 = 0
        for il in range(items):
            qs = il & 1
            yield (f"q_full{qs}0", (il >> 1) & 1)
            yield (f"q_full{qs}1", (il >> 1) & 1)
            yield (f"kv_full{c0 % stages}", (c0 // stages) & 1)
            yield from issue_s(0, c0, qs, il)
            yield from issue_s(1, c0, qs, il)
            for j in range(nkb):
                c = c0 + j
                for g in range(2):
                    yield (f"p_full{g}", c & 1)
                    yield (f"o_empty{g}", (c & 1) ^ 1)
                    s.read(f"P{g}", c)
                    s.read(f"KV{c % stages}", c)
                    s.write(f"O{g}", c)

                    def done(g=g, c=c, last=(g == 1)):
                        B[f"o_full{g}"].arrive()
                        s.free(f"P{g}", c)
                        B[f"p_empty{g}"].arrive()
                        if last:
                            s.free(f"KV{c % stages}", c)
                            B[f"kv_empty{c % stages}"].arrive()
                    s.later(done)
                    if j + 1 < nkb:
                        if g == 0:
                            yield (f"kv_full{(c + 1) % stages}", ((c + 1) // stages) & 1)
                        yield from issue_s(g, c + 1, qs, il)
                    yield None
            while any(fn.__name__ == "done" for _, fn in s.async_q):     # commits complete in issue order
                yield None

            def item_done(qs=qs, il=il):
                for g in range(2):
                    s.free(f"Q{qs}{g}", il)
                    B[f"q_empty{qs}{g}"].arrive()
            s.later(item_done)
            c0 += nkb

    def softmax(g):
        c0 = 0
        for il in range(items):
            for j in range(nkb):
                c = c0 + j
                yield (f"s_full{g}", c & 1)
                s.read(f"S{g}", c)
                yield (f"p_empty{g}", (c & 1) ^ 1)
                s.write(f"P{g}", c)
                s.free(f"S{g}", c)
                B[f"p_full{g}"].arrive(128)
                B[f"s_empty{g}"].arrive(128)
                if j > 0:
                    yield (f"o_full{g}", (c - 1) & 1)
                    s.free(f"O{g}", c - 1)
                    B[f"o_empty{g}"].arrive(128)
                yield None
            c = c0 + nkb - 1
            yield (f"o_full{g}", c & 1)
            s.free(f"O{g}", c)
            B[f"o_empty{g}"].arrive(128)
            c0 += nkb

    s.run({"producer": producer(), "mma": mma(), "A": softmax(0), "B": softmax(1)})


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    n = 0
    for items, nblk in [(1, 1), (1, 2), (2, 3), (3, 5), (4, 9), (5, 10), (2, 65)]:
        for dkv in (False, True):
            for seed in range(runs):
                sim_bwd(seed, items, nblk, dkv)
                n += 1
    for items, nkb in [(1, 1), (1, 2), (2, 3), (3, 5), (5, 4), (2, 33)]:
        for seed in range(runs):
            sim_fwd_long(seed, items, nkb)
            n += 1
    print(f"{n} random schedules: no deadlock, no buffer-discipline violation")


if __name__ == "__main__":
    main()
