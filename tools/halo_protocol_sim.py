"""Randomised discrete simulation of conv_halo.cu's mbarrier protocol.

Models the five roles of one CTA (halo producer, weight producer, MMA issuer,
4*SETS split warps) as coroutines over parity-waited mbarriers with arrival
counts, exactly in the order the kernel issues waits / arrives, and checks

  * no deadlock (every role finishes under any schedule),
  * no barrier receives more arrivals than its count in one phase,
  * every read sees the data it expects (halo slot holds the chunk being read,
    the weight stage and the A slot hold the (chunk, tap[, m]) being multiplied,
    nothing is overwritten before its reader is done).

It runs on the CPU in seconds and is how the "release a skipped chunk only after
its load was issued" rule in the split warps was found and verified, and how the
persistent (several tiles per CTA) version of the protocol was checked before it
first ran on a GPU.

    python tools/halo_protocol_sim.py [n_random_schedules]
"""
import itertools
import random
import sys


class Bar:
    def __init__(self, count, name):
        self.count, self.name = count, name
        self.pending, self.phases = count, 0

    def arrive(self):
        if self.pending <= 0:
            raise AssertionError(f'{self.name}: arrival underflow')
        self.pending -= 1
        if self.pending == 0:
            self.phases += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.phases & 1) != parity


class Sim:
    def __init__(self, sets, mt, T, chunks, nh, nb, rng, fixed=True, tiles=1):
        self.sets, self.mt, self.T, self.chunks, self.nh, self.nb = sets, mt, T, chunks, nh, nb
        self.tiles = tiles                     # tiles one (persistent) CTA runs through
        self.rng, self.fixed = rng, fixed
        self.halo_full = [Bar(1, f'halo_full{i}') for i in range(nh)]
        self.halo_empty = [Bar(4 * sets, f'halo_empty{i}') for i in range(nh)]
        self.b_full = [Bar(1, f'b_full{i}') for i in range(nb)]
        self.ring_empty = [Bar(1, f'ring_empty{i}') for i in range(nb)]
        self.a_full = [Bar(4, f'a_full{i}') for i in range(nb * mt)]
        self.accum = Bar(1, 'accum')
        self.acc_empty = Bar(4 * sets, 'acc_empty')
        self.acc_tile = None                   # tile whose sums sit in the accumulators
        self.acc_readers = 0
        self.halo_slot = [None] * nh           # chunk resident in each halo slot
        self.halo_readers = [0] * nh
        self.b_slot = [None] * nb
        self.a_slot = [[None] * 4 for _ in range(nb * mt)]   # per quadrant warp
        self.asyncq = []                       # TMA completions (any order)
        self.commitq = []                      # tcgen05.commit arrivals (retire in order)

    # every role is a generator yielding ('wait', bar, parity) or ('step',)
    def halo_producer(self):
        s, ph = 0, 1
        for c in range(self.chunks * self.tiles):      # global chunk index
            yield ('wait', self.halo_empty[s], ph)
            def land(s=s, c=c):
                assert self.halo_readers[s] == 0, f'halo slot {s} overwritten under a reader'
                self.halo_slot[s] = c
                self.halo_full[s].arrive()
            self.asyncq.append(land)
            yield ('step',)
            s += 1
            if s == self.nh:
                s, ph = 0, ph ^ 1

    def weight_producer(self):
        s, ph = 0, 1
        for g in range(self.chunks * self.T * self.tiles):
            yield ('wait', self.ring_empty[s], ph)
            def land(s=s, g=g):
                self.b_slot[s] = g
                self.b_full[s].arrive()
            self.asyncq.append(land)
            yield ('step',)
            s += 1
            if s == self.nb:
                s, ph = 0, ph ^ 1

    def mma(self):
        r, ph = 0, 0
        per_tile = self.chunks * self.T
        for g in range(per_tile * self.tiles):
            tile = g // per_tile
            if g % per_tile == 0:
                if tile > 0:
                    yield ('wait', self.acc_empty, (tile - 1) & 1)
                assert self.acc_readers == 0, 'accumulators overwritten under the epilogue'
                self.acc_tile = ('partial', tile)
            yield ('wait', self.b_full[r], ph)
            assert self.b_slot[r] == g, f'weights of {g}: stage holds {self.b_slot[r]}'
            for m in range(self.mt):
                yield ('wait', self.a_full[r * self.mt + m], ph)
                assert self.a_slot[r * self.mt + m] == [(g, m)] * 4, \
                    f'A of {(g, m)}: slot holds {self.a_slot[r * self.mt + m]}'
            def retire(r=r):
                self.ring_empty[r].arrive()
            self.commitq.append(retire)
            if g % per_tile == per_tile - 1:
                def done(tile=tile):
                    self.acc_tile = tile
                    self.accum.arrive()
                self.commitq.append(done)
            yield ('step',)
            r += 1
            if r == self.nb:
                r, ph = 0, ph ^ 1
        self.done_mma = True

    def split_warp(self, set_, quad):
        mt, T, nb, nh, chunks = self.mt, self.T, self.nb, self.nh, self.chunks
        st = dict(c=0, t=0, r=0, eph=1, hs=0, hph=0)
        gc0 = 0                                # chunks of the tiles before this one

        def step_tap():
            st['r'] += 1
            if st['r'] == nb:
                st['r'], st['eph'] = 0, st['eph'] ^ 1
            st['t'] += 1
            if st['t'] < T:
                return
            st['t'] = 0
            st['c'] += 1
            st['hs'] += 1
            if st['hs'] == nh:
                st['hs'], st['hph'] = 0, st['hph'] ^ 1

        def release(ch):
            g = gc0 + ch                       # the halo ring runs on across tiles
            slot = g % nh
            if self.fixed:
                yield ('wait', self.halo_full[slot], (g // nh) & 1)
            self.halo_empty[slot].arrive()

        reading = None

        def begin_read():
            nonlocal reading
            assert self.halo_slot[st['hs']] == gc0 + st['c'], \
                f"set {set_} reads chunk {gc0 + st['c']}: slot holds {self.halo_slot[st['hs']]}"
            reading = st['hs']
            self.halo_readers[reading] += 1

        def end_read():
            nonlocal reading
            if reading is not None:
                self.halo_readers[reading] -= 1
                reading = None

        m_cur = set_ % mt
        for _ in range(set_ // mt):
            step_tap()
        items = mt                             # epilogue items of one tile (one column group)
        for tile in range(self.tiles):
            released = 0
            if self.fixed:
                while released < min(st['c'], chunks):
                    yield from release(released)
                    released += 1
            if st['c'] < chunks:
                yield ('wait', self.halo_full[st['hs']], st['hph'])
                begin_read()
                yield ('step',)
                end_read()
            while st['c'] < chunks:
                g = (gc0 + st['c']) * T + st['t']
                slot = st['r'] * mt + m_cur
                ring_r, ring_ph = st['r'], st['eph']
                m_now = m_cur
                yield ('step',)                   # hi/lo split
                # advance; inside the same chunk the next row is fetched BEFORE the
                # barrier traffic (round 2)
                c_prev = st['c']
                mm = m_cur + self.sets
                taps = mm // mt
                m_cur = mm - taps * mt
                for _ in range(taps):
                    step_tap()
                same_chunk = st['c'] == c_prev
                if same_chunk:
                    begin_read()
                    yield ('step',)
                    end_read()
                yield ('wait', self.ring_empty[ring_r], ring_ph)
                self.a_slot[slot][quad] = (g, m_now)
                self.a_full[slot].arrive()        # published at once (round 2)
                if not same_chunk:
                    upto = min(st['c'], chunks)
                    while released < upto:
                        yield from release(released)
                        released += 1
                    if st['c'] < chunks:
                        yield ('wait', self.halo_full[st['hs']], st['hph'])
                        begin_read()
                        yield ('step',)
                        end_read()
            while released < chunks:
                yield from release(released)
                released += 1
            # epilogue of this tile: EVERY split warp waits for the accumulators (a
            # warp without an item must not run a tile ahead: its acc_empty arrival
            # would land in the previous phase); sets with an item read them
            yield ('wait', self.accum, tile & 1)
            if set_ < items:
                assert self.acc_tile == tile, f'epilogue of tile {tile} reads {self.acc_tile}'
                self.acc_readers += 1
                yield ('step',)
                self.acc_readers -= 1
            self.acc_empty.arrive()
            st['c'] -= chunks
            gc0 += chunks

    def run(self):
        roles = [self.halo_producer(), self.weight_producer(), self.mma()]
        roles += [self.split_warp(s, q) for s in range(self.sets) for q in range(4)]
        state = [None] * len(roles)            # pending wait of each role
        alive = set(range(len(roles)))
        for i in list(alive):
            try:
                state[i] = next(roles[i])
            except StopIteration:
                alive.discard(i)
        while alive:
            ready = [i for i in alive
                     if state[i][0] == 'step' or state[i][1].passed(state[i][2])]
            n_choices = len(ready) + len(self.asyncq) + (1 if self.commitq else 0)
            if n_choices == 0:
                waits = {i: (state[i][1].name, state[i][2], state[i][1].phases) for i in alive}
                raise AssertionError(f'deadlock: {waits}')
            k = self.rng.randrange(n_choices)
            if k >= len(ready) + len(self.asyncq):
                self.commitq.pop(0)()
                continue
            if k >= len(ready):
                self.asyncq.pop(k - len(ready))()
                continue
            i = ready[k]
            try:
                state[i] = next(roles[i])
            except StopIteration:
                alive.discard(i)
        for f in self.asyncq + self.commitq:
            f()


def planner_allows(sets, mt, T, chunks, nh, nb, tiles=1):
    """The ring-depth rules make_plan_uncached() enforces (conv_halo.cu)."""
    if T * chunks > 1 and nb < 2:
        return False
    if T * chunks > nb and nb < -(-sets // mt):
        return False
    # one-tile plans may fall back to ONE halo slot when two do not fit (stride-2 3x3x3 convs:
    # the next chunk's halo then loads after the last tap of the current one, not under it)
    return nh >= 1 if tiles == 1 else nh >= 2


def sweep(sets_list=(2, 3, 4), n=10, seed=0, fixed=True, only_allowed=True, tiles=1):
    """-> (schedules run, {config: first failure})."""
    rng = random.Random(seed)
    bad, cases = {}, 0
    for sets, mt, T, chunks, nb in itertools.product(sets_list, (1, 2), (1, 2, 9, 27),
                                                     (1, 2, 3, 5, 8), (2, 3, 4, 6)):
        for nh in ({1, min(chunks, 2), min(chunks, 3)} if tiles == 1 else {2, 3}):
            if only_allowed and not planner_allows(sets, mt, T, chunks, nh, nb, tiles):
                continue
            for _ in range(n):
                cases += 1
                try:
                    Sim(sets, mt, T, chunks, nh, nb, rng, fixed=fixed, tiles=tiles).run()
                except AssertionError as e:
                    bad.setdefault((sets, mt, T, chunks, nh, nb), str(e)[:200])
    return cases, bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    cases, bad = sweep(n=n)
    print(f'{cases} schedules, {len(bad)} failing configurations')
    for k, v in sorted(bad.items()):
        print(' sets,mt,T,chunks,nh,nb =', k, v)
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
