"""Executable statements of three arguments the CUDA build path relies on (csrc/gx_join.cu,
DESIGN.md §3.1).  They restate the device algorithms in numpy/Python and check them against the
obvious sequential definition, so the reasoning is tested where no GPU is available; the kernels
themselves are tested against the oracle by tests/test_gpu_parity.py.

1. A linear-probing table filled in slot order has the closed form the fill kernel uses.
2. The warp-wide lower-bound search returns the first row of every sub-table on key-ordered data.
3. 31 stored bits identify a key inside its sub-table, and the 16-byte slot can be rebuilt from
   the compact one and its position.
4. The window gx_k_runjoin_seg's producer stages for a chunk of key-ordered probe rows contains the
   home slot group of every key of the chunk, and windows of consecutive chunks tile the table.
5. gx_k_runjoin_tma's branch-free fold (select chains + predicated stores + one continuation add per lane)
   produces the run list of the sequential definition, and the "newest entries first" probe rounds of its
   carry variant visit every run exactly once, always with 32 lanes except for the final flush.
6. The full/empty mbarrier protocol of gx_k_runjoin_seg's ring (one producer, 31 consumer warps, phase parities
   derived from a per-buffer use count) neither deadlocks nor lets a buffer be refilled while a consumer still
   reads it, under arbitrary interleavings; gx_k_runjoin_tma's per-warp barrier is the one-consumer special case.
"""
import numpy as np
import pytest

SUB = 2048
LOG = 11


# ------------------------------------------------------------------ 1. closed-form fill
def sequential_fill(home_slots, nslots):
    """Insert rows in (home slot, arrival) order with linear probing, no wrap handling needed
    as long as nothing runs past the end."""
    tab = -np.ones(nslots, np.int64)
    for row in np.argsort(home_slots, kind="stable"):
        s = home_slots[row]
        while tab[s] >= 0:
            s += 1
        tab[s] = row
    return tab


def closed_form_positions(home_slots, nslots):
    """position = home + e[home] + rank among the rows of that home slot, with
    e[s] = P[s] - min_{j<=s} P[j],  P[s] = sum_{t<s} (c[t] - 1)   (gx_subtable_build)."""
    c = np.bincount(home_slots, minlength=nslots)
    P = np.concatenate([[0], np.cumsum(c - 1)[:-1]])
    e = P - np.minimum.accumulate(P)
    rank = np.zeros(len(home_slots), np.int64)
    seen = {}
    for i in np.argsort(home_slots, kind="stable"):
        s = home_slots[i]
        rank[i] = seen.get(s, 0)
        seen[s] = rank[i] + 1
    return home_slots + e[home_slots] + rank


@pytest.mark.parametrize("load,seed", [(0.3, 1), (0.56, 2), (0.75, 3), (0.56, 4)])
def test_closed_form_equals_sequential_linear_probing(load, seed):
    rng = np.random.default_rng(seed)
    nslots = 4096
    n = int(load * nslots)
    # homes are even (pair buckets) and kept away from the end so the sequential reference needs no wrap
    home = (rng.integers(0, nslots - 600, n) & ~1).astype(np.int64)
    if seed == 4:                                   # a pile-up: 200 rows on one slot
        home[:200] = 1000
    pos = closed_form_positions(home, nslots)
    assert len(np.unique(pos)) == n                  # conflict-free
    tab = sequential_fill(home, nslots)
    for row in range(n):
        # the same SET of slots is occupied, and every key is reachable: no hole between home and position
        assert tab[pos[row]] >= 0
        assert (tab[home[row]:pos[row] + 1] >= 0).all()
    assert set(np.flatnonzero(tab >= 0)) == set(pos)


# ------------------------------------------------------------------ 2. lower-bound search
def make_slot_fn(kmin, krange, nslots):
    ratio = (nslots - SUB) / krange
    sh = 0
    while sh < 62 and ratio * (1 << (sh + 1)) < 4294967295.0:
        sh += 1
    M = int(ratio * (1 << sh))

    def slot(k):
        d = (k - kmin) & 0xFFFFFFFF
        s = ((d * M) >> sh) & (nslots - 1)
        h = ((d * 0x9E3779B9) & 0xFFFFFFFF) >> 24
        return (s ^ (h & 31)) & ~3
    return slot


def warp_lower_bound(keys, slot, s, nsub, rows_per_sub):
    """sorted_lower_bound() with the 32 lanes written out as a loop."""
    nrows = len(keys)
    if s <= 0:
        return 0
    if s >= nsub:
        return nrows

    def pred(r):
        if r >= nrows:
            return True
        if r < 0:
            return False
        return (slot(int(keys[r])) >> LOG) >= s

    def ballot(f):
        return sum(1 << lane for lane in range(32) if f(lane))

    def ffs(b):
        return (b & -b).bit_length()

    g = min(int(s * rows_per_sub), nrows)
    pos = g - 16
    b = ballot(lambda lane: pred(pos + lane))
    if b not in (0, 0xFFFFFFFF):
        return pos + ffs(b) - 1
    lo, hi, e = 0, nrows, 256
    while True:
        cl, ch = max(g - e, 0), min(g + e, nrows)
        b = ballot(lambda lane: pred(cl - 1) if lane == 0 else (pred(ch) if lane == 1 else False))
        lo_ok, hi_ok = cl == 0 or not b & 1, ch == nrows or bool(b & 2)
        if lo_ok:
            lo = cl
        if hi_ok:
            hi = ch
        if (lo_ok and hi_ok) or (cl == 0 and ch == nrows):
            break
        e *= 16
    for _ in range(16):
        if hi - lo <= 32:
            break
        step = (hi - lo + 31) // 32
        b = ballot(lambda lane: True if lo + lane * step >= hi else pred(lo + lane * step))
        if b == 0:
            lo = lo + 31 * step + 1
            continue
        f = ffs(b) - 1
        hi = min(lo + f * step, hi)
        if f > 0:
            lo = lo + (f - 1) * step + 1
    b = ballot(lambda lane: True if lo + lane >= hi else pred(lo + lane))
    return lo + ffs(b) - 1 if b else hi


def key_sets():
    n = 120_000
    i = np.arange(n, dtype=np.int64)
    rng = np.random.default_rng(11)
    yield "tpch order keys (8 of every 32)", ((i >> 3) << 5 | (i & 7)) + 1
    yield "pseudo-random quarter of a key space (one of four datanodes)", np.sort(rng.choice(np.arange(1, 4 * n), n, replace=False)).astype(np.int64)
    yield "dense", np.arange(5, 5 + n, dtype=np.int64)
    yield "two clusters (the guess is far off)", np.sort(np.concatenate([rng.integers(0, 10**6, n // 2), rng.integers(3 * 10**6, 31 * 10**5, n // 2)])).astype(np.int64)


@pytest.mark.parametrize("name,keys", list(key_sets()), ids=[k[0].split(" (")[0] for k in key_sets()])
def test_lower_bound_search_finds_every_sub_table_start(name, keys):
    nrows = len(keys)
    nslots = 1
    while nslots < nrows + nrows // 2 + 16:
        nslots *= 2
    nsub = nslots // SUB
    slot = make_slot_fn(int(keys[0]), float(int(keys[-1]) - int(keys[0]) + 1), nslots)
    subs = np.array([slot(int(k)) >> LOG for k in keys])
    assert (np.diff(subs) >= 0).all()                # the slot function keeps key order at sub-table granularity
    rows_per_sub = nrows * SUB / (nslots - SUB)
    for s in range(nsub + 1):
        assert warp_lower_bound(keys, slot, s, nsub, rows_per_sub) == int(np.searchsorted(subs, s, side="left")), (name, s)


# ------------------------------------------------------------------ 3. compact slots
def test_31_bits_identify_a_key_inside_its_sub_table_and_rebuild_it():
    rng = np.random.default_rng(5)
    n = 100_000
    # a span above 2^32: what 8 datanodes at SF100 each see
    keys = np.sort(rng.choice(np.arange(1, 40 * n, dtype=np.int64) * 1200, n, replace=False))
    kmin, kmax = int(keys[0]), int(keys[-1])
    assert kmax - kmin > 2**32
    nslots = 1
    while nslots < n + n // 2 + 16:
        nslots *= 2
    keys_per_slot = (kmax - kmin + 1) / (nslots - SUB)
    assert keys_per_slot * SUB < 2**28               # the host's eligibility test
    home = ((keys - kmin) / keys_per_slot).astype(np.int64)          # order-preserving interpolation
    d = (((keys - kmin) << 1) | 1) & 0xFFFFFFFF                     # GX_CSLOT_D
    assert (d != 0).all()
    sub = home >> LOG
    for s in np.unique(sub)[:200]:
        ds = d[sub == s]
        assert len(np.unique(ds)) == len(ds)         # no two keys of a sub-table share their 31 bits
    # gx_k_expand_slots: the entry sits within a sub-table of its home; rebuild the key from d and the position
    displaced = np.clip(home + rng.integers(-SUB + 1, SUB, n), 0, nslots - 1)
    est = (displaced * keys_per_slot).astype(np.int64)
    off = (est & ~0x7FFFFFFF) | (d >> 1)
    off = np.where(off - est > 0x40000000, off - 0x80000000, off)
    off = np.where(est - off > 0x40000000, off + 0x80000000, off)
    np.testing.assert_array_equal(kmin + off, keys)


# ------------------------------------------------------------------ 4. the staged table window of a chunk
CHUNK = 31 * 128          # GX_SEG_CW consumer warps x one 128-row tile (csrc/gx_agg.cu)


def producer_window(klo, khi, slot, kmin, kmax, nslots, seg_slots, win=31, amask=3):
    """gx_k_runjoin_seg, producer warp: (lo, len) in slots for a chunk whose first / last key are klo / khi."""
    ka, kz = max(klo, kmin), min(khi, kmax)
    if ka > kz:
        return 0, 0
    wm = win | amask
    lo = slot(ka) & ~wm
    hi = min((slot(kz) | wm) + 1 + 32, nslots)
    if hi <= lo or hi - lo > 4 * seg_slots:
        return lo, 0
    return lo, min(hi - lo, seg_slots)


@pytest.mark.parametrize("name,okeys", [k for k in key_sets() if not k[0].startswith("two clusters")],
                         ids=[k[0].split(" (")[0] for k in key_sets() if not k[0].startswith("two clusters")])
def test_staged_window_holds_every_home_group_of_a_key_ordered_chunk(name, okeys):
    rng = np.random.default_rng(17)
    n = len(okeys)
    nslots = 1
    while nslots < n + n // 2 + 16:
        nslots *= 2
    kmin, kmax = int(okeys[0]), int(okeys[-1])
    slot = make_slot_fn(kmin, float(kmax - kmin + 1), nslots)
    lkeys = np.repeat(okeys, rng.integers(1, 8, n))               # 1..7 lines per order, key order
    # plus probe keys the build side does not hold: below, above and inside its span
    lkeys = np.sort(np.concatenate([lkeys, rng.integers(kmin - 5000, kmax + 5000, 20_000)]))
    avail = (232448 - 1024 - 98304 - 31 * 2560 - 112) // 8        # shared memory left of 227 KB for the ring, in 8-byte slots
    seg_slots = (avail // 3) & ~31
    homes = np.array([slot(int(k)) for k in np.clip(lkeys, kmin, kmax)])
    staged = total = 0
    prev_end = 0
    for c0 in range(0, len(lkeys), CHUNK):
        ks = lkeys[c0:c0 + CHUNK]
        lo, ln = producer_window(int(ks[0]), int(ks[-1]), slot, kmin, kmax, nslots, seg_slots)
        assert lo % 4 == 0 and ln % 4 == 0 and lo + ln <= nslots  # what cp.async.bulk needs: 32-byte pieces inside the table
        inspan = (ks >= kmin) & (ks <= kmax)                      # keys outside the span are rejected before any slot is read
        h = homes[c0:c0 + CHUNK][inspan]
        inside = (h >= lo) & (h + 4 <= lo + ln)
        total += len(h); staged += int(inside.sum())
        if ln and ln < seg_slots:                                 # not clipped by the buffer: everything must be inside
            assert inside.all(), (name, c0)
        if ln:
            assert lo <= prev_end + 64 or prev_end == 0           # consecutive windows overlap or touch: the table is read front to back
            prev_end = lo + ln
    assert staged >= 0.99 * total, (name, staged, total)


def test_unordered_chunks_stage_nothing_or_little():
    """A shuffled probe side: first and last key of a chunk are unrelated, the window is skipped when it would be
    more than four buffers wide — the kernel then probes global memory as gx_k_runjoin does."""
    n = 120_000
    i = np.arange(n, dtype=np.int64)
    okeys = ((i >> 3) << 5 | (i & 7)) + 1
    nslots = 262144
    slot = make_slot_fn(1, float(okeys[-1]), nslots)
    lkeys = np.random.default_rng(2).permutation(np.repeat(okeys, 4))
    skipped = 0
    nchunks = 0
    for c0 in range(0, len(lkeys), CHUNK):
        ks = lkeys[c0:c0 + CHUNK]
        _, ln = producer_window(int(ks[0]), int(ks[-1]), slot, 1, int(okeys[-1]), nslots, 2208)
        skipped += ln == 0
        nchunks += 1
    assert skipped >= 0.9 * nchunks


# ------------------------------------------------------------------ 5. the branch-free fold and the carry rounds
def fold_tile_branch_free(keys, vals, nrows, off=0):
    """One 128-row tile as the 32 lanes of gx_k_runjoin_tma<FOLD2> see it (csrc/gx_agg.cu): returns the run list
    (key, count, sum) written at positions off..off+NR-1, built from per-lane select chains, predicated stores and
    one continuation add per lane — the arithmetic of the kernel, lane by lane."""
    K = {}; C = {}; S = {}
    cont = []                                            # (run index, count, sum) added after the stores (second phase)
    inc = 0
    bases = []
    heads = []
    for lane in range(32):
        act = lane * 4 < nrows
        k = [int(keys[lane * 4 + i]) if lane * 4 + i < len(keys) else 0 for i in range(4)]
        prevk = int(keys[lane * 4 - 1]) if lane > 0 else None
        h = [act and (lane == 0 or k[0] != prevk), act and k[1] != k[0], act and k[2] != k[1], act and k[3] != k[2]]
        heads.append(h)
        bases.append(inc + off)
        inc += sum(h)
    NR = inc
    for lane in range(32):
        act = lane * 4 < nrows
        if not act:
            continue
        k = [int(keys[lane * 4 + i]) for i in range(4)]
        v = [float(vals[lane * 4 + i]) for i in range(4)]
        h0, h1, h2, h3 = heads[lane]
        base = bases[lane]
        r0 = base + int(h0) - 1; r1 = r0 + int(h1); r2 = r1 + int(h2); r3 = r2 + int(h3)
        c1 = 1 if h1 else 2; c2 = 1 if h2 else c1 + 1; c3 = 1 if h3 else c2 + 1
        s0 = v[0]; s1 = v[1] if h1 else s0 + v[1]; s2 = v[2] if h2 else s1 + v[2]; s3 = v[3] if h3 else s2 + v[3]
        q0 = h0; q1 = q0 or h1; q2 = q1 or h2; q3 = q2 or h3
        for hh, r, kk in ((h0, r0, k[0]), (h1, r1, k[1]), (h2, r2, k[2]), (h3, r3, k[3])):
            if hh:
                K[r] = kk
        if h1 and q0: C[r0], S[r0] = 1, s0
        if h2 and q1: C[r1], S[r1] = c1, s1
        if h3 and q2: C[r2], S[r2] = c2, s2
        if q3: C[r3], S[r3] = c3, s3
        if not h0:
            cc = 1 if h1 else c1 if h2 else c2 if h3 else c3
            sc = s0 if h1 else s1 if h2 else s2 if h3 else s3
            cont.append((base - 1, cc, sc))
    for r, cc, sc in cont:                               # after __syncwarp(): atomicAdd onto the run's own entry
        C[r] += cc; S[r] += sc
    assert sorted(K) == sorted(C) == list(range(off, off + NR))
    return [(K[r], C[r], S[r]) for r in range(off, off + NR)]


def fold_tile_sequential(keys, vals, nrows):
    out = []
    for i in range(nrows):
        if i == 0 or keys[i] != keys[i - 1]:
            out.append([int(keys[i]), 0, 0.0])
        out[-1][1] += 1; out[-1][2] += float(vals[i])
    return [tuple(x) for x in out]


@pytest.mark.parametrize("seed,maxrun", [(1, 7), (2, 1), (3, 40), (4, 300), (5, 2)])
def test_branch_free_fold_equals_the_sequential_fold(seed, maxrun):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        runs = rng.integers(1, maxrun + 1, 200)
        keys = np.repeat(np.cumsum(rng.integers(1, 5, 200)), runs)[:128 + 64]
        start = int(rng.integers(0, 64))                 # tiles begin mid-run
        nrows = int(rng.choice([128, 128, 128, 4 * rng.integers(1, 32)]))      # full tiles and partial last tiles
        tile = keys[start:start + 128]
        vals = rng.integers(1, 1000, 128).astype(np.float64)                    # integers: float sums are exact in any order
        got = fold_tile_branch_free(tile, vals, nrows)
        want = fold_tile_sequential(tile, vals, nrows)
        assert got == want


def test_carry_rounds_visit_every_run_once_in_full_rounds():
    """GX_RUNJOIN_TMA=3: T = waiting + new runs; the rounds take entries [T mod 32, T), the first T mod 32 entries wait;
    a tile that would not fit behind the waiting entries (capacity 128) has them probed first."""
    rng = np.random.default_rng(9)
    for regime in ("tpch", "one_row_keys", "mixed"):
        lst = [None] * 128
        nc = 0
        seen = []
        rounds = []
        produced = 0
        for t in range(400):
            NR = {"tpch": int(rng.integers(20, 46)), "one_row_keys": 128, "mixed": int(rng.choice([30, 33, 128, 100, 1]))}[regime]
            if nc + NR > 128:
                seen += lst[:nc]; rounds.append(nc); nc = 0
            for j in range(NR):
                lst[nc + j] = produced; produced += 1
            T = nc + NR
            first = T & 31
            for j0 in range(first, T, 32):
                seen += lst[j0:j0 + 32]; rounds.append(32)
            nc = first
        seen += lst[:nc]
        assert sorted(seen) == list(range(produced))     # every run exactly once
        partial = [r for r in rounds if r != 32]
        assert all(r < 32 for r in partial)
        if regime == "tpch":
            assert not partial and len(rounds) <= produced // 32


# ------------------------------------------------------------------ 6. the mbarrier ring protocol
class MBar:
    """An mbarrier as the kernels use it: `count` arrivals (+ outstanding transaction bytes) complete a phase;
    try_wait.parity(p) is true once the phase with parity p has completed (i.e. the current phase has the other parity)."""
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_complete()

    def try_wait(self, parity):
        return self.phase != parity


@pytest.mark.parametrize("nbuf,nconsumers,seed", [(2, 31, 1), (3, 31, 2), (4, 31, 3), (2, 1, 4), (3, 5, 5)])
def test_ring_protocol_is_deadlock_free_and_never_overwrites_a_buffer_in_use(nbuf, nconsumers, seed):
    """Random scheduling of the producer, the copy engine and the consumers of gx_k_runjoin_seg (csrc/gx_agg.cu):
    producer: for chunk j: (use > 0) wait empty[b] parity (use-1)&1; write lo/len; arrive.expect_tx(full[b]); bulk copy
    consumer: for chunk j: wait full[b] parity use&1; read lo/len + the buffer; arrive(empty[b])
    with b = j % nbuf, use = j // nbuf on both sides."""
    rng = np.random.default_rng(seed)
    niter = 40
    full = [MBar(1) for _ in range(nbuf)]
    empty = [MBar(nconsumers) for _ in range(nbuf)]
    content = [None] * nbuf                   # which chunk a buffer holds (written by the copy engine)
    meta = [None] * nbuf                      # lo/len words written by the producer before its arrive
    inflight = []                             # (buffer, chunk) bulk copies issued and not yet landed
    readers = [0] * nbuf                      # consumers between "wait full" and "arrive empty" on a buffer
    pj = 0
    cj = [0] * nconsumers
    cstate = ["wait"] * nconsumers            # wait -> reading -> (arrive) wait
    consumed = [[] for _ in range(nconsumers)]
    steps = 0
    while pj < niter or any(c < niter for c in cj) or inflight:
        steps += 1
        assert steps < 200_000, "no progress: deadlock"
        actors = []
        if pj < niter:
            actors.append(("p", 0))
        actors += [("e", i) for i in range(len(inflight))]
        actors += [("c", w) for w in range(nconsumers) if cj[w] < niter]
        kind, idx = actors[int(rng.integers(len(actors)))]
        if kind == "p":
            b, use = pj % nbuf, pj // nbuf
            if use > 0 and not empty[b].try_wait((use - 1) & 1):
                continue                      # still waiting for the consumers of chunk pj - nbuf
            assert readers[b] == 0, "refilling a buffer a consumer still reads"
            meta[b] = pj
            full[b].arrive(expect_tx=100)
            inflight.append((b, pj))
            pj += 1
        elif kind == "e":                     # the copy engine lands one outstanding copy
            b, chunk = inflight.pop(idx)
            assert readers[b] == 0
            content[b] = chunk
            full[b].complete_tx(100)
        else:
            w = idx
            b, use = cj[w] % nbuf, cj[w] // nbuf
            if cstate[w] == "wait":
                if full[b].try_wait(use & 1):
                    assert meta[b] == cj[w] and content[b] == cj[w], "consumer sees another chunk's window"
                    readers[b] += 1
                    cstate[w] = "reading"
            else:
                assert content[b] == cj[w]    # nothing changed the buffer while it was being read
                consumed[w].append(cj[w])
                readers[b] -= 1
                empty[b].arrive()
                cstate[w] = "wait"
                cj[w] += 1
    assert all(c == list(range(niter)) for c in consumed)
