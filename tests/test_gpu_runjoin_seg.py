"""gx_k_runjoin_seg — the config-3 probe that streams the join table through a shared-memory ring
(cp.async.bulk + mbarrier) instead of gathering it — against the oracle's HashJoin + HashAggregate
(nodeHashjoin.c:186-742, nodeAgg.c:2609-2648), on the layouts that decide which of its code paths run:

* both sides in key order              every probe answered from the staged window
* outer side shuffled                  chunk key ranges far wider than the ring: nothing staged, global probes
* outer side descending                first key > last key of every chunk: empty windows
* build side with holes / narrow span  misses, keys below and above the build side's key span (clamped windows)
* tiny outer sides                     fewer rows than one chunk, a tail that is not a multiple of four

and with every ring depth.  The switch GX_RUNJOIN_SEG forces the variant on or off; profile name
`probe_agg_seg` proves which kernel ran."""
import numpy as np
import pytest

import opentenbase_b200 as g
import oracle as O
from helpers import assert_agg_equal, to_gpu_plan, lineitem_rel, orders_rel

pytestmark = pytest.mark.gpu

NORD = 200_000        # >= 64 sub-tables: interpolation slots + compact table, what the variant needs


def _plan(count=True, total=True):
    aggs = []
    if count:
        aggs.append((g.GX_AGG_COUNT_STAR, []))
    if total:
        aggs.append((g.GX_AGG_SUM_F8, [(g.GX_OP_COL, g.L_EXTENDEDPRICE, 0)]))
    return O.make_plan(outer_key_col=g.L_ORDERKEY, group_cols=[(1, 0)], aggs=aggs, est_groups=2600)


def _run(gx, o, l, plan, expect_seg=True):
    join = O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1)
    want = O.exec_agg(lineitem_rel(l), plan, orders_rel(o), join)
    ot = gx.table_from(g.SCHEMAS[g.T_ORDERS], o)
    lt = gx.table_from(g.SCHEMAS[g.T_LINEITEM], l)
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    if expect_seg:
        assert ht.info()["slot_mode"] == 2, ht.info()        # key-ordered build: compact slots, order-preserving
    gx.profile(True)
    try:
        got = gx.hash_agg(lt, to_gpu_plan(plan), ht).fetch()
        _, nseg = gx.profile_get("probe_agg_seg")
    finally:
        gx.profile(False)
    if expect_seg is not None:
        assert (nseg > 0) == expect_seg, f"probe_agg_seg launches: {nseg}"
    assert_agg_equal(plan, got, want)
    for t in (ht, lt, ot):
        t.free()
    return got


@pytest.fixture(scope="module")
def base():
    o = [c.copy() for c in O.gen_orders(1, 0, NORD)]
    l = [c.copy() for c in O.gen_lineitem(1, 0, NORD)]
    return o, l


@pytest.mark.parametrize("bufs", ["2", "3", "4"])
def test_key_ordered_both_sides(gx, base, monkeypatch, bufs):
    monkeypatch.setenv("GX_RUNJOIN_SEG", "1")
    monkeypatch.setenv("GX_RUNJOIN_SEG_BUFS", bufs)
    o, l = base
    got = _run(gx, o, l, _plan())
    assert got[1][:, 0].view(np.int64).sum() == len(l[0])     # every line finds its order


@pytest.mark.parametrize("aggs", ["count", "sum"])
def test_single_aggregate_instantiations(gx, base, monkeypatch, aggs):
    monkeypatch.setenv("GX_RUNJOIN_SEG", "1")
    o, l = base
    _run(gx, o, l, _plan(count=aggs == "count", total=aggs == "sum"))


@pytest.mark.parametrize("layout", ["shuffled", "descending", "block_shuffled"])
def test_outer_side_not_in_key_order(gx, base, monkeypatch, layout):
    """The windows are a guess from the first and last key of a chunk; the answer may not depend on it."""
    monkeypatch.setenv("GX_RUNJOIN_SEG", "1")
    o, l = base
    n = len(l[0])
    if layout == "shuffled":
        perm = np.random.default_rng(3).permutation(n)
    elif layout == "descending":
        perm = np.arange(n)[::-1].copy()
    else:                                   # key order inside blocks of 1000 rows, the blocks in random order:
        nb = (n + 999) // 1000              # chunks straddle unrelated key ranges, some windows hold part of the keys
        order = np.random.default_rng(4).permutation(nb)
        perm = np.concatenate([np.arange(b * 1000, min((b + 1) * 1000, n)) for b in order])
    _run(gx, o, [c[perm] for c in l], _plan())


def test_misses_and_keys_outside_the_build_span(gx, base, monkeypatch):
    """Build side = the middle half of the orders with every third order removed: lines of the first and
    last quarter carry keys below / above the build side's span (windows clamped to the span or empty),
    lines of removed orders miss inside it."""
    monkeypatch.setenv("GX_RUNJOIN_SEG", "1")
    o, l = base
    keep = np.zeros(NORD, bool)
    keep[NORD // 4: 3 * NORD // 4] = True
    keep[::3] = False
    got = _run(gx, [c[keep] for c in o], l, _plan(), expect_seg=None)     # a 67 k-row build side is at the edge of the interpolation rule
    assert 0 < got[1][:, 0].view(np.int64).sum() < len(l[0])


@pytest.mark.parametrize("nrows", [1, 3, 4, 127, 130, 3967, 3968, 3971, 4 * 3968 + 5])
def test_short_outer_sides(gx, base, monkeypatch, nrows):
    """Fewer rows than one chunk (31 tiles of 128 rows), one chunk exactly, tails of 1-3 rows."""
    monkeypatch.setenv("GX_RUNJOIN_SEG", "1")
    o, l = base
    start = 123_457                          # somewhere inside the table, mid-run
    _run(gx, o, [c[start:start + nrows] for c in l], _plan())


def test_switch_off_keeps_the_gathering_kernel(gx, base, monkeypatch):
    monkeypatch.setenv("GX_RUNJOIN_SEG", "0")
    o, l = base
    _run(gx, o, l, _plan(), expect_seg=False)


def test_same_answer_as_the_gathering_kernel_at_sf10_slice(gx, monkeypatch):
    """10 M orders generated on the device (no oracle at this size): both variants must agree bit for bit
    on the counts and to 1e-9 on the sums, and count(*) must equal the lineitem rows."""
    nord = 10_000_000
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, 10, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], nord * 7).generate(g.T_LINEITEM, 10, 0, nord)
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    plan = to_gpu_plan(_plan())
    res = {}
    for sw in ("0", "1"):
        monkeypatch.setenv("GX_RUNJOIN_SEG", sw)
        k, a, _ = gx.hash_agg(lt, plan, ht).fetch()
        order = np.argsort(k[:, 0])
        res[sw] = (k[order], a[order])
    np.testing.assert_array_equal(res["0"][0], res["1"][0])
    np.testing.assert_array_equal(res["0"][1][:, 0].view(np.int64), res["1"][1][:, 0].view(np.int64))
    np.testing.assert_allclose(res["0"][1][:, 1], res["1"][1][:, 1], rtol=1e-9, atol=0)
    assert res["1"][1][:, 0].view(np.int64).sum() == lt.nrows
    for t in (ht, lt, ot):
        t.free()
