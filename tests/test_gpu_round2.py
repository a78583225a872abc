"""Round-2 parity cases (all through the C ABI, all against the oracle):
float8 overflow -> ERROR, tuples shorter than the descriptor, the Q3 fragment chain as the
provider issues it, full-SF1 / SF10-slice comparisons of every BASELINE query shape."""
import numpy as np
import pytest

import opentenbase_b200 as g
from opentenbase_b200 import plans as P
import oracle as O
from helpers import assert_agg_equal, to_gpu_plan, lineitem_rel, orders_rel

pytestmark = pytest.mark.gpu

CCOLS = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
OCOLS = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
LCOLS = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}


def test_float8_overflow_is_an_error(gx):
    """float8pl's CHECKFLOATVAL (utils/adt/float.c:970-981): 1e308 + 1e308 from finite inputs is
    'value out of range: overflow' in the reference, GX_ERR_OVERFLOW here; an infinite INPUT is fine."""
    t = gx.table_from([g.GX_INT4, g.GX_FLOAT8], [np.array([1, 1, 2], np.int32), np.array([1e308, 1e308, 5.0])])
    plan = g.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=4)
    with pytest.raises(g.GxError) as ei:
        gx.hash_agg(t, plan).fetch()
    assert ei.value.status == g.GX_ERR_OVERFLOW
    plan = g.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=4)
    with pytest.raises(g.GxError) as ei:
        gx.hash_agg(t, plan).fetch()
    assert ei.value.status == g.GX_ERR_OVERFLOW


def test_tuples_shorter_than_the_descriptor(gx):
    """ALTER TABLE ADD COLUMN without a rewrite: old tuples carry fewer attributes
    (t_infomask2 & HEAP_NATTS_MASK) and read the new ones as NULL — slot_deform_tuple's
    Min(HeapTupleHeaderGetNatts(tup), natts), heaptuple.c:1424,1497-1502,1555."""
    rng = np.random.default_rng(11)
    n = 3000
    rel = O.Rel([O.GX_INT8, O.GX_INT4, O.GX_FLOAT8],
                [rng.integers(0, 2**40, n), rng.integers(0, 100, n).astype(np.int32), rng.random(n)],
                [None, (rng.random(n) < 0.2).astype(np.uint8), None])
    rel.add_column(O.GX_FLOAT8)
    rel.add_column(O.GX_INT4)
    m = 500                                                    # rows written after the ALTER carry all five attributes
    rel.insert([rng.integers(0, 2**40, m), rng.integers(0, 100, m).astype(np.int32), rng.random(m), rng.random(m),
                rng.integers(0, 9, m).astype(np.int32)])
    attnums = [0, 2, 3, 4]
    want_cols, want_nulls = rel.scan(attnums)
    assert want_nulls[2][:n].all() and not want_nulls[2][n:].any()
    t = gx.table([g.GX_INT8, g.GX_FLOAT8, g.GX_FLOAT8, g.GX_INT4], n + m)
    t.append_heap_pages(rel.pages(), [8, 4, 8, 8, 4], [8, 4, 8, 8, 4], attnums)
    assert t.nrows == n + m
    for c in range(len(attnums)):
        got, gn = t.read(c, with_nulls=True)
        np.testing.assert_array_equal(gn, want_nulls[c])
        keep = want_nulls[c] == 0
        np.testing.assert_array_equal(got.view(np.int64 if got.dtype == np.float64 else got.dtype)[keep],
                                      want_cols[c].view(np.int64 if got.dtype == np.float64 else got.dtype)[keep])


def test_drop_column(gx):
    t = gx.table_from([g.GX_INT8, g.GX_INT4, g.GX_FLOAT8], [np.arange(5), np.arange(5, dtype=np.int32) * 2, np.arange(5) * 0.5])
    t.drop_column(1)
    assert t.types == [g.GX_INT8, g.GX_FLOAT8]
    np.testing.assert_array_equal(t.read(1), np.arange(5) * 0.5)


@pytest.mark.parametrize("nord", [30000, 400000])
def test_q3_fragment_chain_one_datanode(gx, nord):
    """customer JOIN orders JOIN lineitem through plans.q3_datanode (scan_filter -> redistribute ->
    build -> probe -> redistribute -> build -> fused probe+aggregate) on a single datanode, where the
    redistributes degenerate to local partition copies; the N > 1 path is checked by bench.py's
    q3_parity_vs_oracle under torchrun and by tests/test_multi_gpu.py."""
    sf, ncust = 1, 150_000
    gx.set_shardmap(1)
    ct = gx.table(g.SCHEMAS[g.T_CUSTOMER], ncust).generate(g.T_CUSTOMER, sf, 0, ncust)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], nord * 7).generate(g.T_LINEITEM, sf, 0, nord)
    stats = {}
    res = P.q3_datanode(gx, ct, ot, lt, CCOLS, OCOLS, LCOLS, stats)
    want = O.q3_reference(sf, nord, ncust, P.DATE_Q3, P.SEGMENT_Q3)
    assert want.ngroups > 50 and stats["joined"] < stats["orders_kept"]
    assert_agg_equal(res.plan, res.fetch(), want)
    for t in (ct, ot, lt):
        t.free()


def test_full_sf1_every_query_shape(gx):
    """Every BASELINE query shape at FULL SF1 (1.5 M orders, 6.0 M lineitem rows) against the oracle:
    config 1, config 2, config 3, Q1."""
    sf, nord = 1, 1_500_000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], len(l[0])).generate(g.T_LINEITEM, sf, 0, nord)
    assert lt.nrows == len(l[0])
    lrel, orel = lineitem_rel(l), orders_rel(o)
    for plan in (P.config1_plan(g.L_RETURNFLAG), P.config2_plan(g.L_SHIPDATE, g.L_EXTENDEDPRICE),
                 P.q1_plan(g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_SHIPDATE, g.L_RETURNFLAG, g.L_LINESTATUS)):
        want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)))
        assert_agg_equal(plan, gx.hash_agg(lt, plan).fetch(), want)
    plan = P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE)
    want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)), orel, O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1))
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    assert_agg_equal(plan, gx.hash_agg(lt, plan, ht).fetch(), want)
    for t in (ot, lt):
        t.free()


def test_sf10_slice_config2_config3_q3(gx):
    """A 3 M-order slice of the SF10 tables (15 M lineitem rows) for configs 2 and 3 and the Q3 chain."""
    sf, nord, ncust = 10, 3_000_000, 1_500_000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], len(l[0])).generate(g.T_LINEITEM, sf, 0, nord)
    ct = gx.table(g.SCHEMAS[g.T_CUSTOMER], ncust).generate(g.T_CUSTOMER, sf, 0, ncust)
    lrel, orel = lineitem_rel(l), orders_rel(o)
    plan = P.config2_plan(g.L_SHIPDATE, g.L_EXTENDEDPRICE)
    assert_agg_equal(plan, gx.hash_agg(lt, plan).fetch(), O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan))))
    plan = P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE)
    want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)), orel, O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1))
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    assert_agg_equal(plan, gx.hash_agg(lt, plan, ht).fetch(), want)
    gx.set_shardmap(1)
    res = P.q3_datanode(gx, ct, ot, lt, CCOLS, OCOLS, LCOLS)
    assert_agg_equal(res.plan, res.fetch(), O.q3_reference(sf, nord, ncust, P.DATE_Q3, P.SEGMENT_Q3))
    for t in (ot, lt, ct):
        t.free()
