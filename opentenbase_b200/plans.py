"""Plan shapes of the BASELINE configs, written as the sequence of C-ABI calls the
CustomScan provider issues for them (host-side mirror of the reference's plan
trees).  Tests, bench.py and the scripts share these so that the thing that is
parity-checked is the thing that is timed.

    config 1   Agg(HASHED) <- SeqScan lineitem              count(*) GROUP BY l_returnflag
    config 2   Agg(HASHED) <- SeqScan lineitem              sum(l_extendedprice) GROUP BY l_shipdate
    config 3   Agg(HASHED) <- HashJoin <- {lineitem, Hash <- orders}
    Q1         Partial Agg per datanode -> Finalize          (xc_groupby.out:193-205)
    Q3         customer JOIN orders JOIN lineitem with two redistributes
               (RemoteSubplan "Distribute results by S", xc_groupby.out:389-394):
                 customer[c_mktsegment = seg]                       -> Hash h1 on c_custkey
                 orders[o_orderdate < D] --Distribute by o_custkey--> probe h1
                     --Distribute by o_orderkey--> Hash h2 (payload o_orderdate, o_shippriority)
                 lineitem[l_shipdate > D] probe h2 -> GROUP BY l_orderkey, o_orderdate, o_shippriority
                                                       sum(l_extendedprice * (1 - l_discount))
"""
from __future__ import annotations

from . import (GX_AGG_AVG_F8, GX_AGG_COUNT_STAR, GX_AGG_SUM_F8, GX_EQ, GX_GT, GX_LE, GX_LT, GX_OP_ADD, GX_OP_COL,
               GX_OP_CONST, GX_OP_MUL, GX_OP_SUB, make_plan)

DATE_Q3 = -1752          # 1995-03-15 as days since 2000-01-01
DATE_Q1 = -517 - 90      # 1998-12-01 - 90 days
SEGMENT_Q3 = ord("B")    # 'BUILDING'


def config1_plan(flag_col):
    return make_plan(group_cols=[(0, flag_col)], aggs=[(GX_AGG_COUNT_STAR, [])], est_groups=3)


def config2_plan(shipdate_col, price_col):
    return make_plan(group_cols=[(0, shipdate_col)], aggs=[(GX_AGG_SUM_F8, [(GX_OP_COL, price_col, 0)])], est_groups=2600)


def config3_plan(okey_col, price_col, maker=make_plan):
    """GROUP BY o_orderdate (payload 0 of the join table): count(*), sum(l_extendedprice)"""
    return maker(outer_key_col=okey_col, group_cols=[(1, 0)],
                 aggs=[(GX_AGG_COUNT_STAR, []), (GX_AGG_SUM_F8, [(GX_OP_COL, price_col, 0)])], est_groups=2500)


def revenue_expr(price_col, disc_col):
    """l_extendedprice * (1 - l_discount) as a postfix program (float8mul(float8mi(..)))"""
    C, K, S, M = GX_OP_COL, GX_OP_CONST, GX_OP_SUB, GX_OP_MUL
    return [(C, price_col, 0), (K, 0, 1.0), (C, disc_col, 0), (S, 0, 0), (M, 0, 0)]


def q1_plan(qty, price, disc, tax, shipdate, flag, status, maker=make_plan):
    """TPC-H Q1: 8 aggregates over 4 groups, l_shipdate <= date '1998-12-01' - 90 days"""
    C, K, A, M = GX_OP_COL, GX_OP_CONST, GX_OP_ADD, GX_OP_MUL
    disc_price = revenue_expr(price, disc)
    charge = disc_price + [(K, 0, 1.0), (C, tax, 0), (A, 0, 0), (M, 0, 0)]
    return maker(preds=[(shipdate, GX_LE, DATE_Q1)], group_cols=[(0, flag), (0, status)],
                 aggs=[(GX_AGG_SUM_F8, [(C, qty, 0)]), (GX_AGG_SUM_F8, [(C, price, 0)]), (GX_AGG_SUM_F8, disc_price),
                       (GX_AGG_SUM_F8, charge), (GX_AGG_AVG_F8, [(C, qty, 0)]), (GX_AGG_AVG_F8, [(C, price, 0)]),
                       (GX_AGG_AVG_F8, [(C, disc, 0)]), (GX_AGG_COUNT_STAR, [])], est_groups=6)


def q3_agg_plan(okey, price, disc, shipdate, est_groups, maker=make_plan):
    return maker(preds=[(shipdate, GX_GT, DATE_Q3)], outer_key_col=okey,
                 group_cols=[(0, okey), (1, 0), (1, 1)], aggs=[(GX_AGG_SUM_F8, revenue_expr(price, disc))],
                 est_groups=est_groups)


def q3_datanode(ctx, cust, orders, line, ccols, ocols, lcols, stats=None):
    """One datanode's fragment chain of the Q3 shape.  cust/orders/line are this datanode's
    shards (SHARD placement on c_custkey / o_orderkey / l_orderkey); *cols map column names
    to column numbers of those tables.  Returns the Result (partial == final: the group key
    contains the distribution key, so every group lives on one datanode —
    grouping_distribution_match, planner.c:10026).  `stats` (dict) receives row counts."""
    import time
    clock = [time.perf_counter()]
    host_ms = {}

    def lap(name):
        now = time.perf_counter()
        host_ms[name] = host_ms.get(name, 0.0) + (now - clock[0]) * 1e3
        clock[0] = now
    t1 = ctx.scan_filter(cust, [(ccols["mktsegment"], GX_EQ, SEGMENT_Q3)], [ccols["custkey"]])
    t2 = ctx.scan_filter(orders, [(ocols["orderdate"], GX_LT, DATE_Q3)],
                         [ocols["orderkey"], ocols["custkey"], ocols["orderdate"], ocols["shippriority"]])
    lap("scan_filter")
    t2r = ctx.redistribute(t2, 1)                                   # Distribute by o_custkey
    lap("redistribute_custkey")
    h1 = ctx.hash_build(t1, 0, [], unique=True)
    lap("build_customer")
    j1 = ctx.hash_probe(t2r, 1, h1, [0, 2, 3])                      # o_orderkey, o_orderdate, o_shippriority, build row
    j1.drop_column(3)                                               # the customer row number is not part of the target list
    lap("probe_customer")
    j1r = ctx.redistribute(j1, 0)                                   # Distribute by o_orderkey
    lap("redistribute_orderkey")
    h2 = ctx.hash_build(j1r, 0, [1, 2], unique=True)
    lap("build_orders")
    plan = q3_agg_plan(lcols["orderkey"], lcols["extendedprice"], lcols["discount"], lcols["shipdate"],
                       est_groups=max(j1r.nrows // 2, 1024))
    res = ctx.hash_agg(line, plan, h2)
    lap("probe_agg_lineitem")
    if stats is not None:
        stats["host_ms_per_call"] = host_ms
        stats.update(cust_kept=t1.nrows, orders_kept=t2.nrows, redistributed_custkey=t2.nrows, joined=j1.nrows,
                     redistributed_orderkey=j1.nrows, build_rows=j1r.nrows,
                     bytes_sent=t2.nrows * 20 + j1.nrows * 16)
    for x in (h2, h1):
        x.free()
    for x in (j1r, j1, t2r, t2, t1):
        x.free()
    return res
