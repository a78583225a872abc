"""Per-kernel measurements for the rows of SURVEY.md §8(a) that bench.py's headline does not
cover: BASELINE configs 1 and 2, the Q1 shape of config 5, K1 (quals + projection), K5
(SHARD routing + partition) and the Q3 shape on one datanode (K0, the heap-page deform, is timed
by a GPU test: it needs page images that only the test-side oracle builds).  One
GPU, inputs resident in HBM, CUDA events on the library's stream, algorithmic bytes as
SURVEY.md §8(d) defines them.  Prints one JSON object; results go to profiles/.

    python scripts/bench_configs.py [--sf 100] [--iters 5]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g  # noqa: E402


LABELS = ("build_sample", "build_bounds", "build_scatter", "build", "build_clear", "build_expand", "probe_agg", "agg", "agg_compact",
          "probe_records", "scan_records", "radix_partition", "radix_agg", "radix_overflow", "filter", "partition", "route", "probe", "probe_count")
LAST_PHASES = {}


def timed(ctx, fn, iters, warm=2):
    """ms per call (CUDA events around `iters` calls) + the per-label kernel times of those calls."""
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.profile(True)
    ctx.timer_start()
    for _ in range(iters):
        fn()
    ms = ctx.timer_stop() / iters
    LAST_PHASES.clear()
    for name in LABELS:
        t, n = ctx.profile_get(name)
        if n:
            LAST_PHASES[name] = round(t / iters, 4)
    ctx.profile(False)
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, default=100)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    peak = 6570.3
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    ctx = g.Context(0)
    sf = a.sf
    no = 1_500_000 * sf
    LT = [g.L_ORDERKEY, g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_RETURNFLAG, g.L_LINESTATUS, g.L_SHIPDATE]
    ltypes = [g.SCHEMAS[g.T_LINEITEM][c] for c in LT]
    lt = ctx.table(ltypes, no * 4 + no // 8).generate(g.T_LINEITEM, sf, 0, no, colmap=LT)
    OT = [g.O_ORDERKEY, g.O_CUSTKEY, g.O_ORDERDATE, g.O_SHIPPRIORITY]
    ot = ctx.table([g.SCHEMAS[g.T_ORDERS][c] for c in OT], no).generate(g.T_ORDERS, sf, 0, no, colmap=OT)
    nl = lt.nrows
    kO, kQ, kP, kD, kT, kF, kS, kSH = range(8)
    C, K, A, S, M = g.GX_OP_COL, g.GX_OP_CONST, g.GX_OP_ADD, g.GX_OP_SUB, g.GX_OP_MUL
    out = {"sf": sf, "lineitem_rows": nl, "orders_rows": no, "hbm_peak_gbs": peak, "iters": a.iters, "kernels": {}}

    def record(name, ms, rows, bytes_per_row, note, extra_bytes=0):
        gb = (rows * bytes_per_row + extra_bytes) / 1e9
        out["kernels"][name] = {"ms": round(ms, 4), "rows": rows, "rows_per_s": rows / ms * 1e3, "algorithmic_gb": round(gb, 3),
                                "achieved_gbs": round(gb / ms * 1e3, 1), "frac_of_hbm_peak": round(gb / ms * 1e3 / peak, 3), "note": note,
                                "kernel_ms": dict(LAST_PHASES)}

    def agg(plan, ht=None, table=lt):
        r = ctx.hash_agg(table, plan, ht)
        k = r.fetch()
        r.free()
        return k

    # ---- config 1: SELECT l_returnflag, count(*) GROUP BY 1
    p1 = g.make_plan(group_cols=[(0, kF)], aggs=[(g.GX_AGG_COUNT_STAR, [])], est_groups=3)
    record("config1 count(*) GROUP BY l_returnflag", timed(ctx, lambda: agg(p1), a.iters), nl, 1, "1 B/row (flag)")
    # ---- config 2: sum(l_extendedprice) GROUP BY l_shipdate
    p2 = g.make_plan(group_cols=[(0, kSH)], aggs=[(g.GX_AGG_SUM_F8, [(C, kP, 0)])], est_groups=2600)
    record("config2 sum(l_extendedprice) GROUP BY l_shipdate", timed(ctx, lambda: agg(p2), a.iters), nl, 12, "8 price + 4 shipdate")
    # ---- Q1 shape
    disc_price = [(C, kP, 0), (K, 0, 1.0), (C, kD, 0), (S, 0, 0), (M, 0, 0)]
    charge = disc_price + [(K, 0, 1.0), (C, kT, 0), (A, 0, 0), (M, 0, 0)]
    pq1 = g.make_plan(preds=[(kSH, g.GX_LE, -517 - 90)], group_cols=[(0, kF), (0, kS)],
                      aggs=[(g.GX_AGG_SUM_F8, [(C, kQ, 0)]), (g.GX_AGG_SUM_F8, [(C, kP, 0)]), (g.GX_AGG_SUM_F8, disc_price),
                            (g.GX_AGG_SUM_F8, charge), (g.GX_AGG_AVG_F8, [(C, kQ, 0)]), (g.GX_AGG_AVG_F8, [(C, kP, 0)]),
                            (g.GX_AGG_AVG_F8, [(C, kD, 0)]), (g.GX_AGG_COUNT_STAR, [])], est_groups=6)
    record("Q1 shape (4 groups, 8 aggregates, shipdate qual)", timed(ctx, lambda: agg(pq1), a.iters), nl, 38,
           "quantity 8 + price 8 + discount 8 + tax 8 + shipdate 4 + flag 1 + status 1")
    # ---- K1: quals + projection
    def k1():
        t = ctx.scan_filter(lt, [(kSH, g.GX_GT, -1752)], [kO, kP, kD])
        n = t.nrows
        t.free()
        return n
    nsel = k1()
    record("K1 scan_filter l_shipdate > D -> (orderkey, price, discount)", timed(ctx, k1, a.iters), nl, 4,
           f"4 B/row qual column + 24 B per selected row read + 24 B written ({nsel} selected)", extra_bytes=nsel * 48)
    # ---- K5: SHARD routing + partition of orders on o_custkey over 4 datanodes
    ctx.set_shardmap(4)
    def k5():
        t, counts = ctx.partition_by_node(ot, 1, 4)
        t.free()
    record("K5 partition_by_node(orders, o_custkey, 4 nodes)", timed(ctx, k5, a.iters), no, 20 + 4 + 20,
           "20 B/row read for the copy + 4 B key read for the histogram pass + 20 B/row written")
    ctx.set_shardmap(1)
    # ---- config 3 pieces, for reference next to bench.py
    def build():
        h = ctx.hash_build(ot, 0, [2], unique=True)
        h.free()
    record("K2 hash build over orders (o_orderkey -> o_orderdate)", timed(ctx, build, a.iters), no, 8 + 4 + 8,
           "8 key + 4 payload read, 8 B compact slot written (16 B wide)")
    ht = ctx.hash_build(ot, 0, [2], unique=True)
    p3 = g.make_plan(outer_key_col=kO, group_cols=[(1, 0)], aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(C, kP, 0)])], est_groups=2500)
    record("K3+K4 probe + aggregate (config 3)", timed(ctx, lambda: agg(p3, ht), a.iters), nl, 24, "8 key + 8 payload on hit + 8 price")
    # ---- Q3 shape on one datanode: orders filtered, join, GROUP BY (l_orderkey, o_orderdate, o_shippriority)
    rev = [(C, kP, 0), (K, 0, 1.0), (C, kD, 0), (S, 0, 0), (M, 0, 0)]
    def q3():
        h = ctx.hash_build(ot, 0, [2, 3], unique=True, preds=[(2, g.GX_LT, -1752)])
        pl = g.make_plan(preds=[(kSH, g.GX_GT, -1752)], outer_key_col=kO, group_cols=[(0, kO), (1, 0), (1, 1)],
                         aggs=[(g.GX_AGG_SUM_F8, rev)], est_groups=max(no // 8, 1024))
        r = ctx.hash_agg(lt, pl, h)
        n = r.ngroups
        r.free(); h.free()
        return n
    ng = q3()
    record("Q3 shape without the customer join (1 datanode)", timed(ctx, q3, max(1, a.iters // 2), warm=1), nl + no, 0,
           f"lineitem 28 B/row + orders 16 B/row + 16 B slot per kept order; {ng} groups", extra_bytes=nl * 28 + no * 16 + no // 2 * 16)
    ht.free()
    # K0 (heap pages -> columns) needs page images, which only the test-side oracle can build:
    # its throughput is measured by tests/test_gpu_parity.py::test_heap_page_deform_throughput
    print(json.dumps(out, indent=1))
    for t in (lt, ot):
        t.free()
    ctx.close()


if __name__ == "__main__":
    main()
