"""TPC-H Q3 shape on N datanodes (one per GPU) with NCCL redistributes — BASELINE
config 4 — checked against the single-node CPU oracle.  Launch with torchrun:

    torchrun --nproc-per-node 2 scripts/q3_multi.py [--sf 1 --orders 20000]

Plan per datanode (the reference's plan shape, xc_groupby.out:389-394 style):
    customer[c_mktsegment = 'B']                       -> hash h1 on c_custkey
    orders[o_orderdate < D] --Distribute by o_custkey--> probe h1
        --Distribute by o_orderkey--> hash h2 (payload o_orderdate, o_shippriority)
    lineitem[l_shipdate > D] probe h2 -> GROUP BY l_orderkey, o_orderdate, o_shippriority
                                          sum(l_extendedprice * (1 - l_discount))
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g  # noqa: E402

DATE = -1752            # 1995-03-15 as days since 2000-01-01


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, default=1)
    ap.add_argument("--orders", type=int, default=20000, help="orders generated (0 = the whole scale factor)")
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--iters", type=int, default=1, help="query repetitions (the first one is cold)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = g.Context(local)
    box = [g.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(rank, world, box[0])
    ctx.set_shardmap(world)

    nord = a.orders or 1_500_000 * a.sf
    ncust = 150_000 * a.sf
    cust = ctx.table(g.SCHEMAS[g.T_CUSTOMER], ncust // world * 2 + 1024).generate(g.T_CUSTOMER, a.sf, 0, ncust, rank, world)
    orders = ctx.table(g.SCHEMAS[g.T_ORDERS], nord // world * 2 + 1024).generate(g.T_ORDERS, a.sf, 0, nord, rank, world)
    line = ctx.table(g.SCHEMAS[g.T_LINEITEM], nord * 7 // world * 2 + 1024).generate(g.T_LINEITEM, a.sf, 0, nord, rank, world)

    C, K, S, M = g.GX_OP_COL, g.GX_OP_CONST, g.GX_OP_SUB, g.GX_OP_MUL
    rev = [(C, g.L_EXTENDEDPRICE, 0), (K, 0, 1.0), (C, g.L_DISCOUNT, 0), (S, 0, 0), (M, 0, 0)]
    plan = g.make_plan(preds=[(g.L_SHIPDATE, g.GX_GT, DATE)], outer_key_col=g.L_ORDERKEY,
                       group_cols=[(0, g.L_ORDERKEY), (1, 0), (1, 1)], aggs=[(g.GX_AGG_SUM_F8, rev)],
                       est_groups=max(nord // world // 4, 1024))
    for it in range(a.iters):
        ctx.sync(); dist.barrier()
        t0 = time.perf_counter()
        t1 = ctx.scan_filter(cust, [(g.C_MKTSEGMENT, g.GX_EQ, ord("B"))], [g.C_CUSTKEY])
        t2 = ctx.scan_filter(orders, [(g.O_ORDERDATE, g.GX_LT, DATE)], [g.O_ORDERKEY, g.O_CUSTKEY, g.O_ORDERDATE, g.O_SHIPPRIORITY])
        t2r = ctx.redistribute(t2, 1)                                   # all-to-all on o_custkey
        h1 = ctx.hash_build(t1, 0, [], unique=True)
        j1 = ctx.hash_probe(t2r, 1, h1, [0, 2, 3])                      # o_orderkey, o_orderdate, o_shippriority (+ build row)
        j1r = ctx.redistribute(j1, 0)                                   # all-to-all back on o_orderkey
        h2 = ctx.hash_build(j1r, 0, [1, 2], unique=True)
        res = ctx.hash_agg(line, plan, h2)
        keys, aggs, nulls = res.fetch()
        ctx.sync(); dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0 and a.iters > 1:
            print(f"q3 iteration {it}: {dt * 1e3:.2f} ms")
        if it + 1 < a.iters:
            for x in (res, h2, h1): x.free()
            for x in (j1r, j1, t2r, t2, t1): x.free()
    rows_in = cust.nrows + orders.nrows + line.nrows

    gathered = [None] * world
    dist.all_gather_object(gathered, (keys, aggs, rows_in, t2.nrows, j1.nrows))
    if rank == 0:
        allk = np.concatenate([x[0] for x in gathered]); alla = np.concatenate([x[1] for x in gathered])
        total_rows = sum(x[2] for x in gathered)
        print(f"q3: {world} datanodes, {total_rows} input rows, {len(allk)} groups, {dt * 1e3:.2f} ms, "
              f"{total_rows / dt / 1e9:.2f} G rows/s; redistributed {sum(x[3] for x in gathered)} + {sum(x[4] for x in gathered)} rows")
        if a.check:
            import oracle as O
            from oracle import make_join, make_plan
            c, o, l = O.gen_customer(a.sf, 0, ncust), O.gen_orders(a.sf, 0, nord), O.gen_lineitem(a.sf, 0, nord)
            good = np.isin(o[1], c[0][c[1] == ord("B")]) & (o[2] < DATE)     # the customer join, restated with numpy
            oj = [x[good] for x in o]
            oplan = make_plan(preds=[(g.L_SHIPDATE, g.GX_GT, DATE)], outer_key_col=g.L_ORDERKEY,
                              group_cols=[(0, g.L_ORDERKEY), (1, 0), (1, 1)], aggs=[(g.GX_AGG_SUM_F8, rev)])
            want = O.exec_agg(O.Rel(g.SCHEMAS[g.T_LINEITEM], l), oplan, O.Rel(g.SCHEMAS[g.T_ORDERS], oj),
                              make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE, g.O_SHIPPRIORITY], inner_unique=1)).sorted()
            order = np.lexsort([allk[:, 2], allk[:, 1], allk[:, 0]])
            np.testing.assert_array_equal(allk[order], want.keys)
            np.testing.assert_allclose(alla[order, 0], want.aggs[:, 0], rtol=1e-9, atol=0)
            print(f"OK q3 parity vs oracle: {want.ngroups} groups")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
