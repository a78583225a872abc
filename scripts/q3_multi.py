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
    ctx.pool_reserve(0)

    from opentenbase_b200 import plans as P
    nord = a.orders or 1_500_000 * a.sf
    ncust = 150_000 * a.sf
    cust = ctx.table(g.SCHEMAS[g.T_CUSTOMER], ncust // world * 2 + 1024).generate(g.T_CUSTOMER, a.sf, 0, ncust, rank, world)
    orders = ctx.table(g.SCHEMAS[g.T_ORDERS], nord // world * 2 + 1024).generate(g.T_ORDERS, a.sf, 0, nord, rank, world)
    line = ctx.table(g.SCHEMAS[g.T_LINEITEM], nord * 7 // world * 2 + 1024).generate(g.T_LINEITEM, a.sf, 0, nord, rank, world)
    ccols = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
    ocols = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
    lcols = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}
    stats = {}
    ctx.profile(True)
    for it in range(a.iters):
        ctx.sync(); dist.barrier()
        t0 = time.perf_counter()
        res = P.q3_datanode(ctx, cust, orders, line, ccols, ocols, lcols, stats)
        keys, aggs, nulls = res.fetch()
        res.free()
        ctx.sync(); dist.barrier()
        dt = time.perf_counter() - t0
        if a.iters > 1:
            print(f"rank {rank} q3 iteration {it}: {dt * 1e3:.2f} ms; " + ", ".join(f"{k} {v:.2f}" for k, v in stats["host_ms_per_call"].items()), flush=True)
    rows_in = cust.nrows + orders.nrows + line.nrows
    npeer, nnccl = ctx.profile_get("peer_gather")[1], ctx.profile_get("alltoall")[1]
    ctx.profile(False)
    if rank == 0:
        print(f"transport: peer windows x{npeer}, nccl all-to-all x{nnccl}", flush=True)

    gathered = [None] * world
    dist.all_gather_object(gathered, (keys, aggs, rows_in, stats["redistributed_custkey"], stats["redistributed_orderkey"]))
    if rank == 0:
        allk = np.concatenate([x[0] for x in gathered]); alla = np.concatenate([x[1] for x in gathered])
        total_rows = sum(x[2] for x in gathered)
        print(f"q3: {world} datanodes, {total_rows} input rows, {len(allk)} groups, {dt * 1e3:.2f} ms, "
              f"{total_rows / dt / 1e9:.2f} G rows/s; redistributed {sum(x[3] for x in gathered)} + {sum(x[4] for x in gathered)} rows")
        if a.check:
            import oracle as O
            want = O.q3_reference(a.sf, nord, ncust, P.DATE_Q3, P.SEGMENT_Q3)
            order = np.lexsort([allk[:, 2], allk[:, 1], allk[:, 0]])
            np.testing.assert_array_equal(allk[order], want.keys)
            np.testing.assert_allclose(alla[order, 0], want.aggs[:, 0], rtol=1e-9, atol=0)
            print(f"OK q3 parity vs oracle: {want.ngroups} groups")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
