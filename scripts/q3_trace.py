"""Host-side time per C-ABI call of the Q3 fragment chain on one GPU, several iterations (diagnostics)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g
from opentenbase_b200 import plans as P

sf = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = g.Context(0)
ctx.pool_reserve(0)
ctx.set_shardmap(1)
no = 1_500_000 * sf
lt = ctx.table(g.SCHEMAS[g.T_LINEITEM], no * 4 + no // 8).generate(g.T_LINEITEM, sf, 0, no)
ot = ctx.table(g.SCHEMAS[g.T_ORDERS], no).generate(g.T_ORDERS, sf, 0, no)
ct = ctx.table(g.SCHEMAS[g.T_CUSTOMER], no // 10).generate(g.T_CUSTOMER, sf, 0, no // 10)
cc = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
oc = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
lc = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}
for it in range(6):
    st = {}
    ctx.sync(); t0 = time.perf_counter()
    r = P.q3_datanode(ctx, ct, ot, lt, cc, oc, lc, st)
    t1 = time.perf_counter()
    out = r.fetch(); t2 = time.perf_counter()
    r.free(); ctx.sync(); t3 = time.perf_counter()
    print(f"iter {it}: chain {1e3*(t1-t0):.2f} ms, fetch {1e3*(t2-t1):.2f}, free+sync {1e3*(t3-t2):.2f}; "
          + ", ".join(f"{k} {v:.2f}" for k, v in st["host_ms_per_call"].items()), flush=True)
ctx.close()
