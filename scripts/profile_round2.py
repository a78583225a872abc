"""One pass over the round-2 kernels for an ncu capture (config 1, Q1, the Q3 chain, config 3), SF from argv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g
from opentenbase_b200 import plans as P

sf = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = g.Context(0)
ctx.set_shardmap(1)
ctx.pool_reserve(0)
no = 1_500_000 * sf
lt = ctx.table(g.SCHEMAS[g.T_LINEITEM], no * 4 + no // 8).generate(g.T_LINEITEM, sf, 0, no)
ot = ctx.table(g.SCHEMAS[g.T_ORDERS], no).generate(g.T_ORDERS, sf, 0, no)
ct = ctx.table(g.SCHEMAS[g.T_CUSTOMER], no // 10).generate(g.T_CUSTOMER, sf, 0, no // 10)
cc = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
oc = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
lc = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}
ctx.hash_agg(lt, P.config1_plan(g.L_RETURNFLAG)).free()
ctx.hash_agg(lt, P.q1_plan(g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_SHIPDATE, g.L_RETURNFLAG, g.L_LINESTATUS)).free()
P.q3_datanode(ctx, ct, ot, lt, cc, oc, lc).free()
h = ctx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
ctx.hash_agg(lt, P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE), h).free()
h.free()
ctx.sync()
print("done")
ctx.close()
