"""Small driver for ncu captures: one build + a few probe_agg launches at the given SF."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g
sf = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = g.Context(0)
no = 1_500_000 * sf
ot = ctx.table([g.GX_INT8, g.GX_DATE], no); ot.generate(g.T_ORDERS, sf, 0, no, colmap=[g.O_ORDERKEY, g.O_ORDERDATE])
lt = ctx.table([g.GX_INT8, g.GX_FLOAT8, g.GX_DATE], int(no * 4.02) + 1024)
lt.generate(g.T_LINEITEM, sf, 0, no, colmap=[g.L_ORDERKEY, g.L_EXTENDEDPRICE, g.L_SHIPDATE])
aggs = [(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])]
for _ in range(reps):
    ht = ctx.hash_build(ot, 0, [1], unique=True)
    r = ctx.hash_agg(lt, g.make_plan(outer_key_col=0, group_cols=[(1, 0)], aggs=aggs, est_groups=2500, strategy=1), ht)
    r2 = ctx.hash_agg(lt, g.make_plan(group_cols=[(0, 2)], aggs=[aggs[1]], est_groups=2600, strategy=1))
    ctx.sync()
print("groups", r.ngroups, r2.ngroups)
