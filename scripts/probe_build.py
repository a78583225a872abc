"""Development probe: build-phase breakdown at SF100 under GX_SCATTER_MODE."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g
sf = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = g.Context(0)
no = 1_500_000 * sf
ot = ctx.table([g.GX_INT8, g.GX_DATE], no); ot.generate(g.T_ORDERS, sf, 0, no, colmap=[g.O_ORDERKEY, g.O_ORDERDATE])
ctx.profile(True)
for _ in range(4):
    ht = ctx.hash_build(ot, 0, [1], unique=True); info = ht.info(); ht.free()
print(json.dumps({"mode": os.environ.get("GX_SCATTER_MODE", "0"), **info,
                  **{k: round(ctx.profile_get(k)[0] / max(ctx.profile_get(k)[1], 1), 3) for k in ("build_scatter", "build")}}))
