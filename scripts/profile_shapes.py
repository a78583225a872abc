"""Runs the config-1, Q1 and Q3 shapes a few times on one GPU at SF `--sf` (inputs resident) so that
ncu can capture their kernels:
    ncu --set full --clock-control none --import-source on -k regex:'gx_k_(count_char|fewgroups|runagg)' -c 6 \
        -o gpurun_out/shapes python scripts/profile_shapes.py --sf 100
Prints CUDA-event times per shape without ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g  # noqa: E402
from opentenbase_b200 import plans as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, default=100)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--shapes", default="config1,q1,q3,config2,config3")
    ap.add_argument("--envs", default="", help="A/B runs: ';'-separated settings, each a ','-separated list of NAME=VALUE "
                                               "(the library reads its switches per call); an empty setting is the default build")
    a = ap.parse_args()
    ctx = g.Context(0)
    ctx.pool_reserve(0)
    no = 1_500_000 * a.sf
    lt = ctx.table(g.SCHEMAS[g.T_LINEITEM], no * 4 + no // 8).generate(g.T_LINEITEM, a.sf, 0, no)
    ot = ctx.table(g.SCHEMAS[g.T_ORDERS], no).generate(g.T_ORDERS, a.sf, 0, no)
    shapes = {}
    shapes["config1"] = lambda: ctx.hash_agg(lt, P.config1_plan(g.L_RETURNFLAG))
    shapes["config2"] = lambda: ctx.hash_agg(lt, P.config2_plan(g.L_SHIPDATE, g.L_EXTENDEDPRICE))
    shapes["q1"] = lambda: ctx.hash_agg(lt, P.q1_plan(g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_SHIPDATE, g.L_RETURNFLAG, g.L_LINESTATUS))

    def q3():
        h = ctx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE, g.O_SHIPPRIORITY], unique=True, preds=[(g.O_ORDERDATE, g.GX_LT, P.DATE_Q3)])
        r = ctx.hash_agg(lt, P.q3_agg_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_SHIPDATE, est_groups=no // 8), h)
        h.free()
        return r
    shapes["q3"] = q3

    def c3():
        h = ctx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
        r = ctx.hash_agg(lt, P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE), h)
        h.free()
        return r
    shapes["config3"] = c3
    if "q3chain" in a.shapes:
        ctx.set_shardmap(1)
        ct = ctx.table(g.SCHEMAS[g.T_CUSTOMER], no // 10).generate(g.T_CUSTOMER, a.sf, 0, no // 10)
        cc = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
        oc = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
        lc = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}
        shapes["q3chain"] = lambda: P.q3_datanode(ctx, ct, ot, lt, cc, oc, lc)
    for setting, name in [(e, n) for e in a.envs.split(";") for n in a.shapes.split(",")]:
        pairs = [kv.split("=", 1) for kv in setting.split(",") if kv]
        for k, v in pairs:
            os.environ[k] = v
        _run(ctx, name, shapes[name], a.iters, setting)
        for k, _ in pairs:
            del os.environ[k]
    ctx.close()


def _run(ctx, name, fn, iters, setting):
    if True:
        fn().free()
        ctx.sync()
        ctx.profile(True)
        ctx.timer_start()
        for _ in range(iters):
            r = fn()
            n = r.ngroups
            r.free()
        ms = ctx.timer_stop() / iters
        ph = {}
        for lab in ("agg", "probe_agg", "runagg", "build", "build_scatter", "agg_compact", "runagg_merge", "filter", "partition", "probe"):
            t, k = ctx.profile_get(lab)
            if k:
                ph[lab] = round(t / iters, 3)
        ctx.profile(False)
        print(f"{name} [{setting or 'default'}]: {ms:.3f} ms per call, {n} groups, kernels {ph}", flush=True)


if __name__ == "__main__":
    main()
