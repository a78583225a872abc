"""Development probe (not the bench contract): times the hot kernels per
strategy and scale on one GPU and prints one JSON line per measurement.
Usage: python scripts/gpu_probe.py [--sf 10,100] [--out gpurun_out/probe.jsonl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opentenbase_b200 as g  # noqa: E402


def timed(ctx, fn, reps=5, flush=True):
    best, times = None, []
    for _ in range(reps):
        if flush:
            ctx.l2_flush()
        ctx.timer_start()
        r = fn()
        ms = ctx.timer_stop()
        times.append(ms)
        if hasattr(r, "free"):
            keep = r
    return min(times), float(np.median(times)), keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", default="10,100")
    ap.add_argument("--out", default="gpurun_out/probe.jsonl")
    ap.add_argument("--strategies", default="1,3,2")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    out = open(a.out, "a")
    ctx = g.Context(0)

    def emit(**kw):
        line = json.dumps(kw)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()

    emit(event="device", **ctx.device_info())
    for sf in [int(x) for x in a.sf.split(",")]:
        no = 1_500_000 * sf
        t0 = time.time()
        ot = ctx.table([g.GX_INT8, g.GX_DATE], no)
        ot.generate(g.T_ORDERS, sf, 0, no, colmap=[g.O_ORDERKEY, g.O_ORDERDATE])
        lt = ctx.table([g.GX_INT8, g.GX_FLOAT8, g.GX_DATE], int(no * 4.02) + 1024)
        lt.generate(g.T_LINEITEM, sf, 0, no, colmap=[g.L_ORDERKEY, g.L_EXTENDEDPRICE, g.L_SHIPDATE])
        nl = lt.nrows
        emit(event="generated", sf=sf, orders=no, lineitem=nl, secs=round(time.time() - t0, 2))
        aggs = [(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])]

        # config 2: SUM(l_extendedprice) GROUP BY l_shipdate — 12 B/row algorithmic
        for s in [int(x) for x in a.strategies.split(",")]:
            plan = g.make_plan(group_cols=[(0, 2)], aggs=[aggs[1]], est_groups=2600, strategy=s)
            try:
                best, med, r = timed(ctx, lambda: ctx.hash_agg(lt, plan))
                emit(event="config2", sf=sf, strategy=s, ms_best=round(best, 3), ms_med=round(med, 3),
                     grows_per_s=round(nl / best / 1e6, 2), alg_GBps=round(nl * 12 / best / 1e6, 1), groups=r.ngroups)
            except g.GxError as e:
                emit(event="config2", sf=sf, strategy=s, error=str(e))

        # config 3: build orders, probe lineitem, GROUP BY o_orderdate
        ctx.profile(True)
        bbest, bmed, ht = timed(ctx, lambda: ctx.hash_build(ot, 0, [1], unique=True), reps=3)
        emit(event="build_breakdown", sf=sf, **{k: round(ctx.profile_get(k)[0] / max(ctx.profile_get(k)[1], 1), 3)
                                                for k in ("build_count", "build_scan", "build_scatter", "build", "build_clear")})
        ctx.profile(False)
        emit(event="build", sf=sf, ms_best=round(bbest, 3), ms_med=round(bmed, 3), nslots=ht.nslots,
             alg_GBps=round(no * 24 / bbest / 1e6, 1))
        for s in [int(x) for x in a.strategies.split(",")]:
            plan = g.make_plan(outer_key_col=0, group_cols=[(1, 0)], aggs=aggs, est_groups=2500, strategy=s)
            try:
                best, med, r = timed(ctx, lambda: ctx.hash_agg(lt, plan, ht), reps=3)
                emit(event="config3_probe_agg", sf=sf, strategy=s, ms_best=round(best, 3), ms_med=round(med, 3),
                     grows_per_s=round(nl / best / 1e6, 2), alg_GBps=round(nl * 24 / best / 1e6, 1), groups=r.ngroups,
                     whole_query_grows_per_s=round((nl + no) / (best + bbest) / 1e6, 2))
            except g.GxError as e:
                emit(event="config3_probe_agg", sf=sf, strategy=s, error=str(e))
        # count(*) only and join-only variants to separate probe cost from aggregation cost
        plan = g.make_plan(outer_key_col=0, aggs=[aggs[0]], strategy=1)
        best, med, r = timed(ctx, lambda: ctx.hash_agg(lt, plan, ht), reps=3)
        emit(event="config3_probe_countstar_1group", sf=sf, ms_best=round(best, 3), grows_per_s=round(nl / best / 1e6, 2),
             alg_GBps=round(nl * 16 / best / 1e6, 1))
        plan = g.make_plan(aggs=[aggs[1]], strategy=1)
        best, med, r = timed(ctx, lambda: ctx.hash_agg(lt, plan), reps=3)
        emit(event="scan_sum_1group", sf=sf, ms_best=round(best, 3), alg_GBps=round(nl * 8 / best / 1e6, 1))
        ht.free(); ot.free(); lt.free()
    ctx.close()


if __name__ == "__main__":
    main()
