#!/usr/bin/env python
"""bench.py — rows/sec of the lineitem JOIN orders hash join + GROUP BY at SF100
per GPU (BASELINE.json metric), one process per GPU.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):

    SELECT o_orderdate, count(*), sum(l_extendedprice)
    FROM lineitem JOIN orders ON l_orderkey = o_orderkey GROUP BY o_orderdate

on synthetic TPC-H-shaped tables (include/gx_tpch_gen.h, seed 20240922).  A
"step" is one complete pass of the hot path over the datanode's resident
tables: hash build over orders, fused probe + hash aggregate over lineitem,
Finalize across datanodes (N > 1), result read back.  With N GPUs there are N
datanodes holding a SF(100*N) database placed by the reference's SHARD rule
(weak scaling; lineitem and orders are co-located on the order key, so the only
exchange is the partial-aggregate redistribute).

    python bench.py --gpus 1 --steps 10 --warmup 3
    torchrun --nproc-per-node N bench.py --gpus N ...
    python bench.py --impl reference        # the CPU executor port on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rows/sec lineitem JOIN orders hashjoin+groupby SF100"
try:                                   # BASELINE.json names the metric; use its wording when it is there
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"] or METRIC
except (OSError, ValueError, KeyError):
    pass
ALG_BYTES_PER_PROBE_ROW = 24     # 8 B key + 8 B payload on hit + 8 B l_extendedprice
ALG_BYTES_PER_BUILD_ROW = 24     # 8 B key read + 16 B slot write


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------ clocks
REAL_STDOUT = sys.stdout


class ClockSampler:
    """nvidia-smi clocks / throttle reasons.  The sampler needs a few hundred ms to come up
    and the timed region is short, so it is started before the warm-up steps (the GPU is under
    the same load there); stop(t0, t1) reports the samples that fall inside the timed region
    and, when the region was too short to catch any, the ones taken under load since start."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not inside:
            inside = [r for r in rows if t1 is None or r[0] <= t1]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        if inside:
            reasons = set()
            for r in inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            out = {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": max(r[2] for r in inside),
                   "reasons": sorted(reasons), "samples": len(inside), "window": window}
        return out


# ------------------------------------------------------- CPU baseline (oracle)
def cpu_reference_run(sample_orders: int, threads: int, steps: int, warmup: int):
    """The reference CPU executor's path (oracle/: tuple-at-a-time SeqScan ->
    HashJoin -> HashAgg, the only stand-in that exists: the reference itself
    cannot be built here) on `threads` host threads.  Each
    thread owns one co-partitioned slice (like one datanode / parallel worker),
    partial aggregates are combined at the end (Partial -> Finalize)."""
    import oracle as O
    import opentenbase_b200 as g
    from concurrent.futures import ThreadPoolExecutor
    sf = 100
    bounds = np.linspace(0, sample_orders, threads + 1).astype(np.int64)
    plan = O.make_plan(outer_key_col=g.L_ORDERKEY, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=2500)
    join = O.make_join(0, payload_cols=[1], inner_unique=1)

    def make(i):
        o = O.gen_orders(sf, int(bounds[i]), int(bounds[i + 1]))
        l = O.gen_lineitem(sf, int(bounds[i]), int(bounds[i + 1]))
        orel = O.Rel([O.GX_INT8, O.GX_DATE], [o[0], o[2]])
        lrel = O.Rel([O.GX_INT8, O.GX_FLOAT8], [l[0], l[2]])
        return orel, lrel, len(o[0]) + len(l[0])

    with ThreadPoolExecutor(threads) as ex:
        rels = list(ex.map(make, range(threads)))
        nrows = sum(r[2] for r in rels)

        def one(i):
            res, raw = O.exec_agg(rels[i][1], plan, rels[i][0], join, keep_raw=True)
            return raw

        times = []
        final = None
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            raws = list(ex.map(one, range(threads)))
            final = O.combine(plan, raws)
            dt = time.perf_counter() - t0
            for r in raws:
                O.lib().orc_result_free(r)
            if it >= warmup:
                times.append(dt)
    total_count = int(final.aggs[:, 0].view(np.int64).sum())
    return {"rows": nrows, "secs_per_step": float(np.mean(times)), "rows_per_sec": nrows / float(np.mean(times)),
            "count_star_total": total_count, "groups": final.ngroups}


def effective_cpus() -> int:
    """Host threads this process can really run at once: the affinity mask, capped by a cgroup
    CPU quota if there is one (a container may show 128 CPUs and be allowed 16 of them)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]              # cgroup v2
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())            # cgroup v1
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0 and period > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, n)


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = effective_cpus()
    sample_orders = (args.cpu_sample_orders // 2) * threads          # ~0.75 M orders (3.75 M rows) per thread
    r = cpu_reference_run(sample_orders, threads, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rows_per_sec"], "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["secs_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": r["rows_per_sec"], "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_orders} orders + their lineitem rows ({r['rows']} rows) of the SF100 tables, "
                                   f"{threads} threads, one co-partitioned slice each, partial->final combine"},
        "e2e": {"value": r["rows_per_sec"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=REAL_STDOUT, flush=True)
    return 0


def workload_config(args, world):
    return {"workload": "configs[2]: lineitem JOIN orders ON l_orderkey=o_orderkey, hash build on orders, probe lineitem, "
                        "GROUP BY o_orderdate: count(*), sum(l_extendedprice)",
            "sf_per_gpu": args.sf, "sf_total": args.sf * world, "datanodes": world,
            "parallelism": f"{world} datanode(s), one per GPU, SHARD placement on the order key",
            "l2": "inputs (>= 11 GB per GPU at SF100) are far larger than the 126 MB L2; no flush needed between steps"}


# ---------------------------------------------------------------- GPU arm
def run_ours(args):
    import opentenbase_b200 as g
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version banner to stdout otherwise
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = g.Context(local_rank)
    if world > 1:
        box = [g.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    ctx.set_shardmap(world)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sf_total = args.sf * world
    n_orders_total = 1_500_000 * sf_total
    cap_o = int(1_500_000 * args.sf * 1.03) + 1_000_000
    cap_l = int(6_000_000 * args.sf * 1.03) + 4_000_000
    ot = ctx.table([g.GX_INT8, g.GX_DATE], cap_o)
    lt = ctx.table([g.GX_INT8, g.GX_FLOAT8], cap_l)
    ot.generate(g.T_ORDERS, sf_total, 0, n_orders_total, rank, world, colmap=[g.O_ORDERKEY, g.O_ORDERDATE])
    lt.generate(g.T_LINEITEM, sf_total, 0, n_orders_total, rank, world, colmap=[g.L_ORDERKEY, g.L_EXTENDEDPRICE])
    no, nl = ot.nrows, lt.nrows
    plan = g.make_plan(outer_key_col=0, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=2500)

    def step():
        ht = ctx.hash_build(ot, 0, [1], unique=True)
        r = ctx.hash_agg(lt, plan, ht)
        r.combine()
        out = r.fetch()
        r.free(); ht.free()
        return out

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    # ---- timed region: exactly K steps, device events on the library's stream
    ctx.profile(True)
    launches0 = ctx.launches
    barrier()
    t0 = time.perf_counter(); wall0 = time.time()
    ctx.timer_start()
    last = None
    for _ in range(args.steps):
        last = step()
    dev_ms = ctx.timer_stop()
    ctx.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3; wall1 = time.time()
    barrier()
    launches = ctx.launches - launches0
    clk = clocks.stop(wall0, wall1) if clocks else None
    probe_ms, probe_n = ctx.profile_get("probe_agg")
    build_ms, build_n = ctx.profile_get("build")
    phases = {}
    for name in ("build_sample", "build_bounds", "build_scatter", "build", "build_clear", "probe_agg", "agg", "agg_compact",
                 "radix_partition", "radix_agg", "combine_partition", "alltoall"):
        ms, n = ctx.profile_get(name)
        if n:
            phases[name] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps}
    ctx.profile(False)
    dev_ms_max = allmax(dev_ms)
    wall_ms_max = allmax(wall_ms)
    rows_all = allsum(float(no + nl))
    ms_per_step = dev_ms_max / args.steps
    value = rows_all / (ms_per_step / 1e3)
    # size-independent correctness properties on the last step's result
    keys, aggs, nulls = last
    count_local = int(aggs[:, 0].view(np.int64).sum())
    count_total = allsum(float(count_local))
    nl_total = allsum(float(nl))
    checks = {"count_star_equals_lineitem_rows": int(count_total) == int(nl_total),
              "groups_this_node": int(len(keys))}

    # ---- end to end: HOST (pinned) buffers -> result, copies inside the timed region
    e2e = None
    try:
        e2e = run_e2e(ctx, g, ot, lt, no, nl, plan, args, barrier, allmax, rows_all)
    except g.GxError as ex:
        e2e = {"value": None, "unit": "rows/s", "error": str(ex)}

    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    probe_avg_ms = probe_ms / max(probe_n, 1)
    achieved = nl * ALG_BYTES_PER_PROBE_ROW / (probe_avg_ms / 1e3) / 1e9 if probe_n else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "probe_agg_traffic.json"))).get("dram_bytes_per_launch_sf100")
    except (OSError, ValueError):
        pass
    roofline = {"kernel": "gx_k_runjoin (fused hash probe + hash aggregate over lineitem)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nl * ALG_BYTES_PER_PROBE_ROW, "avg_launch_ms": probe_avg_ms,
                "build_kernel_avg_ms": build_ms / max(build_n, 1),
                "build_alg_GBps": (no * ALG_BYTES_PER_BUILD_ROW / (build_ms / max(build_n, 1) / 1e3) / 1e9) if build_n else None}

    cpu = None
    if True:
        r = cpu_reference_run(args.cpu_sample_orders, 1, 1, 0)
        cpu = {"value": r["rows_per_sec"], "unit": "rows/s", "cores": 1, "kind": "port",
               "sample": f"first {args.cpu_sample_orders} orders and their lineitem rows ({r['rows']} rows) of the SF100 tables, "
                         "oracle/ tuple-at-a-time executor, 1 thread (one backend per datanode fragment)"}

    line = {"metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": workload_config(args, world), "rows_per_step": rows_all,
            "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "checks": checks,
            "phases_ms": phases}
    print(json.dumps(line), file=REAL_STDOUT, flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_e2e(ctx, g, ot, lt, no, nl, plan, args, barrier, allmax, rows_all):
    import ctypes as C
    sizes = [(ot, 0, 8, no), (ot, 1, 4, no), (lt, 0, 8, nl), (lt, 1, 8, nl)]
    bufs = []
    for t, col, sz, n in sizes:
        p = ctx.host_alloc(max(n * sz, 8))
        ctx._chk(g.lib().gx_table_read_column(t.h, col, 0, n, p, None))     # fill the pinned staging buffer
        bufs.append(p)
    h2d = sum(sz * n for _, _, sz, n in sizes)
    steps = max(1, min(args.steps, args.e2e_steps))

    def estep():
        r = ctx.exec_host([g.GX_INT8, g.GX_FLOAT8], bufs[2:4], nl, plan, [g.GX_INT8, g.GX_DATE], bufs[0:2], no,
                          inner_key_col=0, payload_cols=[1], inner_unique=True)
        r.combine()
        out = r.fetch()
        r.free()
        return out

    estep()                                   # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = estep()
    ctx.sync()
    dt = time.perf_counter() - t0
    barrier()
    dt = allmax(dt)
    d2h = int(out[0].nbytes + out[1].nbytes + out[2].nbytes)
    placement = staging_placement(bufs)
    for p in bufs:
        ctx.host_free(p)
    return {"value": rows_all / (dt / steps), "unit": "rows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": dt / steps * 1e3, "h2d_gb_per_s": h2d / (dt / steps) / 1e9, "staging": placement,
            "note": "gx_exec_host: pinned host columns -> HBM -> build -> probe+agg -> result on the host, every step"}


def staging_placement(bufs):
    """Where the pinned staging buffers ended up (pages per NUMA node, from /proc/self/numa_maps) and the
    node the GPU hangs off: the e2e figure halves when the two differ."""
    info = {}
    try:
        want = {int(p) for p in bufs}
        pages = {}
        for line in open("/proc/self/numa_maps"):
            f = line.split()
            if int(f[0], 16) in want:
                for tok in f[1:]:
                    if tok[0] == "N" and "=" in tok:
                        node, n = tok[1:].split("=")
                        pages[node] = pages.get(node, 0) + int(n)
        info["pages_by_node"] = pages
    except (OSError, ValueError, IndexError):
        pass
    try:
        bus = subprocess.check_output(["nvidia-smi", f"--id={env_int('LOCAL_RANK', 0)}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, stderr=subprocess.DEVNULL).strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus                     # 00000000:1B:00.0 -> 0000:1b:00.0
        info["gpu_numa_node"] = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sf", type=int, default=100, help="scale factor per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-orders", type=int, default=1_500_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly one JSON line: anything a library prints on fd 1 meanwhile (NCCL's
    # version banner, for one) is sent to stderr instead
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
