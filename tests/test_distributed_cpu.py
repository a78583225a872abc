"""CPU coverage of the N > 1 path with world_size-2 gloo processes.

The data path itself (partition kernel + NCCL all-to-all) needs GPUs; what runs
here is the host-side logic it relies on: SHARD placement of the generator's
rows on datanodes, the Partial -> redistribute -> Finalize aggregation contract
(each group is finalized on exactly one datanode, the union equals the one-node
answer), and bench.py's launch contract for the reference arm under torchrun."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import opentenbase_b200 as g
import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
sf, nord = 1, 5000
o, l = O.gen_orders(sf, 0, nord, rank, world), O.gen_lineitem(sf, 0, nord, rank, world)
assert (O.route_nodes(o[0], O.GX_INT8, world) == rank).all()          # SHARD placement
plan = O.make_plan(outer_key_col=g.L_ORDERKEY, group_cols=[(1, 0)],
                   aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, g.L_EXTENDEDPRICE, 0)]),
                         (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, g.L_QUANTITY, 0)])])
join = O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1)
part = O.exec_agg(O.Rel(g.SCHEMAS[g.T_LINEITEM], l), plan, O.Rel(g.SCHEMAS[g.T_ORDERS], o), join)   # Partial HashAggregate
# "Distribute results by S: o_orderdate": every partial group goes to the datanode owning its key
dest = O.route_nodes(part.keys[:, 0].astype(np.int32), O.GX_DATE, world)
outbox = [{"keys": part.keys[dest == d].tolist(), "states": part.states[dest == d].tolist()} for d in range(world)]
inbox = [None] * world
dist.all_to_all_object = None
gathered = [None] * world
dist.all_gather_object(gathered, outbox)
mine = [gathered[src][rank] for src in range(world)]
# Finalize HashAggregate: int8pl / float8pl / float8_combine over the received states
acc = {}
for m in mine:
    for k, st in zip(m["keys"], m["states"]):
        a = acc.setdefault(k[0], [0, 0.0, None, [0.0, 0.0, 0.0]])
        a[0] += int(np.float64(st[0][0]).view(np.int64))
        a[1] = st[1][1] if a[2] is None else a[1] + st[1][1]; a[2] = True
        N1, S1 = a[3][0], a[3][1]; N2, S2 = st[2][0], st[2][1]
        a[3][0], a[3][1] = N1 + N2, (S2 if N1 == 0 else S1 + S2)
final = {k: (v[0], v[1], v[3][1] / v[3][0]) for k, v in acc.items()}
allfinal = [None] * world
dist.all_gather_object(allfinal, final)
if rank == 0:
    assert not (set(allfinal[0]) & set(allfinal[1]))                   # each group finalized on exactly one datanode
    merged = {**allfinal[0], **allfinal[1]}
    whole = O.exec_agg(O.Rel(g.SCHEMAS[g.T_LINEITEM], O.gen_lineitem(sf, 0, nord)), plan,
                       O.Rel(g.SCHEMAS[g.T_ORDERS], O.gen_orders(sf, 0, nord)), join)
    assert len(merged) == whole.ngroups
    for k, c, s, a in zip(whole.keys[:, 0], whole.aggs[:, 0].view(np.int64), whole.aggs[:, 1], whole.aggs[:, 2]):
        mc, ms, ma = merged[int(k)]
        assert mc == c and abs(ms - s) <= 1e-9 * abs(s) and abs(ma - a) <= 1e-9 * abs(a)
    print("OK", len(merged))
dist.destroy_process_group()
'''


def test_partial_redistribute_finalize_two_datanodes(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": _free_port()})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    assert "OK" in outs[0][0]


def test_bench_reference_arm_under_torchrun():
    """`bench.py --impl reference` launched like the driver does for N = 2: rank 0 alone prints the line."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "3", "--cpu-sample-orders", "20000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["value"] > 0 and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
