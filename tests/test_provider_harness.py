"""The CustomScan provider EXECUTED: provider/harness/gpuexec_harness is a stub-linked fake backend
(fake EState, heapgetpage() over oracle-built heap pages, real libgpuexec.so) that drives
CreateCustomScanState / BeginCustomScan / ExecCustomScan / ReScan / EndCustomScan and prints the
slots it gets back.  The rows are compared with the oracle's tuple-at-a-time executor:
final values when the plan is pushed down whole, transition states (N, Sx) when the provider
runs the partial half of a two-phase aggregate.  The binary is built where /root/reference
exists (make -C opentenbase_b200/provider harness) and travels to the GPU box."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import opentenbase_b200 as g
from opentenbase_b200 import plans as P
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "opentenbase_b200", "provider", "harness", "gpuexec_harness")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(HARNESS), reason="harness binary not built (needs /root/reference at build time)")]

# oracle type -> (attlen, attalign, pg_type oid)
PG = {O.GX_INT4: (4, 4, 23), O.GX_INT8: (8, 8, 20), O.GX_FLOAT8: (8, 8, 701), O.GX_DATE: (4, 4, 1082), O.GX_CHAR: (1, 1, 18),
      O.ORC_BPCHAR1: (-1, 4, 1042)}
GXT = {O.GX_INT4: g.GX_INT4, O.GX_INT8: g.GX_INT8, O.GX_FLOAT8: g.GX_FLOAT8, O.GX_DATE: g.GX_DATE, O.GX_CHAR: g.GX_CHAR, O.ORC_BPCHAR1: g.GX_CHAR}
INT8OID, FLOAT8OID, FLOAT8ARRAYOID = 20, 701, 1022


def write_case(path, rels, outer, inner, plan, partial, out_types, notnull=True):
    """rels: [(oracle types, O.Rel)]; notnull: the columns are declared NOT NULL (as in the TPC-H DDL); outer/inner: dict(rti, attnums); inner also key_col, payload_cols, unique, preds"""
    with open(path, "wb") as f:
        f.write(b"GXH1")
        f.write(struct.pack("<i", len(rels)))
        for types, rel in rels:
            f.write(struct.pack("<i", len(types)))
            for t in types:
                f.write(struct.pack("<iiii", *PG[t], 1 if notnull else 0))
            f.write(struct.pack("<q", rel.npages))
            f.write(struct.pack("<f", float(rel.ntuples)))
            f.write(rel.pages().tobytes())

        def relinfo(d, types):
            f.write(struct.pack("<ii", d["rti"], len(d["attnums"])))
            for a in d["attnums"]:
                f.write(struct.pack("<ii", a, GXT[types[a]]))
        f.write(struct.pack("<i", 1 if inner else 0))
        relinfo(outer, rels[outer["rti"] - 1][0])
        if inner:
            relinfo(inner, rels[inner["rti"] - 1][0])
            f.write(struct.pack("<ii", inner["key_col"], len(inner["payload_cols"])))
            for c in inner["payload_cols"]:
                f.write(struct.pack("<i", c))
            f.write(struct.pack("<ii", 1 if inner["unique"] else 0, len(inner.get("preds", []))))
            for pr in inner.get("preds", []):
                f.write(bytes(g.mk_pred(*pr)))
        f.write(struct.pack("<i", 1 if partial else 0))
        f.write(bytes(plan))
        f.write(struct.pack("<i", len(out_types)))
        for t in out_types:
            f.write(struct.pack("<i", t))


def run_harness(case):
    r = subprocess.run([HARNESS, case], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:] + r.stderr[-3000:])
    rows = [line.split("\t") for line in r.stdout.splitlines() if line]
    return rows, r.stderr


def test_harness_config3_join_pushed_down_whole():
    """lineitem JOIN orders GROUP BY o_orderdate with an inner-side qual: final values out of ExecCustomScan"""
    sf, nord = 1, 60000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    otypes = [O.GX_INT8, O.GX_INT4, O.GX_DATE, O.GX_INT4]
    ltypes = [O.GX_INT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_DATE, O.ORC_BPCHAR1, O.ORC_BPCHAR1]   # bpchar(1) flags as in the TPC-H DDL
    orel, lrel = O.Rel(otypes, o), O.Rel(ltypes, l)
    # the provider stages only referenced attributes: outer (l_orderkey, l_extendedprice), inner (o_orderkey, o_orderdate)
    plan = g.make_plan(outer_key_col=0, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=2500)
    with tempfile.TemporaryDirectory() as d:
        case = os.path.join(d, "c3.case")
        write_case(case, [(ltypes, lrel), (otypes, orel)], {"rti": 1, "attnums": [0, 2]},
                   {"rti": 2, "attnums": [0, 2], "key_col": 0, "payload_cols": [1], "unique": True, "preds": [(1, g.GX_LT, -1752)]}, plan, False,
                   [1082, INT8OID, FLOAT8OID, FLOAT8OID])
        rows, err = run_harness(case)
    oplan = O.make_plan(outer_key_col=g.L_ORDERKEY, group_cols=[(1, 0)],
                        aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, g.L_EXTENDEDPRICE, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, g.L_EXTENDEDPRICE, 0)])])
    want = O.exec_agg(lrel, oplan, orel, O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1,
                                                     inner_preds=[(g.O_ORDERDATE, g.GX_LT, -1752)])).sorted()
    got = sorted((int(r[0]), int(r[1]), float(r[2]), float(r[3])) for r in rows)
    assert len(got) == want.ngroups > 100
    np.testing.assert_array_equal([x[0] for x in got], want.keys[:, 0])
    np.testing.assert_array_equal([x[1] for x in got], want.aggs[:, 0].view(np.int64))
    np.testing.assert_allclose([x[2] for x in got], want.aggs[:, 1], rtol=1e-9, atol=0)
    np.testing.assert_allclose([x[3] for x in got], want.aggs[:, 2], rtol=1e-9, atol=0)
    assert "GPU Output: final values" in err


def test_harness_q1_partial_states_for_a_two_phase_plan():
    """Q1 shape with a qual, GROUP BY two bpchar(1) columns, partial mode: the slots carry what the reference's
    combine functions expect — int8 counts, float8 sums, float8[3] {N, Sx, Sxx} for avg — checked against the
    oracle through a restated Finalize (float8_avg = Sx / N, float.c:2991)."""
    sf, nord = 1, 40000
    l = O.gen_lineitem(sf, 0, nord)
    ltypes = [O.GX_INT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_DATE, O.ORC_BPCHAR1, O.ORC_BPCHAR1]
    lrel = O.Rel(ltypes, l)
    attnums = [1, 2, 3, 4, 5, 6, 7]                       # qty, price, disc, tax, shipdate, flag, status
    plan = P.q1_plan(0, 1, 2, 3, 4, 5, 6)
    out_types = [1042, 1042, FLOAT8OID, FLOAT8OID, FLOAT8OID, FLOAT8OID, FLOAT8ARRAYOID, FLOAT8ARRAYOID, FLOAT8ARRAYOID, INT8OID]
    with tempfile.TemporaryDirectory() as d:
        case = os.path.join(d, "q1.case")
        write_case(case, [(ltypes, lrel)], {"rti": 1, "attnums": attnums}, None, plan, True, out_types)
        rows, err = run_harness(case)
    want = O.exec_agg(lrel, P.q1_plan(g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_SHIPDATE, g.L_RETURNFLAG, g.L_LINESTATUS,
                                      maker=O.make_plan)).sorted()
    assert len(rows) == want.ngroups == 4 and "partial states" in err
    rows.sort(key=lambda r: (int(r[0]), int(r[1])))
    for r, wk, wa in zip(rows, want.keys, want.aggs):
        assert (int(r[0]), int(r[1])) == (wk[0], wk[1])
        n = int(r[9])
        assert n == wa[7:8].view(np.int64)[0]                                   # count(*) bit-exact
        np.testing.assert_allclose([float(x) for x in r[2:6]], wa[0:4], rtol=1e-9, atol=0)   # sums
        for col, a in ((6, 4), (7, 5), (8, 6)):                                # avg states -> float8_avg
            N, sx, sxx = (float(x) for x in r[col].strip("{}").split(","))
            assert N == n and sxx == 0.0
            np.testing.assert_allclose(sx / N, wa[a], rtol=1e-9, atol=0)
