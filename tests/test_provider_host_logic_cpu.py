"""The provider's HOST logic executed without a GPU: provider/harness/gpuexec_harness_double is the stub-linked fake
backend of tests/test_provider_harness.py linked against a TEST DOUBLE of the C ABI (harness/gx_double.c) instead of
libgpuexec.so.  The double computes nothing — gx_hash_agg() hands back the groups this test wrote into a file — so
what runs for real is the provider: BeginCustomScan, the heap-page loader (heap_beginscan / heapgetpage over page images,
batches of at most 4096 pages through the two-slot staging ring, visibility lists, NOT NULL hints, referenced attributes
only), the order of the C-ABI calls, ExecCustomScan's Datum encoding of every key and aggregate type in final and in
partial (transition-state) mode, NULL flags, ReScan, EndCustomScan's bookkeeping, and the ERROR path.

The GPU run of the same harness (tests/test_provider_harness.py) checks the values; this one checks everything around
them, on every box."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import opentenbase_b200 as g
import oracle as O
from test_provider_harness import write_case, HARNESS

DOUBLE = HARNESS + "_double"
pytestmark = pytest.mark.skipif(not os.path.exists(DOUBLE), reason="harness binary not built (needs /root/reference at build time)")

OTYPES = [O.GX_INT8, O.GX_INT4, O.GX_DATE, O.GX_INT4]
LTYPES = [O.GX_INT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_DATE, O.ORC_BPCHAR1, O.ORC_BPCHAR1]
INT4OID, INT8OID, FLOAT8OID, DATEOID, BPCHAROID, F8ARRAY = 23, 20, 701, 1082, 1042, 1022


def canned(path, keys, aggs, cnts, nulls):
    keys, aggs, cnts, nulls = (np.ascontiguousarray(x) for x in (keys, aggs, cnts, nulls))
    n = keys.shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<qii", n, keys.shape[1], aggs.shape[1]))
        f.write(keys.astype(np.int64).tobytes()); f.write(aggs.astype(np.float64).tobytes())
        f.write(cnts.astype(np.int64).tobytes()); f.write(nulls.astype(np.uint8).tobytes())


def run(case, result, env=None, expect_rc=0):
    e = dict(os.environ, GX_DOUBLE_RESULT=result, **(env or {}))
    r = subprocess.run([DOUBLE, case], capture_output=True, text=True, timeout=120, env=e)
    assert r.returncode == expect_rc, r.stdout[-1500:] + r.stderr[-3000:]
    return [l.split("\t") for l in r.stdout.splitlines() if l], r.stderr


def as_f64(int_values):
    return np.array(int_values, np.int64).view(np.float64)


def test_final_mode_every_key_and_aggregate_type_with_a_join():
    o, l = O.gen_orders(1, 0, 3000), O.gen_lineitem(1, 0, 3000)
    orel, lrel = O.Rel(OTYPES, o), O.Rel(LTYPES, l)
    # outer staged attributes: l_orderkey int8, l_extendedprice float8, l_shipdate date, l_returnflag bpchar(1)
    # GROUP BY l_orderkey, l_shipdate, l_returnflag, o_custkey (int4, join payload)
    plan = g.make_plan(outer_key_col=0, group_cols=[(0, 0), (0, 2), (0, 3), (1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)]),
                             (g.GX_AGG_MIN_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_MAX_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_COUNT, [(g.GX_OP_COL, 1, 0)])],
                       est_groups=10)
    out_types = [INT8OID, DATEOID, BPCHAROID, INT4OID, INT8OID, FLOAT8OID, FLOAT8OID, FLOAT8OID, FLOAT8OID, INT8OID]
    keys = np.array([[-2**63 + 1, -2922, ord("A"), -5],
                     [2**62 + 7, 0, 200, 2**31 - 1],             # byte 200: the key column is one raw byte
                     [0, 2406, ord("N"), 0]], np.int64)
    aggs = np.zeros((3, 6))
    aggs[:, 0] = as_f64([1, 2**40, 0]); aggs[:, 5] = as_f64([7, 0, 2**62])     # integer results travel as int64 bits
    aggs[:, 1] = [1.5e300, -0.0, 0.0]; aggs[:, 2] = [0.1, 5e-324, 0.0]
    aggs[:, 3] = [float("-inf"), 1.0, 0.0]; aggs[:, 4] = [float("nan"), 2.0, 0.0]
    nulls = np.zeros((3, 10), np.uint8)
    nulls[2, 1] = 1                                                             # a NULL group key
    nulls[2, 5] = nulls[2, 6] = nulls[2, 7] = nulls[2, 8] = 1                   # sum / avg / min / max over no non-NULL input
    with tempfile.TemporaryDirectory() as d:
        case, res = os.path.join(d, "c.case"), os.path.join(d, "c.res")
        write_case(case, [(LTYPES, lrel), (OTYPES, orel)], {"rti": 1, "attnums": [0, 2, 5, 6]},
                   {"rti": 2, "attnums": [0, 1], "key_col": 0, "payload_cols": [1], "unique": True, "preds": [(1, g.GX_GT, 17)]}, plan, False, out_types)
        canned(res, keys, aggs, np.zeros((3, 6), np.int64), nulls)
        rows, err = run(case, res)
    assert rows[0] == [str(-2**63 + 1), "-2922", "65", "-5", "1", "1.5000000000000001e+300", "0.10000000000000001", "-inf", "nan", "7"]
    assert rows[1] == [str(2**62 + 7), "0", "-56", str(2**31 - 1), str(2**40), "-0", "4.9406564584124654e-324", "1", "2", "0"]
    assert rows[2] == ["0", "\\N", "78", "0", "0", "\\N", "\\N", "\\N", "\\N", str(2**62)]
    # the calls, in order: inner relation first (its table feeds the build), then the outer one
    trace = [x for x in err.splitlines() if x.startswith("double:")]
    kinds = [x.split()[1] for x in trace]
    assert kinds[:3] == ["init", "table_create", "append_heap_pages"]
    assert kinds.index("hash_build") > [i for i, k in enumerate(kinds) if k == "load_finish"][1]      # both relations loaded before the build
    assert kinds[-4:] == ["hash_build", "hash_agg", "result_fetch", "shutdown"]
    creates = [x for x in trace if " table_create " in x]
    assert f"types {g.GX_INT8} {g.GX_INT4} " in creates[0]                                            # inner: o_orderkey, o_custkey
    assert f"types {g.GX_INT8} {g.GX_FLOAT8} {g.GX_DATE} {g.GX_CHAR} " in creates[1]                  # outer: only the referenced attributes
    appends = [x for x in trace if " append_heap_pages " in x]
    assert " attnums 0 1 " in appends[0] and " attnums 0 2 5 6 " in appends[1] and appends[1].endswith("notnull 1 1 1 1")
    fin = [x for x in trace if " load_finish " in x]
    assert f"pages {orel.npages} rows {orel.ntuples} batches 1" in fin[0] and f"pages {lrel.npages} rows {lrel.ntuples} batches 1" in fin[1]
    assert "hash_build key 0 unique 1 payload 1 preds (1 4 17)" in "\n".join(trace) or "payload 1 preds (1" in "\n".join(trace)
    assert f"hash_agg outer rows {lrel.ntuples} join 1 groups 4 aggs 6" in "\n".join(trace)
    assert "shutdown live tables 0 hashes 0 results 0" in trace[-1]                                   # EndCustomScan released every handle
    assert "harness: 3 rows" in err and "GPU Output: final values" in err


def test_partial_mode_emits_transition_states():
    """AGGSPLIT_INITIAL_SERIAL columns: int8 N for count(*) / count(x), float8 for sum / min, float8[3] {N, Sx, 0} for avg."""
    l = O.gen_lineitem(1, 0, 500)
    lrel = O.Rel(LTYPES, l)
    plan = g.make_plan(preds=[(2, g.GX_LE, -607)], group_cols=[(0, 3), (0, 4)],
                       aggs=[(g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 0, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_COUNT_STAR, []),
                             (g.GX_AGG_COUNT, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_MIN_F8, [(g.GX_OP_COL, 0, 0)])], est_groups=6)
    out_types = [BPCHAROID, BPCHAROID, FLOAT8OID, F8ARRAY, INT8OID, INT8OID, FLOAT8OID]
    keys = np.array([[ord("A"), ord("F")], [ord("N"), ord("O")]], np.int64)
    vals = np.array([[10.5, 99.25, 0.0, 0.0, -3.0], [0.0, 0.0, 0.0, 0.0, 0.0]])
    cnts = np.array([[4, 4, 5, 4, 4], [0, 0, 3, 0, 0]], np.int64)
    nulls = np.zeros((2, 7), np.uint8)
    nulls[1, 2] = nulls[1, 6] = 1                            # sum / min states stay NULL without input; avg's {0,0,0} and the counts do not
    with tempfile.TemporaryDirectory() as d:
        case, res = os.path.join(d, "p.case"), os.path.join(d, "p.res")
        write_case(case, [(LTYPES, lrel)], {"rti": 1, "attnums": [1, 2, 5, 6, 7]}, None, plan, True, out_types)
        canned(res, keys, vals, cnts, nulls)
        rows, err = run(case, res)
    assert rows[0] == ["65", "70", "10.5", "{4,99.25,0}", "5", "4", "-3"]
    assert rows[1] == ["78", "79", "\\N", "{0,0,0}", "3", "0", "\\N"]
    assert "result_fetch_states" in err and "GPU Output: partial states" in err and "hash_build" not in err


def test_loader_batches_of_4096_pages_through_the_two_slot_ring():
    nord = 600_000
    o = O.gen_orders(1, 0, nord)
    orel = O.Rel(OTYPES, o)
    assert 4096 < orel.npages < 2 * 4096
    plan = g.make_plan(group_cols=[(0, 1)], aggs=[(g.GX_AGG_COUNT_STAR, [])], est_groups=4)
    with tempfile.TemporaryDirectory() as d:
        case, res = os.path.join(d, "b.case"), os.path.join(d, "b.res")
        write_case(case, [(OTYPES, orel)], {"rti": 1, "attnums": [0, 3]}, None, plan, False, [INT4OID, INT8OID])
        canned(res, np.array([[0]], np.int64), as_f64([nord]).reshape(1, 1), np.zeros((1, 1), np.int64), np.zeros((1, 2), np.uint8))
        rows, err = run(case, res)
    assert rows == [["0", str(nord)]]
    appends = [x for x in err.splitlines() if " append_heap_pages " in x]
    assert len(appends) == 2 and " pages 4096 " in appends[0] and f" pages {orel.npages - 4096} " in appends[1]
    assert sum(int(x.split()[5]) for x in appends) == nord            # every visible tuple of every page, once
    assert f"load_finish pages {orel.npages} rows {nord} batches 2" in err
    assert f"capacity {int(nord * 1.1) + 1024}" in err                # sized from pg_class.reltuples, not from MaxHeapTuplesPerPage


def test_library_errors_become_ereport():
    """float8 overflow inside the library (GX_ERR_OVERFLOW) surfaces as an ERROR carrying the library's message;
    the resource-owner bookkeeping is what a longjmp out of ExecCustomScan leaves behind."""
    l = O.gen_lineitem(1, 0, 100)
    lrel = O.Rel(LTYPES, l)
    plan = g.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=4)
    with tempfile.TemporaryDirectory() as d:
        case, res = os.path.join(d, "e.case"), os.path.join(d, "e.res")
        write_case(case, [(LTYPES, lrel)], {"rti": 1, "attnums": [5, 2]}, None, plan, False, [DATEOID, FLOAT8OID])
        canned(res, np.zeros((1, 1), np.int64), np.zeros((1, 1)), np.zeros((1, 1), np.int64), np.zeros((1, 2), np.uint8))
        rows, err = run(case, res, env={"GX_DOUBLE_FAIL_AGG": "1"}, expect_rc=3)
    assert rows == [] and "ereport(ERROR): gpuexec: double: value out of range: overflow" in err
