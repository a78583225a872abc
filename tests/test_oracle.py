"""CPU tests of the oracle (the parity checker): golden vectors transcribed from
the reference's own regression outputs, known-answer hashes, and an independent
numpy cross-check of the tuple-at-a-time executor."""
import json
import os
from decimal import Decimal, getcontext

import numpy as np

import opentenbase_b200 as g
import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_sql_goldens.json")))


def s32(x):
    return x - (1 << 32) if x >= (1 << 31) else x


def test_hash_known_answers():
    L = O.lib()
    k = GOLD["hash_kats"]
    for v, want in k["hashint4"].items():
        assert s32(L.orc_hashint4(int(v))) == want
    for v, want in k["hashint4new"].items():
        assert L.orc_hashint4new(int(v)) == want
    assert L.orc_murmurhash32(L.orc_hashint4(1)) == k["murmurhash32_of_hashint4_1"]
    assert L.orc_shard_index(L.orc_hashint4(1)) == k["shard_of_int4_1"]
    assert L.orc_hash_any_new(b"123456789", 9) == k["crc32c_123456789"]
    # hashint8 folds the high half into the low one so int4/int8 keys collide (hashfunc.c:92-110)
    assert L.orc_hashint8(1) == L.orc_hashint4(1)
    assert L.orc_hashint8((1 << 32) + 1) == L.orc_hashint4(0)
    assert L.orc_hashint8(-1) == L.orc_hashint4(-1)
    # hash_uint32(k) == hash_any(&k, 4)
    for v in (0, 1, 0xDEADBEEF, 0x7FFFFFFF):
        assert L.orc_hash_uint32(v) == L.orc_hash_any(int(v).to_bytes(4, "little"), 4)
    # abs(INT_MIN) % 4096 == 0 (shardmap.c:1156)
    assert L.orc_shard_index(0x80000000) == 0
    assert L.orc_shard_index(0xFFFFFFFF) == 1


def numeric_avg(total, count):
    """int8_avg's numeric division as psql prints it (16 fractional digits)."""
    getcontext().prec = 40
    return str((Decimal(total) / Decimal(count)).quantize(Decimal("0.0000000000000001")))


def test_golden_xc_groupby():
    t = GOLD["xc_groupby_tab1"]
    rel = O.Rel([O.GX_INT4, O.GX_INT4], [np.array(t["columns"]["val"], np.int32), np.array(t["columns"]["val2"], np.int32)])
    plan = O.make_plan(group_cols=[(0, 1)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_I4, [(g.GX_OP_COL, 0, 0)]),
                             (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 0, 0)])])
    r = O.exec_agg(rel, plan)
    got = {int(k): (int(c), int(sm), float(av)) for k, c, sm, av in
           zip(r.keys[:, 0], r.aggs[:, 0].view(np.int64), r.aggs[:, 1].view(np.int64), r.aggs[:, 2])}
    assert len(got) == len(t["expected"])
    for e in t["expected"]:
        c, sm, av = got[e["val2"]]
        assert (c, sm) == (e["count"], e["sum"])
        assert numeric_avg(sm, c) == e["avg_numeric"]                 # avg(int4): the finalfn divides the same two integers
        assert repr(float(sm) / c).removesuffix(".0") == e["sum_f8_div_count"]
        assert av == float(sm) / c or abs(av - float(sm) / c) <= 1e-15 * abs(av)
    n = GOLD["xc_groupby_nested_sum"]["expected"]
    by_parity = {}
    for k, (c, sm, av) in got.items():
        by_parity[k % 2] = by_parity.get(k % 2, 0) + sm
    assert sorted(by_parity.values()) == n


def test_golden_join_groupby():
    t1, t2 = GOLD["xc_groupby_tab1"]["columns"], GOLD["xc_groupby_tab2"]["columns"]
    r1 = O.Rel([O.GX_INT4, O.GX_INT4], [np.array(t1["val"], np.int32), np.array(t1["val2"], np.int32)])
    r2 = O.Rel([O.GX_INT4, O.GX_INT4], [np.array(t2["val"], np.int32), np.array(t2["val2"], np.int32)])
    # tab1 JOIN tab2 ON tab1.val2 = tab2.val2, N:M
    join = O.make_join(1, payload_cols=[0], inner_unique=0)
    plan = O.make_plan(outer_key_col=1, group_cols=[(0, 1)], aggs=[(g.GX_AGG_COUNT_STAR, [])])
    r = O.exec_agg(r1, plan, r2, join)
    got = dict(zip(r.keys[:, 0].tolist(), r.aggs[:, 0].view(np.int64).tolist()))
    want = {e["val2"]: e["count"] for e in GOLD["xc_groupby_join_inner_part"]["expected_matched_groups"]}
    assert got == want
    cols = O.exec_join(r1, 1, r2, join, [0, 1])
    prod = {}
    for v1, k, v2 in zip(*[c.tolist() for c in cols]):
        prod[k] = prod.get(k, 0) + v1 * v2
    assert prod == {e["val2"]: e["sum_val1_times_val2"] for e in GOLD["xc_groupby_join_inner_part"]["expected_matched_groups"]}


def test_golden_opentenbase_c_aggregation():
    lo, hi = GOLD["opentenbase_c_aggregation_float8"]["series"]
    v = np.arange(lo, hi + 1, dtype=np.float64)
    plan = O.make_plan(aggs=[(g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 0, 0)]), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 0, 0)]),
                             (g.GX_AGG_MAX_F8, [(g.GX_OP_COL, 0, 0)]), (g.GX_AGG_MIN_F8, [(g.GX_OP_COL, 0, 0)])])
    r = O.exec_agg(O.Rel([O.GX_FLOAT8], [v]), plan)
    e = GOLD["opentenbase_c_aggregation_float8"]["expected"]
    # float8out prints the shortest round-trip representation ("2.5", "65", ...)
    assert [repr(float(x)).removesuffix(".0") for x in r.aggs[0]] == [e["avg"], e["sum"], e["max"], e["min"]]
    i = np.arange(lo, hi + 1, dtype=np.int32)
    plan = O.make_plan(aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_I4, [(g.GX_OP_COL, 0, 0)])])
    r = O.exec_agg(O.Rel([O.GX_INT4], [i]), plan)
    c, sm = int(r.aggs[0, 0].view(np.int64)), int(r.aggs[0, 1].view(np.int64))
    ei = GOLD["opentenbase_c_aggregation_int4"]["expected"]
    assert sm == ei["sum"] and numeric_avg(sm, c) == ei["avg_numeric"]


def test_heap_pages_round_trip():
    """heap_form_tuple + PageAddItem -> heapgetpage + slot_deform_tuple is the identity,
    page and tuple headers have the OpenTenBase layout (t_hoff 48, 44-byte page header)."""
    rng = np.random.default_rng(1)
    n = 3000
    types = [O.GX_INT8, O.GX_INT4, O.ORC_BPCHAR1, O.GX_FLOAT8, O.GX_CHAR, O.GX_DATE]
    cols = [rng.integers(-2**62, 2**62, n), rng.integers(-2**31, 2**31, n).astype(np.int32),
            rng.integers(65, 90, n).astype(np.int8), rng.normal(size=n), rng.integers(32, 127, n).astype(np.int8),
            rng.integers(-3000, 0, n).astype(np.int32)]
    nulls = [None, (rng.random(n) < 0.2).astype(np.uint8), (rng.random(n) < 0.2).astype(np.uint8), None, None,
             (rng.random(n) < 0.2).astype(np.uint8)]
    rel = O.Rel(types, cols, nulls)
    assert rel.ntuples == n and rel.npages > 10
    pages = rel.pages().reshape(-1, 8192)
    pd_lower = pages[:, 16:20].copy().view(np.uint32)[:, 0].astype(int)      # LocationIndex is uint32 in OpenTenBase-C
    pd_upper = pages[:, 20:24].copy().view(np.uint32)[:, 0].astype(int)
    assert (pd_lower >= 44).all() and (pd_upper <= 8192).all() and (pd_lower <= pd_upper).all()
    assert ((pd_lower - 44) % 4 == 0).all()
    lp0 = int.from_bytes(pages[0, 44:48].tobytes(), "little")
    off, flags, ln = lp0 & 0x7FFF, (lp0 >> 15) & 3, lp0 >> 17
    assert flags == 1 and off % 8 == 0 and off + ln <= 8192
    assert pages[0, off + 46] in (48, 56)                       # t_hoff: 48 without NULLs (htup_details.h)
    got, gn = rel.scan(list(range(len(types))))
    for c in range(len(types)):
        want_n = nulls[c] if nulls[c] is not None else np.zeros(n, np.uint8)
        np.testing.assert_array_equal(gn[c], want_n)
        keep = want_n == 0
        if got[c].dtype == np.float64:
            np.testing.assert_array_equal(got[c].view(np.int64)[keep], cols[c].view(np.int64)[keep])
        else:
            np.testing.assert_array_equal(got[c][keep], np.asarray(cols[c])[keep])
    assert rel.delete(0, 1) == 0 and rel.ntuples == n - 1
    got2, _ = rel.scan([0])
    np.testing.assert_array_equal(got2[0], np.asarray(cols[0])[1:])


def test_executor_against_numpy():
    """Independent restatement in numpy of join + group-by on random data (incl. NULLs)."""
    rng = np.random.default_rng(2)
    n, m = 20000, 300
    key = rng.integers(0, 400, n).astype(np.int64); knull = (rng.random(n) < 0.05).astype(np.uint8)
    val = rng.normal(10, 3, n); vnull = (rng.random(n) < 0.1).astype(np.uint8)
    bkey = np.arange(m, dtype=np.int64); bgrp = (bkey % 17).astype(np.int32)
    outer = O.Rel([O.GX_INT8, O.GX_FLOAT8], [key, val], [knull, vnull])
    inner = O.Rel([O.GX_INT8, O.GX_INT4], [bkey, bgrp])
    plan = O.make_plan(outer_key_col=0, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_COUNT, [(g.GX_OP_COL, 1, 0)]),
                             (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)])])
    r = O.exec_agg(outer, plan, inner, O.make_join(0, payload_cols=[1], inner_unique=1)).sorted()
    ok = (knull == 0) & (key < m)
    grp = (key % 17)[ok]
    for i, gk in enumerate(r.keys[:, 0]):
        sel = grp == gk
        vv, vn = val[ok][sel], vnull[ok][sel]
        assert int(r.aggs[i, 0].view(np.int64)) == int(sel.sum())
        assert int(r.aggs[i, 1].view(np.int64)) == int((vn == 0).sum())
        # float8pl in scan order
        acc = None
        for x in vv[vn == 0]:
            acc = x if acc is None else acc + x
        assert r.aggs[i, 2] == acc
        np.testing.assert_allclose(r.aggs[i, 3], vv[vn == 0].mean(), rtol=1e-12)
    assert r.ngroups == 17


def test_partial_final_combine_equals_single_node():
    """Partial HashAggregate per datanode + Finalize == one-node aggregate (xc_groupby.out:193-205)."""
    sf, nord, nn = 1, 6000, 4
    plan = O.make_plan(outer_key_col=g.L_ORDERKEY, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, g.L_EXTENDEDPRICE, 0)]),
                             (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, g.L_QUANTITY, 0)])])
    join = O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1)
    whole = O.exec_agg(O.Rel(g.SCHEMAS[g.T_LINEITEM], O.gen_lineitem(sf, 0, nord)), plan,
                       O.Rel(g.SCHEMAS[g.T_ORDERS], O.gen_orders(sf, 0, nord)), join).sorted()
    raws, rows = [], 0
    for node in range(nn):
        o, l = O.gen_orders(sf, 0, nord, node, nn), O.gen_lineitem(sf, 0, nord, node, nn)
        rows += len(l[0])
        assert (O.route_nodes(o[0], O.GX_INT8, nn) == node).all()
        raws.append(O.exec_agg(O.Rel(g.SCHEMAS[g.T_LINEITEM], l), plan, O.Rel(g.SCHEMAS[g.T_ORDERS], o), join, keep_raw=True)[1])
    comb = O.combine(plan, raws).sorted()
    assert rows == int(whole.aggs[:, 0].view(np.int64).sum())
    np.testing.assert_array_equal(comb.keys, whole.keys)
    np.testing.assert_array_equal(comb.aggs[:, 0].view(np.int64), whole.aggs[:, 0].view(np.int64))
    np.testing.assert_allclose(comb.aggs[:, 1:], whole.aggs[:, 1:], rtol=1e-9)


def test_generator_shape():
    sf, n = 1, 30000
    o, l = O.gen_orders(sf, 0, n), O.gen_lineitem(sf, 0, n)
    assert (np.diff(o[0]) > 0).all() and o[0][0] == 1 and o[0][8] == 33          # sparse TPC-H order keys
    assert (o[1] % 3 != 0).all() and o[1].min() >= 1 and o[1].max() <= 150000
    assert o[2].min() >= -2922 and o[2].max() <= -517
    assert 3.9 < len(l[0]) / n < 4.1
    assert set(np.unique(l[6]).tolist()) <= {ord("R"), ord("A"), ord("N")}
    assert set(np.unique(l[7]).tolist()) <= {ord("O"), ord("F")}
    cents = np.round(l[2] * 100).astype(np.int64)
    np.testing.assert_array_equal(cents / 100.0, l[2])                           # exact 2-decimal values
    assert np.isin(l[0], o[0]).all()


def test_bloom_filter_restatement():
    L = O.lib()
    b = L.orc_bloom_create(100000)
    assert L.orc_bloom_log_num_buckets(b) >= 10
    keys = [L.orc_hashint8new(i) for i in range(0, 20000, 2)]
    for k in keys:
        L.orc_bloom_insert(b, k)
    assert all(L.orc_bloom_find(b, k) for k in keys)
    fp = sum(L.orc_bloom_find(b, L.orc_hashint8new(i)) for i in range(1, 20001, 2))
    assert fp < 0.02 * 10000
    L.orc_bloom_free(b)
    assert not L.orc_bloom_create(10**9)             # logNumBuckets > 20: "give up using bloom filter"


def test_oracle_matches_committed_reference_object_vectors():
    """tests/golden/reference_object_vectors.json holds values computed by the reference's own
    object code (generator: tests/golden/make_reference_object_vectors.py).  Unlike
    tests/test_oracle_vs_ref.py this needs neither /root/reference nor oracle/_ref."""
    import ctypes as C
    import json
    import os
    import struct
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_object_vectors.json")))
    L = O.lib()
    for v, h in g["hashint4"]:
        assert L.orc_hashint4(v) == h
    for v, h in g["hashint4new"]:
        assert L.orc_hashint4new(v) == h
    for v, h in g["hashint8"]:
        assert L.orc_hashint8(v) == h
    for v, h in g["hashint8new"]:
        assert L.orc_hashint8new(v) == h
    for v, h in g["murmurhash32"]:
        assert L.orc_murmurhash32(v) == h
    for v, h in g["hashchar"]:
        assert L.orc_hashchar(v) == h
    for bits, h in g["hashfloat8_bits"]:
        assert L.orc_hashfloat8(struct.unpack("<d", struct.pack("<q", bits))[0]) == h
    for a, b, h in g["hash_combine"]:
        assert L.orc_hash_combine(a, b) == h
    for hx, h_old, h_new in g["hash_any"]:
        b = bytes.fromhex(hx)
        assert L.orc_hash_any(b, len(b)) == h_old and L.orc_hash_any_new(b, len(b)) == h_new
    for v, h in g["evaluate_hashkey_int4"]:
        assert L.orc_evaluate_hashkey((C.c_int * 1)(O.GX_INT4), None, (C.c_int64 * 1)(v), 1) == h
    for v, h in g["evaluate_hashkey_int8"]:
        assert L.orc_evaluate_hashkey((C.c_int * 1)(O.GX_INT8), None, (C.c_int64 * 1)(v), 1) == h
    for a, b, h in g["evaluate_hashkey_int8_int4"]:
        assert L.orc_evaluate_hashkey((C.c_int * 2)(O.GX_INT8, O.GX_INT4), None, (C.c_int64 * 2)(a, b), 2) == h
    # float8_accum / float8pl: the oracle's aggregate states over the same inputs in the same order
    import opentenbase_b200 as gpu
    acc = g["float8_accum"]
    vals = np.array(acc["inputs_bits"], np.int64).view(np.float64)
    plan = O.make_plan(aggs=[(gpu.GX_AGG_AVG_F8, [(gpu.GX_OP_COL, 0, 0)]), (gpu.GX_AGG_SUM_F8, [(gpu.GX_OP_COL, 0, 0)])])
    r, _ = O.exec_agg(O.Rel([O.GX_FLOAT8], [vals]), plan, keep_raw=True)
    np.testing.assert_array_equal(r.states[0, 0].view(np.int64), np.array(acc["state_bits"], np.int64))
    assert np.float64(r.aggs[0, 1]).view(np.int64) == acc["float8pl_sum_bits"]


def test_short_tuples_read_added_columns_as_null():
    """heap_deform_tuple / slot_deform_tuple stop at the tuple's own attribute count
    (heaptuple.c:1424, :1555); the rest is NULL when the column has no missing value (:1497-1502)."""
    rel = O.Rel([O.GX_INT8, O.GX_INT4], [np.arange(10), np.arange(10, dtype=np.int32)])
    rel.add_column(O.GX_FLOAT8)
    rel.insert([np.arange(3), np.arange(3, dtype=np.int32), np.full(3, 2.5)])
    cols, nulls = rel.scan([0, 2])
    np.testing.assert_array_equal(nulls[1], [1] * 10 + [0] * 3)
    np.testing.assert_array_equal(cols[1][10:], [2.5] * 3)
    np.testing.assert_array_equal(cols[0], list(range(10)) + [0, 1, 2])


def test_q3_reference_agrees_with_a_numpy_restatement():
    """The whole-query Q3 reference used by bench.py and the multi-datanode tests, against plain numpy."""
    sf, nord, ncust, date, seg = 1, 20000, 150000, -1752, ord("B")
    want = O.q3_reference(sf, nord, ncust, date, seg)
    c, o, l = O.gen_customer(sf, 0, ncust), O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    good = np.isin(o[1], c[0][c[1] == seg]) & (o[2] < date)
    okeys = dict(zip(o[0][good].tolist(), zip(o[2][good].tolist(), o[3][good].tolist())))
    sums = {}
    for k, p, d, sd in zip(l[0].tolist(), l[2].tolist(), l[3].tolist(), l[5].tolist()):
        if sd > date and k in okeys:
            sums[k] = sums.get(k, 0.0) + p * (1.0 - d)
    assert want.ngroups == len(sums) > 50
    for (k, od, sp), s in zip(want.keys.tolist(), want.aggs[:, 0].tolist()):
        assert okeys[k] == (od, sp)
        assert abs(sums[k] - s) <= 1e-9 * abs(s)


def _fnpage_case_arrays(case):
    """fixture case -> (types, columns, null arrays, OrcFnPageId, full pages)"""
    types = case["types"]
    vals = np.array(case["values"], np.int64).reshape(-1, len(types))
    isn = np.array(case["isnull"], np.uint8).reshape(-1, len(types))
    cols = []
    for i, t in enumerate(types):
        c = vals[:, i]
        cols.append(c.view(np.float64) if t == O.GX_FLOAT8 else c.astype(O.NP_DTYPES[t]))
    nulls = [isn[:, i].copy() if isn[:, i].any() else None for i in range(len(types))]
    pid = O.OrcFnPageId(case["page_id"]["qid_ts"], case["page_id"]["qid_seq"], case["page_id"]["fid"], case["page_id"]["nodeid"],
                        case["page_id"]["workerid"], case["page_id"]["virtualid"], 0)
    pages = np.zeros((len(case["pages_used_hex"]), 8192), np.uint8)
    for p, h in enumerate(case["pages_used_hex"]):
        b = np.frombuffer(bytes.fromhex(h), np.uint8)
        pages[p, :len(b)] = b
    return types, cols, nulls, isn, pid, pages


def test_fnpage_oracle_matches_reference_vectors():
    """tests/golden/fnpage_vectors.json holds forward-node pages (the redistribute wire format) written by the
    reference's own heaptuple.o + fnbufpage.o (generator: tests/golden/make_fnpage_vectors.py).  The oracle's sender
    must produce the same bytes, and its receiver must read the reference's pages back into the inputs."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fnpage_vectors.json")))
    assert len(g["cases"]) >= 5
    for case in g["cases"]:
        types, cols, nulls, isn, pid, want = _fnpage_case_arrays(case)
        got = O.fnpage_pack(types, cols, nulls if any(x is not None for x in nulls) else None, pid, case["end_marker"])
        assert got.shape == want.shape, case["name"]
        np.testing.assert_array_equal(got, want, err_msg=case["name"])
        c2, n2 = O.fnpage_unpack(want, types)
        assert len(c2[0]) == len(cols[0]), case["name"]
        for i in range(len(types)):
            np.testing.assert_array_equal(n2[i], isn[:, i], err_msg=case["name"])
            keep = isn[:, i] == 0
            np.testing.assert_array_equal(np.asarray(c2[i])[keep].view(np.uint8), np.asarray(cols[i])[keep].view(np.uint8), err_msg=case["name"])
