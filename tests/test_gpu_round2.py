"""Round-2 parity cases (all through the C ABI, all against the oracle):
float8 overflow -> ERROR, tuples shorter than the descriptor, the Q3 fragment chain as the
provider issues it, full-SF1 / SF10-slice comparisons of every BASELINE query shape."""
import numpy as np
import pytest

import opentenbase_b200 as g
from opentenbase_b200 import plans as P
import oracle as O
from helpers import assert_agg_equal, to_gpu_plan, lineitem_rel, orders_rel

pytestmark = pytest.mark.gpu

CCOLS = {"custkey": g.C_CUSTKEY, "mktsegment": g.C_MKTSEGMENT}
OCOLS = {"orderkey": g.O_ORDERKEY, "custkey": g.O_CUSTKEY, "orderdate": g.O_ORDERDATE, "shippriority": g.O_SHIPPRIORITY}
LCOLS = {"orderkey": g.L_ORDERKEY, "extendedprice": g.L_EXTENDEDPRICE, "discount": g.L_DISCOUNT, "shipdate": g.L_SHIPDATE}


def test_float8_overflow_is_an_error(gx):
    """float8pl's CHECKFLOATVAL (utils/adt/float.c:970-981): 1e308 + 1e308 from finite inputs is
    'value out of range: overflow' in the reference, GX_ERR_OVERFLOW here; an infinite INPUT is fine."""
    t = gx.table_from([g.GX_INT4, g.GX_FLOAT8], [np.array([1, 1, 2], np.int32), np.array([1e308, 1e308, 5.0])])
    plan = g.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=4)
    with pytest.raises(g.GxError) as ei:
        gx.hash_agg(t, plan).fetch()
    assert ei.value.status == g.GX_ERR_OVERFLOW
    plan = g.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)])], est_groups=4)
    with pytest.raises(g.GxError) as ei:
        gx.hash_agg(t, plan).fetch()
    assert ei.value.status == g.GX_ERR_OVERFLOW


def test_tuples_shorter_than_the_descriptor(gx):
    """ALTER TABLE ADD COLUMN without a rewrite: old tuples carry fewer attributes
    (t_infomask2 & HEAP_NATTS_MASK) and read the new ones as NULL — slot_deform_tuple's
    Min(HeapTupleHeaderGetNatts(tup), natts), heaptuple.c:1424,1497-1502,1555."""
    rng = np.random.default_rng(11)
    n = 3000
    rel = O.Rel([O.GX_INT8, O.GX_INT4, O.GX_FLOAT8],
                [rng.integers(0, 2**40, n), rng.integers(0, 100, n).astype(np.int32), rng.random(n)],
                [None, (rng.random(n) < 0.2).astype(np.uint8), None])
    rel.add_column(O.GX_FLOAT8)
    rel.add_column(O.GX_INT4)
    m = 500                                                    # rows written after the ALTER carry all five attributes
    rel.insert([rng.integers(0, 2**40, m), rng.integers(0, 100, m).astype(np.int32), rng.random(m), rng.random(m),
                rng.integers(0, 9, m).astype(np.int32)])
    attnums = [0, 2, 3, 4]
    want_cols, want_nulls = rel.scan(attnums)
    assert want_nulls[2][:n].all() and not want_nulls[2][n:].any()
    t = gx.table([g.GX_INT8, g.GX_FLOAT8, g.GX_FLOAT8, g.GX_INT4], n + m)
    t.append_heap_pages(rel.pages(), [8, 4, 8, 8, 4], [8, 4, 8, 8, 4], attnums)
    assert t.nrows == n + m
    for c in range(len(attnums)):
        got, gn = t.read(c, with_nulls=True)
        np.testing.assert_array_equal(gn, want_nulls[c])
        keep = want_nulls[c] == 0
        np.testing.assert_array_equal(got.view(np.int64 if got.dtype == np.float64 else got.dtype)[keep],
                                      want_cols[c].view(np.int64 if got.dtype == np.float64 else got.dtype)[keep])


def test_drop_column(gx):
    t = gx.table_from([g.GX_INT8, g.GX_INT4, g.GX_FLOAT8], [np.arange(5), np.arange(5, dtype=np.int32) * 2, np.arange(5) * 0.5])
    t.drop_column(1)
    assert t.types == [g.GX_INT8, g.GX_FLOAT8]
    np.testing.assert_array_equal(t.read(1), np.arange(5) * 0.5)


@pytest.mark.parametrize("nord", [30000, 400000])
def test_q3_fragment_chain_one_datanode(gx, nord):
    """customer JOIN orders JOIN lineitem through plans.q3_datanode (scan_filter -> redistribute ->
    build -> probe -> redistribute -> build -> fused probe+aggregate) on a single datanode, where the
    redistributes degenerate to local partition copies; the N > 1 path is checked by bench.py's
    q3_parity_vs_oracle under torchrun and by tests/test_multi_gpu.py."""
    sf, ncust = 1, 150_000
    gx.set_shardmap(1)
    ct = gx.table(g.SCHEMAS[g.T_CUSTOMER], ncust).generate(g.T_CUSTOMER, sf, 0, ncust)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], nord * 7).generate(g.T_LINEITEM, sf, 0, nord)
    stats = {}
    res = P.q3_datanode(gx, ct, ot, lt, CCOLS, OCOLS, LCOLS, stats)
    want = O.q3_reference(sf, nord, ncust, P.DATE_Q3, P.SEGMENT_Q3)
    assert want.ngroups > 50 and stats["joined"] < stats["orders_kept"]
    assert_agg_equal(res.plan, res.fetch(), want)
    for t in (ct, ot, lt):
        t.free()


def test_full_sf1_every_query_shape(gx):
    """Every BASELINE query shape at FULL SF1 (1.5 M orders, 6.0 M lineitem rows) against the oracle:
    config 1, config 2, config 3, Q1."""
    sf, nord = 1, 1_500_000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], len(l[0])).generate(g.T_LINEITEM, sf, 0, nord)
    assert lt.nrows == len(l[0])
    lrel, orel = lineitem_rel(l), orders_rel(o)
    for plan in (P.config1_plan(g.L_RETURNFLAG), P.config2_plan(g.L_SHIPDATE, g.L_EXTENDEDPRICE),
                 P.q1_plan(g.L_QUANTITY, g.L_EXTENDEDPRICE, g.L_DISCOUNT, g.L_TAX, g.L_SHIPDATE, g.L_RETURNFLAG, g.L_LINESTATUS)):
        want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)))
        assert_agg_equal(plan, gx.hash_agg(lt, plan).fetch(), want)
    plan = P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE)
    want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)), orel, O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1))
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    assert_agg_equal(plan, gx.hash_agg(lt, plan, ht).fetch(), want)
    for t in (ot, lt):
        t.free()


def test_sf10_slice_config2_config3_q3(gx):
    """A 3 M-order slice of the SF10 tables (15 M lineitem rows) for configs 2 and 3 and the Q3 chain."""
    sf, nord, ncust = 10, 3_000_000, 1_500_000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    ot = gx.table(g.SCHEMAS[g.T_ORDERS], nord).generate(g.T_ORDERS, sf, 0, nord)
    lt = gx.table(g.SCHEMAS[g.T_LINEITEM], len(l[0])).generate(g.T_LINEITEM, sf, 0, nord)
    ct = gx.table(g.SCHEMAS[g.T_CUSTOMER], ncust).generate(g.T_CUSTOMER, sf, 0, ncust)
    lrel, orel = lineitem_rel(l), orders_rel(o)
    plan = P.config2_plan(g.L_SHIPDATE, g.L_EXTENDEDPRICE)
    assert_agg_equal(plan, gx.hash_agg(lt, plan).fetch(), O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan))))
    plan = P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE)
    want = O.exec_agg(lrel, O.GxAggPlan.from_buffer_copy(bytes(plan)), orel, O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1))
    ht = gx.hash_build(ot, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    assert_agg_equal(plan, gx.hash_agg(lt, plan, ht).fetch(), want)
    gx.set_shardmap(1)
    res = P.q3_datanode(gx, ct, ot, lt, CCOLS, OCOLS, LCOLS)
    assert_agg_equal(res.plan, res.fetch(), O.q3_reference(sf, nord, ncust, P.DATE_Q3, P.SEGMENT_Q3))
    for t in (ot, lt, ct):
        t.free()


def _runagg_case(gx, okeys, odate, oprio, lkeys, lprice, ldisc, lship, aggs, group_cols, preds=()):
    """orders(key, date, prio) JOIN lineitem(key, price, disc, ship) with the group key containing the join key"""
    otypes = [g.GX_INT8, g.GX_DATE, g.GX_INT4]
    ltypes = [g.GX_INT8, g.GX_FLOAT8, g.GX_FLOAT8, g.GX_DATE]
    ocols, lcols = [okeys, odate, oprio], [lkeys, lprice, ldisc, lship]
    plan = O.make_plan(preds=list(preds), outer_key_col=0, group_cols=group_cols, aggs=aggs, est_groups=max(len(okeys), 16))
    want = O.exec_agg(O.Rel(ltypes, lcols), plan, O.Rel(otypes, ocols), O.make_join(0, payload_cols=[1, 2], inner_unique=1))
    ot, lt = gx.table_from(otypes, ocols), gx.table_from(ltypes, lcols)
    ht = gx.hash_build(ot, 0, [1, 2], unique=True)
    gx.profile(True)
    got = gx.hash_agg(lt, to_gpu_plan(plan), ht).fetch()
    used = gx.profile_get("runagg")[1] > 0
    fell_back = gx.profile_get("probe_records")[1] > 0
    gx.profile(False)
    assert_agg_equal(plan, got, want)
    for t in (ot, lt):
        t.free()
    return used, fell_back


def test_runagg_group_key_contains_join_key(gx):
    """gx_k_runagg: runs of equal keys are groups.  Long runs that span warp chunks and slabs, runs of one,
    rows that fail the qual, keys without a partner, several aggregates; then the layouts that must NOT
    take the kernel's answer: keys that descend somewhere (shuffled, or clustered but unordered)."""
    rng = np.random.default_rng(3)
    C, K, S, M = g.GX_OP_COL, g.GX_OP_CONST, g.GX_OP_SUB, g.GX_OP_MUL
    rev = [(C, 1, 0), (K, 0, 1.0), (C, 2, 0), (S, 0, 0), (M, 0, 0)]
    nkeys = 40000
    okeys = np.sort(rng.choice(np.arange(1, 10 * nkeys), nkeys, replace=False)).astype(np.int64)
    odate = rng.integers(-3000, -1000, nkeys).astype(np.int32); oprio = rng.integers(0, 3, nkeys).astype(np.int32)
    # run lengths: mostly 1..7, a few of thousands (span several 32-row slabs and whole warp chunks), keys missing from orders
    reps = rng.integers(1, 8, nkeys); reps[rng.choice(nkeys, 12, replace=False)] = rng.integers(2000, 9000, 12)
    lk = np.repeat(okeys + (rng.random(nkeys) < 0.1), reps).astype(np.int64)      # 10 % of the keys shifted off their partner
    lk.sort()
    n = len(lk)
    lprice, ldisc = np.round(rng.random(n) * 1e5, 2), np.round(rng.random(n) * 0.1, 2)
    lship = rng.integers(-3000, -1000, n).astype(np.int32)
    aggs3 = [(g.GX_AGG_SUM_F8, rev), (g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_AVG_F8, [(C, 1, 0)]), (g.GX_AGG_SUM_F8, [(C, 2, 0)])]
    full_key = [(0, 0), (1, 0), (1, 1)]
    used, fb = _runagg_case(gx, okeys, odate, oprio, lk, lprice, ldisc, lship, [(g.GX_AGG_SUM_F8, rev)], full_key, preds=[(3, g.GX_GT, -1752)])
    assert used and not fb
    used, fb = _runagg_case(gx, okeys, odate, oprio, lk, lprice, ldisc, lship, aggs3, full_key)
    assert used and not fb
    used, fb = _runagg_case(gx, okeys, odate, oprio, lk, lprice, ldisc, lship, aggs3, [(0, 0)], preds=[(3, g.GX_GT, -2500), (1, g.GX_LT, 9e4, True)])
    assert used and not fb
    # clustered but not ordered: the runs are intact, their order is not -> the flag must fire and the general path answers
    order = np.argsort(rng.permutation(len(okeys))[np.searchsorted(okeys, lk - (~np.isin(lk, okeys)))], kind="stable")
    used, fb = _runagg_case(gx, okeys, odate, oprio, lk[order], lprice[order], ldisc[order], lship[order], aggs3, full_key)
    assert used and fb
    perm = rng.permutation(n)
    used, fb = _runagg_case(gx, okeys, odate, oprio, lk[perm], lprice[perm], ldisc[perm], lship[perm], aggs3, full_key)
    assert used and fb
    # tiny inputs: fewer rows than one slab, a single row
    for m in (1, 5, 33):
        used, fb = _runagg_case(gx, okeys, odate, oprio, lk[:m], lprice[:m], ldisc[:m], lship[:m], aggs3, full_key)
        assert used and not fb


@pytest.mark.parametrize("ndistinct", [1, 3, 4, 5, 8, 9, 40])
def test_few_groups_register_kernels(gx, ndistinct):
    """gx_k_fewgroups / gx_k_count_char (<= 4 resp. 8 groups in registers) and their fall-back when the data
    holds more groups than the planner promised; 1-byte keys with the sign bit set; int4 keys; row counts that
    are not a multiple of the vector width."""
    rng = np.random.default_rng(ndistinct)
    n = 100003
    vals = rng.choice(np.arange(-128, 128), ndistinct, replace=False).astype(np.int8)
    flag = rng.choice(vals, n)
    k4 = rng.choice(rng.integers(-2**31, 2**31, ndistinct), n).astype(np.int32)
    x, y = np.round(rng.random(n) * 1e4, 2), np.round(rng.random(n), 2)
    types = [g.GX_CHAR, g.GX_INT4, g.GX_FLOAT8, g.GX_FLOAT8]
    cols = [flag, k4, x, y]
    rel = O.Rel(types, cols)
    t = gx.table_from(types, cols)
    C, K, S, M = g.GX_OP_COL, g.GX_OP_CONST, g.GX_OP_SUB, g.GX_OP_MUL
    expr = [(C, 2, 0), (K, 0, 1.0), (C, 3, 0), (S, 0, 0), (M, 0, 0)]
    plans = [
        O.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_COUNT_STAR, [])], est_groups=3),                       # config 1 shape
        O.make_plan(group_cols=[(0, 0)], aggs=[(g.GX_AGG_SUM_F8, expr), (g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_AVG_F8, [(C, 3, 0)])],
                    preds=[(2, g.GX_LT, 9000.0, True)], est_groups=4),
        O.make_plan(group_cols=[(0, 1)], aggs=[(g.GX_AGG_SUM_F8, [(C, 2, 0)]), (g.GX_AGG_COUNT_STAR, [])], est_groups=2),
        O.make_plan(group_cols=[(0, 0), (0, 1)], aggs=[(g.GX_AGG_AVG_F8, [(C, 2, 0)])], est_groups=6),
    ]
    for plan in plans:
        assert_agg_equal(plan, gx.hash_agg(t, to_gpu_plan(plan)).fetch(), O.exec_agg(rel, plan))
    # short and unaligned inputs through the byte-vector kernel
    for m in (0, 1, 15, 16, 17, 33):
        sub = gx.table_from(types, [c[:m] for c in cols]) if m else gx.table(types, 1)
        assert_agg_equal(plans[0], gx.hash_agg(sub, to_gpu_plan(plans[0])).fetch(), O.exec_agg(O.Rel(types, [c[:m] for c in cols]), plans[0]))
        sub.free()
    t.free()


def test_permuted_tables_give_the_same_answer(gx, ):
    """gx_table_permute is a bijection, and the join + aggregate does not depend on the row order of either side
    (key-ordered build and run-folding probe are optimisations that verify their precondition)."""
    sf, nord = 1, 50000
    o, l = O.gen_orders(sf, 0, nord), O.gen_lineitem(sf, 0, nord)
    ot = gx.table_from(g.SCHEMAS[g.T_ORDERS], o); lt = gx.table_from(g.SCHEMAS[g.T_LINEITEM], l)
    ou, lu = ot.permuted(3), lt.permuted(5)
    assert sorted(ou.read(0).tolist()) == sorted(o[0].tolist())
    assert not np.array_equal(lu.read(0), l[0])
    plan = P.config3_plan(g.L_ORDERKEY, g.L_EXTENDEDPRICE)
    want = O.exec_agg(lineitem_rel(l), O.GxAggPlan.from_buffer_copy(bytes(plan)), orders_rel(o),
                      O.make_join(g.O_ORDERKEY, payload_cols=[g.O_ORDERDATE], inner_unique=1))
    ht = gx.hash_build(ou, g.O_ORDERKEY, [g.O_ORDERDATE], unique=True)
    assert_agg_equal(plan, gx.hash_agg(lu, plan, ht).fetch(), want)
    for t in (ot, lt, ou, lu):
        t.free()


def test_bloom_filter_bit_identical_with_the_reference_geometry(gx):
    """gx_bloom_build / gx_bloom_test against the oracle's BlockBloomFilter (itself pinned to the reference's
    bloomfilter.o): same number of buckets, the SAME directory words bit for bit, the same Find answers — for int8
    and int4 keys, NULL keys, a build-side qual, and the size at which the reference gives the filter up."""
    import ctypes as C
    L = O.lib()
    rng = np.random.default_rng(8)
    n = 60000
    keys = rng.choice(np.arange(1, 10**7), n, replace=False).astype(np.int64)
    nulls = (rng.random(n) < 0.05).astype(np.uint8)
    flag = rng.integers(0, 2, n).astype(np.int32)
    t = gx.table_from([g.GX_INT8, g.GX_INT4], [keys, flag], [nulls, None])
    for preds, keep in (((), nulls == 0), (((1, g.GX_EQ, 1),), (nulls == 0) & (flag == 1))):
        b = gx.bloom_build(t, 0, preds)
        ob = L.orc_bloom_create(n)
        assert g.lib().gx_bloom_log_num_buckets(b) == L.orc_bloom_log_num_buckets(ob)
        for k in keys[keep].tolist():
            L.orc_bloom_insert(ob, L.orc_hashint8new(k))
        nw = C.c_int64()
        L.orc_bloom_words.restype = C.POINTER(C.c_uint32)
        want = np.ctypeslib.as_array(L.orc_bloom_words(ob, C.byref(nw)), (8 << L.orc_bloom_log_num_buckets(ob),)).copy()
        np.testing.assert_array_equal(gx.bloom_words(b), want)
        probe = np.concatenate([keys[:2000], rng.integers(10**7, 2 * 10**7, 20000)]).astype(np.int64)
        pt = gx.table_from([g.GX_INT8], [probe])
        got = gx.bloom_test(b, pt, 0)
        exp = np.array([L.orc_bloom_find(ob, L.orc_hashint8new(k)) for k in probe.tolist()], np.uint8)
        np.testing.assert_array_equal(got, exp)
        assert got[:2000][keep[:2000]].all()                       # no false negatives
        assert got[2000:].mean() < 0.12                            # false positives near the 5 % the sizing aims at
        # an int4 probe column of the same values hashes alike (hashint4new widens to int64)
        small = rng.integers(1, 10**7, 5000).astype(np.int32)
        p4 = gx.table_from([g.GX_INT4], [small])
        exp4 = np.array([L.orc_bloom_find(ob, L.orc_hashint4new(int(k))) for k in small.tolist()], np.uint8)
        np.testing.assert_array_equal(gx.bloom_test(b, p4, 0), exp4)
        g.lib().gx_bloom_free(b); L.orc_bloom_free(ob)
        pt.free(); p4.free()
    t.free()
    # "give up using bloom filter": more than 2^20 buckets wanted (bloomfilter.c:68)
    big = gx.table([g.GX_INT8], 60_000_000).generate(g.T_ORDERS, 40, 0, 60_000_000, colmap=[g.O_ORDERKEY])
    assert gx.bloom_build(big, 0) is None and not L.orc_bloom_create(60_000_000)
    big.free()


def test_left_semi_anti_joins(gx):
    """gx_hash_probe_ex against the oracle's ExecHashJoinImpl restatement (nodeHashjoin.c:466-689): LEFT emits the
    unmatched outer rows with a NULL inner side, SEMI each outer row once, ANTI the rows without a partner — with
    NULL outer keys (never match), duplicate build keys (SEMI/ANTI), outer quals, and unmatched keys."""
    rng = np.random.default_rng(12)
    ni, no = 30000, 120000
    uk = rng.choice(np.arange(1, 200000), ni, replace=False).astype(np.int64)
    dk = np.concatenate([uk[:ni // 2], uk[:ni // 2]])                       # every key twice
    ipay = rng.integers(0, 1000, ni).astype(np.int32)
    okey = rng.integers(1, 200000, no).astype(np.int64)
    onull = (rng.random(no) < 0.03).astype(np.uint8)
    oval = np.arange(no, dtype=np.int32)
    types = [g.GX_INT8, g.GX_INT4]
    ot = gx.table_from(types, [okey, oval], [onull, None])
    orel = O.Rel(types, [okey, oval], [onull, None])
    preds = [(1, g.GX_GE, 1000)]
    NULL = np.iinfo(np.int64).min

    def gpu_rows(t, with_inner):
        cols = [t.read(c, with_nulls=True) for c in range(len(t.types))]
        n = t.nrows
        out = np.full((n, len(cols)), NULL, np.int64)
        for c, (v, nl) in enumerate(cols):
            out[:, c] = np.where(nl == 0, v.astype(np.int64), NULL)
        return out[np.lexsort(out.T[::-1])]

    def want_rows(cols):
        w = np.stack(cols, 1) if len(cols[0]) else np.zeros((0, len(cols)), np.int64)
        return w[np.lexsort(w.T[::-1])]

    for keys, unique, jts in ((uk, True, (g.GX_JOIN_INNER, g.GX_JOIN_LEFT, g.GX_JOIN_SEMI, g.GX_JOIN_ANTI)),
                              (dk, False, (g.GX_JOIN_INNER, g.GX_JOIN_SEMI, g.GX_JOIN_ANTI))):
        it = gx.table_from(types, [keys, ipay])
        irel = O.Rel(types, [keys, ipay])
        ht = gx.hash_build(it, 0, [1], unique=unique)
        for jt in jts:
            got = gx.hash_probe(ot, 0, ht, [0, 1], preds=preds, join_type=jt)
            w = O.exec_join(orel, 0, irel, O.make_join(0, payload_cols=[1], inner_unique=int(unique), jointype=jt), [0, 1], outer_preds=preds)
            if jt in (g.GX_JOIN_SEMI, g.GX_JOIN_ANTI):
                w = w[:2]                                                   # the inner side is not part of the target list
            np.testing.assert_array_equal(gpu_rows(got, jt), want_rows(w), err_msg=f"join type {jt}, unique={unique}")
            got.free()
        ht.free(); it.free()
    # LEFT over a non-unique build side is declined, RIGHT/FULL are declined
    it = gx.table_from(types, [dk, ipay]); ht = gx.hash_build(it, 0, [1], unique=False)
    for jt in (g.GX_JOIN_LEFT, 2, 3):
        with pytest.raises(g.GxError) as ei:
            gx.hash_probe(ot, 0, ht, [0, 1], join_type=jt)
        assert ei.value.status == g.GX_ERR_ARG
    ot.free(); it.free()


def test_not_null_columns_are_staged_without_null_arrays(gx):
    """pg_attribute.attnotnull reaches the loader (gx_heap_desc.att_notnull): such a column gets no NULL array, so the
    no-NULL fast paths stay open for heap-loaded tables; a NULL arriving in it is reported, not silently kept."""
    rng = np.random.default_rng(4)
    n = 4000
    types = [O.GX_INT8, O.GX_FLOAT8]
    cols = [rng.integers(0, 2**40, n), rng.random(n)]
    rel = O.Rel(types, cols)
    t = gx.table([g.GX_INT8, g.GX_FLOAT8], n)
    t.append_heap_pages(rel.pages(), [8, 8], [8, 8], [0, 1], notnull=[True, True])
    k, kn = t.read(0, with_nulls=True)
    np.testing.assert_array_equal(k, cols[0]); assert not kn.any()
    # the same relation with a NULL in the "NOT NULL" column
    rel2 = O.Rel(types, cols, [None, (np.arange(n) == 17).astype(np.uint8)])
    t2 = gx.table([g.GX_INT8, g.GX_FLOAT8], n)
    with pytest.raises(g.GxError) as ei:
        t2.append_heap_pages(rel2.pages(), [8, 8], [8, 8], [0, 1], notnull=[True, True])
    assert ei.value.status == g.GX_ERR_STATE
    t3 = gx.table([g.GX_INT8, g.GX_FLOAT8], n)
    t3.append_heap_pages(rel2.pages(), [8, 8], [8, 8], [0, 1], notnull=[True, False])
    assert t3.read(1, with_nulls=True)[1].sum() == 1
    for x in (t, t2, t3):
        x.free()
