"""CPU oracle — TEST INFRASTRUCTURE, not part of the product.

ctypes binding of ``oracle/liboracle.so`` (a plain-C restatement of the
reference executor hot path, see ``otb_oracle.h``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this package; ``opentenbase_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# mirror of include/gpuexec.h enums (the shared plan vocabulary)
GX_INT4, GX_INT8, GX_FLOAT8, GX_DATE, GX_CHAR, ORC_BPCHAR1 = 1, 2, 3, 4, 5, 6
NP_DTYPES = {GX_INT4: np.int32, GX_INT8: np.int64, GX_FLOAT8: np.float64,
             GX_DATE: np.int32, GX_CHAR: np.int8, ORC_BPCHAR1: np.int8}


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (building the checker is not using it)."""
    # make is incremental and tracks the shared headers: always let it decide
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


# ---- POD structs shared with include/gpuexec.h ---------------------------
class GxPred(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("ival", C.c_int64), ("fval", C.c_double)]


class GxExprOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32), ("k", C.c_double)]


class GxExpr(C.Structure):
    _fields_ = [("nops", C.c_int32), ("_pad", C.c_int32), ("ops", GxExprOp * 12)]


class GxAgg(C.Structure):
    _fields_ = [("fn", C.c_int32), ("_pad", C.c_int32), ("arg", GxExpr)]


class GxColRef(C.Structure):
    _fields_ = [("side", C.c_int32), ("col", C.c_int32)]


class GxAggPlan(C.Structure):
    _fields_ = [("n_preds", C.c_int32), ("outer_key_col", C.c_int32),
                ("preds", GxPred * 4),
                ("n_group_cols", C.c_int32), ("n_aggs", C.c_int32),
                ("group_cols", GxColRef * 4),
                ("aggs", GxAgg * 8),
                ("est_groups", C.c_int64),
                ("strategy", C.c_int32), ("_pad", C.c_int32)]


class OrcJoinSpec(C.Structure):
    _fields_ = [("inner_key_col", C.c_int32), ("n_inner_preds", C.c_int32),
                ("inner_preds", GxPred * 4),
                ("n_payload", C.c_int32), ("payload_cols", C.c_int32 * 2),
                ("inner_unique", C.c_int32), ("jointype", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [("n_group_cols", C.c_int32), ("n_aggs", C.c_int32),
                ("ngroups", C.c_int64),
                ("keys", C.POINTER(C.c_int64)), ("aggs", C.POINTER(C.c_double)),
                ("nulls", C.POINTER(C.c_uint8)), ("states", C.POINTER(C.c_double))]


class OrcFnPageId(C.Structure):
    """what FnPageInit stamps on every page of a fragment's stream (forward/fnbufpage.h:54-65)"""
    _fields_ = [("qid_timestamp_nodeid", C.c_int64), ("qid_sequence", C.c_int64), ("fid", C.c_uint16), ("nodeid", C.c_uint16),
                ("workerid", C.c_uint16), ("virtualid", C.c_uint8), ("pad", C.c_uint8)]


def _declare(L: C.CDLL) -> None:
    u32, i32, i64, dbl, vp = C.c_uint32, C.c_int32, C.c_int64, C.c_double, C.c_void_p
    for name, res, args in [
        ("orc_hash_any", u32, [C.c_char_p, C.c_int]),
        ("orc_hash_uint32", u32, [u32]),
        ("orc_hashint4", u32, [i32]), ("orc_hashint8", u32, [i64]),
        ("orc_hashchar", u32, [C.c_int8]), ("orc_hashfloat8", u32, [dbl]),
        ("orc_crc32c", u32, [u32, vp, C.c_size_t]),
        ("orc_hash_any_new", u32, [C.c_char_p, C.c_int]),
        ("orc_hashint4new", u32, [i32]), ("orc_hashint8new", u32, [i64]),
        ("orc_hashcharnew", u32, [C.c_int8]), ("orc_hashfloat8new", u32, [dbl]),
        ("orc_murmurhash32", u32, [u32]), ("orc_hash_combine", u32, [u32, u32]),
        ("orc_hash_datum", u32, [C.c_int, i64]), ("orc_hash_datum_new", u32, [C.c_int, i64]),
        ("orc_evaluate_hashkey", u32, [vp, vp, vp, C.c_int]),
        ("orc_shard_index", i32, [u32]),
        ("orc_default_shardmap", None, [vp, C.c_int]),
        ("orc_route_node", i32, [vp, C.c_int, i64, C.c_int]),
        ("orc_gen_orders", i64, [C.c_int, i64, i64, C.c_int, C.c_int, vp, vp, vp, vp]),
        ("orc_gen_lineitem_count", i64, [C.c_int, i64, i64, C.c_int, C.c_int]),
        ("orc_gen_lineitem", i64, [C.c_int, i64, i64, C.c_int, C.c_int] + [vp] * 8),
        ("orc_gen_customer", i64, [C.c_int, i64, i64, C.c_int, C.c_int, vp, vp]),
        ("orc_rel_create", vp, [C.c_int, vp]), ("orc_rel_free", None, [vp]),
        ("orc_rel_add_column", C.c_int, [vp, i32]),
        ("orc_rel_insert_columns", C.c_int, [vp, vp, vp, i64]),
        ("orc_rel_ntuples", i64, [vp]), ("orc_rel_npages", i64, [vp]),
        ("orc_rel_page", vp, [vp, i64]), ("orc_rel_copy_pages", None, [vp, vp]),
        ("orc_rel_delete_tuple", C.c_int, [vp, i64, C.c_int]),
        ("orc_rel_scan_columns", i64, [vp, C.c_int, vp, vp, vp]),
        ("orc_result_free", None, [C.POINTER(OrcResult)]),
        ("orc_exec_agg", C.c_int, [vp, vp, C.POINTER(OrcJoinSpec), C.POINTER(GxAggPlan), C.POINTER(OrcResult)]),
        ("orc_exec_join", i64, [vp, C.c_int, C.c_int, vp, vp, C.POINTER(OrcJoinSpec), C.c_int, vp, vp]),
        ("orc_combine_results", C.c_int, [C.POINTER(GxAggPlan), C.POINTER(OrcResult), C.c_int, C.POINTER(OrcResult)]),
        ("orc_last_exec_seconds", dbl, []),
        ("orc_last_hash_stats", None, [vp, vp, vp, vp]),
        ("orc_bloom_create", vp, [i64]), ("orc_bloom_insert", None, [vp, u32]),
        ("orc_bloom_find", C.c_int, [vp, u32]), ("orc_bloom_log_num_buckets", C.c_int, [vp]),
        ("orc_bloom_words", vp, [vp, vp]), ("orc_bloom_free", None, [vp]),
        ("orc_fnpage_pack", i64, [C.c_int, vp, vp, vp, i64, C.POINTER(OrcFnPageId), C.c_int, vp, i64]),
        ("orc_fnpage_unpack", i64, [vp, i64, C.c_int, vp, vp, vp, i64]),
    ]:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args


# ---- numpy-friendly helpers ------------------------------------------------
def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_shardmap(nnodes: int) -> np.ndarray:
    m = np.empty(4096, np.int32)
    lib().orc_default_shardmap(_ptr(m), nnodes)
    return m


def route_nodes(keys: np.ndarray, gx_type: int, nnodes: int) -> np.ndarray:
    """Datanode of each key under the reference SHARD rule (default map)."""
    L = lib()
    m = default_shardmap(nnodes)
    out = np.empty(len(keys), np.int32)
    mp = _ptr(m)
    if gx_type == GX_FLOAT8:
        datums = np.asarray(keys, np.float64).view(np.int64)
    else:
        datums = np.asarray(keys).astype(np.int64)
    for i, k in enumerate(datums.tolist()):
        out[i] = L.orc_route_node(mp, gx_type, k, 0)
    return out


def gen_orders(sf, o0, o1, node=0, nnodes=1):
    n = o1 - o0
    cols = [np.empty(n, np.int64), np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)]
    k = lib().orc_gen_orders(sf, o0, o1, node, nnodes, *[_ptr(c) for c in cols])
    return [c[:k] for c in cols]


def gen_lineitem(sf, o0, o1, node=0, nnodes=1):
    L = lib()
    n = L.orc_gen_lineitem_count(sf, o0, o1, node, nnodes)
    cols = [np.empty(n, np.int64)] + [np.empty(n, np.float64) for _ in range(4)] + \
           [np.empty(n, np.int32), np.empty(n, np.int8), np.empty(n, np.int8)]
    k = L.orc_gen_lineitem(sf, o0, o1, node, nnodes, *[_ptr(c) for c in cols])
    assert k == n
    return cols


def gen_customer(sf, c0, c1, node=0, nnodes=1):
    n = c1 - c0
    cols = [np.empty(n, np.int32), np.empty(n, np.int8)]
    k = lib().orc_gen_customer(sf, c0, c1, node, nnodes, *[_ptr(c) for c in cols])
    return [c[:k] for c in cols]


class Rel:
    """An OpenTenBase heap relation (8 KB pages) built from columns."""

    def __init__(self, types, cols=None, nulls=None):
        self.types = list(types)
        arr = (C.c_int32 * len(types))(*types)
        self.h = lib().orc_rel_create(len(types), arr)
        if cols is not None:
            self.insert(cols, nulls)

    def insert(self, cols, nulls=None):
        n = len(cols[0]) if cols else 0
        keep = [np.ascontiguousarray(c, NP_DTYPES[t]) for c, t in zip(cols, self.types)]
        cp = (C.c_void_p * len(keep))(*[c.ctypes.data for c in keep])
        npp = None
        if nulls is not None:
            kn = [None if x is None else np.ascontiguousarray(x, np.uint8) for x in nulls]
            npp = (C.c_void_p * len(kn))(*[None if x is None else x.ctypes.data for x in kn])
        rc = lib().orc_rel_insert_columns(self.h, cp, npp, n)
        assert rc == 0

    def add_column(self, gx_type):
        """ALTER TABLE ADD COLUMN without a rewrite: old tuples read the new attribute as NULL"""
        assert lib().orc_rel_add_column(self.h, gx_type) == 0
        self.types.append(gx_type)

    @property
    def ntuples(self):
        return lib().orc_rel_ntuples(self.h)

    @property
    def npages(self):
        return lib().orc_rel_npages(self.h)

    def pages(self) -> np.ndarray:
        out = np.empty(self.npages * 8192, np.uint8)
        lib().orc_rel_copy_pages(self.h, _ptr(out))
        return out

    def delete(self, pageno, lineoff):
        return lib().orc_rel_delete_tuple(self.h, pageno, lineoff)

    def scan(self, attnums):
        n = self.ntuples
        cols = [np.empty(n, NP_DTYPES[self.types[a]]) for a in attnums]
        nulls = [np.zeros(n, np.uint8) for _ in attnums]
        an = (C.c_int32 * len(attnums))(*attnums)
        cp = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        npp = (C.c_void_p * len(cols))(*[c.ctypes.data for c in nulls])
        k = lib().orc_rel_scan_columns(self.h, len(attnums), an, cp, npp)
        return [c[:k] for c in cols], [x[:k] for x in nulls]

    def free(self):
        if self.h:
            lib().orc_rel_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fnpage_pack(types, cols, nulls=None, page_id=None, end_marker=True) -> np.ndarray:
    """columns -> forward-node pages as the reference's sender fills them; returns a (npages, 8192) uint8 array"""
    n = len(cols[0]) if cols else 0
    keep = [np.ascontiguousarray(c, NP_DTYPES[t]) for c, t in zip(cols, types)]
    cp = (C.c_void_p * len(keep))(*[c.ctypes.data for c in keep])
    npp, kn = None, None
    if nulls is not None:
        kn = [None if x is None else np.ascontiguousarray(x, np.uint8) for x in nulls]
        npp = (C.c_void_p * len(kn))(*[None if x is None else x.ctypes.data for x in kn])
    tarr = (C.c_int32 * len(types))(*types)
    pid = page_id or OrcFnPageId()
    cap = n // 20 + 4                                   # a tuple takes at least 24 bytes: <= 340 per page
    while True:
        out = np.empty((cap, 8192), np.uint8)
        k = lib().orc_fnpage_pack(len(types), tarr, cp, npp, n, C.byref(pid), int(bool(end_marker)), _ptr(out), cap)
        if k == -1:
            cap *= 2
            continue
        assert k >= 0, k
        return out[:k].copy()


def fnpage_unpack(pages: np.ndarray, types):
    """forward-node pages -> (columns, null arrays) as the reference's receiver reads them"""
    pages = np.ascontiguousarray(pages, np.uint8).reshape(-1, 8192)
    cap = len(pages) * 510 + 1                          # a tuple takes at least 16 bytes (header only: every attribute NULL)
    cols = [np.empty(cap, NP_DTYPES[t]) for t in types]
    nulls = [np.zeros(cap, np.uint8) for _ in types]
    tarr = (C.c_int32 * len(types))(*types)
    cp = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    npp = (C.c_void_p * len(cols))(*[c.ctypes.data for c in nulls])
    k = lib().orc_fnpage_unpack(_ptr(pages), len(pages), len(types), tarr, cp, npp, cap)
    assert k >= 0, k
    return [c[:k] for c in cols], [x[:k] for x in nulls]


def make_plan(preds=(), outer_key_col=-1, group_cols=(), aggs=(), est_groups=0, strategy=0) -> GxAggPlan:
    """preds: (col, op, const[, is_float]); group_cols: (side, col);
    aggs: (fn, [ (op, col, k), ... ])"""
    p = GxAggPlan()
    p.n_preds = len(preds)
    p.outer_key_col = outer_key_col
    for i, pr in enumerate(preds):
        p.preds[i] = mk_pred(*pr)
    p.n_group_cols = len(group_cols)
    for i, (side, col) in enumerate(group_cols):
        p.group_cols[i].side = side
        p.group_cols[i].col = col
    p.n_aggs = len(aggs)
    for i, (fn, ops) in enumerate(aggs):
        p.aggs[i].fn = fn
        p.aggs[i].arg.nops = len(ops)
        for j, (op, col, k) in enumerate(ops):
            p.aggs[i].arg.ops[j].op = op
            p.aggs[i].arg.ops[j].col = col
            p.aggs[i].arg.ops[j].k = float(k)
    p.est_groups = est_groups
    p.strategy = strategy
    return p


def mk_pred(col, op, const, is_float=False) -> GxPred:
    q = GxPred()
    q.col, q.op = col, op
    if is_float:
        q.fval = float(const)
    else:
        q.ival = int(const)
    return q


def make_join(inner_key_col, payload_cols=(), inner_unique=0, inner_preds=(), jointype=0) -> OrcJoinSpec:
    j = OrcJoinSpec()
    j.jointype = jointype
    j.inner_key_col = inner_key_col
    j.n_inner_preds = len(inner_preds)
    for i, pr in enumerate(inner_preds):
        j.inner_preds[i] = mk_pred(*pr)
    j.n_payload = len(payload_cols)
    for i, c in enumerate(payload_cols):
        j.payload_cols[i] = c
    j.inner_unique = inner_unique
    return j


class AggResult:
    """keys int64[ngroups, ng], aggs float64[ngroups, na] (int results bit-cast),
    nulls uint8[ngroups, ng+na], states float64[ngroups, na, 3]"""

    def __init__(self, keys, aggs, nulls, states=None):
        self.keys, self.aggs, self.nulls, self.states = keys, aggs, nulls, states

    @property
    def ngroups(self):
        return self.keys.shape[0]

    def sorted(self):
        ng = self.keys.shape[1]
        if ng == 0 or self.ngroups == 0:
            return self
        order = np.lexsort([self.keys[:, c] for c in reversed(range(ng))] +
                           [self.nulls[:, c] for c in reversed(range(ng))])
        return AggResult(self.keys[order], self.aggs[order], self.nulls[order],
                         None if self.states is None else self.states[order])


def _result_to_np(r: OrcResult) -> AggResult:
    n, ng, na = r.ngroups, r.n_group_cols, r.n_aggs
    keys = np.ctypeslib.as_array(r.keys, (n * max(ng, 1),))[: n * ng].reshape(n, ng).copy() if n else np.zeros((0, ng), np.int64)
    aggs = np.ctypeslib.as_array(r.aggs, (n * max(na, 1),))[: n * na].reshape(n, na).copy() if n else np.zeros((0, na))
    nulls = np.ctypeslib.as_array(r.nulls, (n * max(ng + na, 1),))[: n * (ng + na)].reshape(n, ng + na).copy() if n else np.zeros((0, ng + na), np.uint8)
    states = np.ctypeslib.as_array(r.states, (n * max(na, 1) * 3,))[: n * na * 3].reshape(n, na, 3).copy() if n else np.zeros((0, na, 3))
    return AggResult(keys, aggs, nulls, states)


def exec_agg(outer: Rel, plan: GxAggPlan, inner: Rel | None = None, join: OrcJoinSpec | None = None,
             keep_raw: bool = False):
    r = OrcResult()
    rc = lib().orc_exec_agg(outer.h, inner.h if inner else None,
                            C.byref(join) if join is not None else None, C.byref(plan), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"oracle exec failed: status {rc}")
    out = _result_to_np(r)
    if keep_raw:
        return out, r
    lib().orc_result_free(C.byref(r))
    return out


def combine(plan: GxAggPlan, raws) -> AggResult:
    arr = (OrcResult * len(raws))(*raws)
    out = OrcResult()
    rc = lib().orc_combine_results(C.byref(plan), arr, len(raws), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle combine failed: status {rc}")
    res = _result_to_np(out)
    lib().orc_result_free(C.byref(out))
    return res


def exec_join(outer: Rel, outer_key_col, inner: Rel, join: OrcJoinSpec, out_outer_cols, outer_preds=()):
    L = lib()
    pr = (GxPred * max(len(outer_preds), 1))(*[mk_pred(*p) for p in outer_preds])
    oc = (C.c_int32 * max(len(out_outer_cols), 1))(*out_outer_cols)
    n = L.orc_exec_join(outer.h, outer_key_col, len(outer_preds), pr, inner.h, C.byref(join),
                        len(out_outer_cols), oc, None)
    ncols = len(out_outer_cols) + join.n_payload
    cols = [np.empty(n, np.int64) for _ in range(ncols)]
    cp = (C.c_void_p * max(ncols, 1))(*[c.ctypes.data for c in cols])
    n2 = L.orc_exec_join(outer.h, outer_key_col, len(outer_preds), pr, inner.h, C.byref(join),
                         len(out_outer_cols), oc, cp)
    assert n2 == n
    return cols


def last_exec_seconds() -> float:
    return lib().orc_last_exec_seconds()


# ---- whole-query references for the multi-datanode plan shapes ---------------
def q3_reference(sf, n_orders, n_cust, date, segment, orders_cols=None):
    """TPC-H Q3 shape on ONE node, tuple at a time: the customer semi-join is restated with
    numpy (c_custkey is unique), the orders JOIN lineitem + GROUP BY runs through the
    tuple-at-a-time executor.  Returns the sorted AggResult over
    (l_orderkey, o_orderdate, o_shippriority): sum(l_extendedprice * (1 - l_discount))."""
    GX_OP_COL, GX_OP_CONST, GX_OP_SUB, GX_OP_MUL, GX_AGG_SUM_F8, GX_GT = 1, 2, 4, 5, 3, 5
    c = gen_customer(sf, 0, n_cust)
    o = gen_orders(sf, 0, n_orders)
    l = gen_lineitem(sf, 0, n_orders)
    good = np.isin(o[1], c[0][c[1] == segment]) & (o[2] < date)
    oj = [x[good] for x in o]
    rev = [(GX_OP_COL, 2, 0), (GX_OP_CONST, 0, 1.0), (GX_OP_COL, 3, 0), (GX_OP_SUB, 0, 0), (GX_OP_MUL, 0, 0)]
    plan = make_plan(preds=[(5, GX_GT, date)], outer_key_col=0, group_cols=[(0, 0), (1, 0), (1, 1)],
                     aggs=[(GX_AGG_SUM_F8, rev)])
    ltypes = [GX_INT8, GX_FLOAT8, GX_FLOAT8, GX_FLOAT8, GX_FLOAT8, GX_DATE, GX_CHAR, GX_CHAR]
    otypes = [GX_INT8, GX_INT4, GX_DATE, GX_INT4]
    return exec_agg(Rel(ltypes, l), plan, Rel(otypes, oj), make_join(0, payload_cols=[2, 3], inner_unique=1)).sorted()
