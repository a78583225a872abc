"""Forward-node pages (the reference's redistribute wire format) on the device, against
 * tests/golden/fnpage_vectors.json - pages written by the reference's own heaptuple.o + fnbufpage.o, and
 * oracle/orc_fnpage.c - the restatement pinned to those objects (tests/test_oracle_vs_ref.py).
Sender (gx_fnpage_pack): byte-identical pages when no column carries NULLs; with NULLs the pages are checked by
reading them back with the oracle's receiver.  Receiver (gx_fnpage_unpack): reads the reference's pages."""
import json
import os

import numpy as np
import pytest

import opentenbase_b200 as g
import oracle as O

pytestmark = pytest.mark.gpu
ATT = {O.GX_INT8: (8, 8), O.GX_INT4: (4, 4), O.GX_FLOAT8: (8, 8), O.GX_DATE: (4, 4), O.GX_CHAR: (1, 1), O.ORC_BPCHAR1: (-1, 4)}


def _cases():
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fnpage_vectors.json")))["cases"]


def _arrays(case):
    types = case["types"]
    vals = np.array(case["values"], np.int64).reshape(-1, len(types))
    isn = np.array(case["isnull"], np.uint8).reshape(-1, len(types))
    cols = [vals[:, i].copy().view(np.float64) if t == O.GX_FLOAT8 else vals[:, i].astype(O.NP_DTYPES[t]) for i, t in enumerate(types)]
    pages = np.zeros((len(case["pages_used_hex"]), 8192), np.uint8)
    for p, h in enumerate(case["pages_used_hex"]):
        b = np.frombuffer(bytes.fromhex(h), np.uint8)
        pages[p, :len(b)] = b
    return types, cols, isn, pages


def _gx_types(types):
    return [g.GX_CHAR if t == O.ORC_BPCHAR1 else t for t in types]


def _pid(d):
    return g.GxFnPageId(d["qid_ts"], d["qid_seq"], d["fid"], d["nodeid"], d["workerid"], d["virtualid"], 0)


def _table(ctx, types, cols, isn):
    nulls = [isn[:, i].copy() if isn[:, i].any() else None for i in range(len(types))]
    return ctx.table_from(_gx_types(types), cols, nulls if any(x is not None for x in nulls) else None)


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def test_pack_equals_reference_pages_when_no_nulls(ctx):
    done = 0
    for case in _cases():
        types, cols, isn, want = _arrays(case)
        if isn.any():
            continue
        t = _table(ctx, types, cols, isn)
        got = ctx.fnpage_pack(t, [ATT[x][0] for x in types], [ATT[x][1] for x in types], _pid(case["page_id"]), case["end_marker"])
        t.free()
        assert got.shape == want.shape, case["name"]
        np.testing.assert_array_equal(got, want, err_msg=case["name"])
        done += 1
    assert done >= 3


def test_pack_with_nulls_is_read_back_by_the_oracle_receiver(ctx):
    done = 0
    for case in _cases():
        types, cols, isn, _ = _arrays(case)
        if not isn.any():
            continue
        t = _table(ctx, types, cols, isn)
        got = ctx.fnpage_pack(t, [ATT[x][0] for x in types], [ATT[x][1] for x in types], _pid(case["page_id"]), case["end_marker"])
        t.free()
        # page headers: FnPageInit's fields, lower inside the page, FNPAGE_END only on the last page
        for p in range(len(got)):
            lower = int(got[p, 0:4].copy().view(np.uint32)[0])
            assert 32 <= lower <= 8192 and not got[p, lower:].any()
            assert int(got[p, 4:6].copy().view(np.uint16)[0]) == case["page_id"]["fid"]
            assert int(got[p, 8:16].copy().view(np.int64)[0]) == case["page_id"]["qid_ts"]
            assert (int(got[p, 24:28].copy().view(np.uint32)[0]) == 8) == (p == len(got) - 1)
        c2, n2 = O.fnpage_unpack(got, types)
        assert len(c2[0]) == len(cols[0]), case["name"]
        for i in range(len(types)):
            np.testing.assert_array_equal(n2[i], isn[:, i], err_msg=case["name"])
            keep = isn[:, i] == 0
            np.testing.assert_array_equal(np.asarray(c2[i])[keep].view(np.uint8), np.asarray(cols[i])[keep].view(np.uint8), err_msg=case["name"])
        done += 1
    assert done >= 2


def test_unpack_reads_the_reference_pages(ctx):
    for case in _cases():
        types, cols, isn, pages = _arrays(case)
        t = ctx.fnpage_unpack(pages, [ATT[x][0] for x in types], [ATT[x][1] for x in types], list(range(len(types))), _gx_types(types))
        assert t.nrows == len(cols[0]), case["name"]
        for i in range(len(types)):
            v, nl = t.read(i, with_nulls=True)
            np.testing.assert_array_equal(nl, isn[:, i], err_msg=case["name"])
            keep = isn[:, i] == 0
            np.testing.assert_array_equal(np.asarray(v)[keep].view(np.uint8), np.asarray(cols[i])[keep].view(np.uint8), err_msg=case["name"])
        t.free()


def test_unpack_projects_and_honours_not_null(ctx):
    case = _cases()[0]                                     # no NULLs
    types, cols, isn, pages = _arrays(case)
    want = [3, 0]                                          # two of the four attributes, reordered
    t = ctx.fnpage_unpack(pages, [ATT[x][0] for x in types], [ATT[x][1] for x in types], want, [_gx_types(types)[a] for a in want], notnull=[1] * len(types))
    for c, a in enumerate(want):
        np.testing.assert_array_equal(t.read(c), cols[a])
    t.free()
    case = _cases()[1]                                     # NULLs present but every attribute declared NOT NULL
    types, cols, isn, pages = _arrays(case)
    with pytest.raises(g.GxError):
        ctx.fnpage_unpack(pages, [ATT[x][0] for x in types], [ATT[x][1] for x in types], list(range(len(types))), _gx_types(types), notnull=[1] * len(types))


def test_unpack_rejects_huge_and_corrupt_pages(ctx):
    types, cols, isn, pages = _arrays(_cases()[0])
    al, ag = [ATT[x][0] for x in types], [ATT[x][1] for x in types]
    bad = pages.copy(); bad[0, 24] |= 1                    # FNPAGE_HUGE
    with pytest.raises(g.GxError):
        ctx.fnpage_unpack(bad, al, ag, list(range(len(types))), _gx_types(types))
    bad = pages.copy(); bad[0, 32:36] = np.frombuffer(np.uint32(9000).tobytes(), np.uint8)   # a tuple longer than the page
    with pytest.raises(g.GxError):
        ctx.fnpage_unpack(bad, al, ag, list(range(len(types))), _gx_types(types))
    bad = pages.copy(); bad[0, 0:4] = np.frombuffer(np.uint32(70000).tobytes(), np.uint8)    # lower past the page
    with pytest.raises(g.GxError):
        ctx.fnpage_unpack(bad, al, ag, list(range(len(types))), _gx_types(types))


def test_round_trip_at_size_and_against_the_oracle_sender(ctx):
    """3 M orders rows (the Q3 redistribute's tuple: orderkey, custkey, orderdate, shippriority) -> pages -> rows;
    the first pages equal the oracle sender's (which is pinned to the reference's objects); page count is the closed form."""
    n = 3_000_000
    t = ctx.table(g.SCHEMAS[g.T_ORDERS], n).generate(g.T_ORDERS, 2, 0, n)
    types = list(t.types)
    al, ag = [ATT[x][0] for x in types], [ATT[x][1] for x in types]
    pid = g.GxFnPageId(42, 43, 5, 1, 0, 0, 0)
    pages = ctx.fnpage_pack(t, al, ag, pid, True)
    tuple_len = int(pages[0, 32:36].copy().view(np.uint32)[0])
    per_page = 8160 // ((tuple_len + 7) // 8 * 8)
    assert len(pages) == (t.nrows + per_page - 1) // per_page + (1 if (t.nrows % per_page == 0 or 8192 - 32 - (t.nrows % per_page) * ((tuple_len + 7) // 8 * 8) < 4) else 0)
    k = 5000
    head = [t.read(c)[:k] for c in range(len(types))]
    want = O.fnpage_pack(types, head, None, O.OrcFnPageId(42, 43, 5, 1, 0, 0, 0), False)
    np.testing.assert_array_equal(pages[: len(want) - 1], want[:-1])          # all but the oracle's partly filled last page
    back = ctx.fnpage_unpack(pages, al, ag, list(range(len(types))), types, notnull=[1] * len(types))
    assert back.nrows == t.nrows
    for c in range(len(types)):
        np.testing.assert_array_equal(back.read(c), t.read(c))
    back.free(); t.free()


def test_round_trip_with_random_nulls(ctx):
    rng = np.random.default_rng(5)
    n = 400_000
    types = [g.GX_INT8, g.GX_INT4, g.GX_FLOAT8, g.GX_CHAR, g.GX_DATE, g.GX_INT8, g.GX_CHAR, g.GX_INT4, g.GX_FLOAT8, g.GX_INT8]
    cols = [rng.integers(-2**62, 2**62, n) if x == g.GX_INT8 else rng.integers(-2**31, 2**31 - 1, n).astype(np.int32) if x in (g.GX_INT4, g.GX_DATE)
            else rng.normal(size=n) if x == g.GX_FLOAT8 else rng.integers(-128, 127, n).astype(np.int8) for x in types]
    nulls = [(rng.random(n) < 0.3).astype(np.uint8) if i % 2 else None for i in range(len(types))]
    t = ctx.table_from(types, cols, nulls)
    wire = [O.ORC_BPCHAR1 if (x == g.GX_CHAR and i == 6) else x for i, x in enumerate(types)]      # one char column travels as bpchar(1)
    al, ag = [ATT[x][0] for x in wire], [ATT[x][1] for x in wire]
    pages = ctx.fnpage_pack(t, al, ag, g.GxFnPageId(1, 2, 3, 4, 5, 6, 0), True)
    c2, n2 = O.fnpage_unpack(pages[:50], wire)                                                     # the oracle reads the first pages
    for i in range(len(types)):
        k = len(c2[i])
        want_null = nulls[i][:k] if nulls[i] is not None else np.zeros(k, np.uint8)
        np.testing.assert_array_equal(n2[i], want_null)
        keep = want_null == 0
        np.testing.assert_array_equal(np.asarray(c2[i])[keep].view(np.uint8), np.asarray(cols[i])[:k][keep].view(np.uint8))
    back = ctx.fnpage_unpack(pages, al, ag, list(range(len(types))), types)
    assert back.nrows == n
    for i in range(len(types)):
        v, nl = back.read(i, with_nulls=True)
        want_null = nulls[i] if nulls[i] is not None else np.zeros(n, np.uint8)
        np.testing.assert_array_equal(nl, want_null)
        keep = want_null == 0
        np.testing.assert_array_equal(np.asarray(v)[keep].view(np.uint8), np.asarray(cols[i])[keep].view(np.uint8))
    back.free(); t.free()
