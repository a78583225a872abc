"""Coordinator -> datanode plan shipping of the CustomScan provider, without a GPU.

`provider/harness/gpuexec_harness --ship-only <case>` serialises the plan descriptor into `custom_private` Value nodes
(gpuexec_serialise), writes them the way nodeToString() does for a List of Values (nodes/outfuncs.c:443-477, 5103-5117),
reads the text back with the reference's OWN parser — src/backend/nodes/read.c (stringToNode / nodeRead), compiled from
where it lies — deserialises and serialises again; the two wire texts and every qual constant must be identical.

What this pins: 64-bit qual constants beyond 2^53 and beyond int32 (nodeRead turns wide integer tokens into T_Float
nodes, and _outValue prints T_Integer with %d), float8 constants that "%.17g" cannot carry through nodeRead's number
test (NaN, +-Infinity), negative zero, denormals, est_groups above 2^32.  The binary is built where /root/reference
exists (make -C opentenbase_b200/provider harness)."""
import math
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import opentenbase_b200 as g
from opentenbase_b200 import plans as P
import oracle as O
from test_provider_harness import HARNESS, write_case

pytestmark = pytest.mark.skipif(not os.path.exists(HARNESS), reason="harness binary not built (needs /root/reference at build time)")

OTYPES = [O.GX_INT8, O.GX_INT4, O.GX_DATE, O.GX_INT4]
LTYPES = [O.GX_INT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_FLOAT8, O.GX_DATE, O.ORC_BPCHAR1, O.ORC_BPCHAR1]


def ship(case):
    r = subprocess.run([HARNESS, "--ship-only", case], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    head, wire = r.stdout.splitlines()[:2]
    assert head.startswith("ship ok")
    return wire


@pytest.fixture(scope="module")
def rels():
    o, l = O.gen_orders(1, 0, 200), O.gen_lineitem(1, 0, 200)
    return O.Rel(OTYPES, o), O.Rel(LTYPES, l)


def test_join_plan_with_extreme_constants_survives_the_reference_parser(rels):
    orel, lrel = rels
    big = 2**62 + 12345                                    # not representable as a double, far outside int32
    plan = g.make_plan(preds=[(0, g.GX_GE, -big), (1, g.GX_LT, float("inf"), True), (1, g.GX_NE, float("nan"), True),
                              (1, g.GX_GT, -0.0, True)],
                       outer_key_col=0, group_cols=[(1, 0)],
                       aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)]), (g.GX_AGG_AVG_F8, [(g.GX_OP_COL, 1, 0)])],
                       est_groups=2**40 + 7)
    with tempfile.TemporaryDirectory() as d:
        case = os.path.join(d, "c3.case")
        write_case(case, [(LTYPES, lrel), (OTYPES, orel)], {"rti": 1, "attnums": [0, 2]},
                   {"rti": 2, "attnums": [0, 2], "key_col": 0, "payload_cols": [1], "unique": True,
                    "preds": [(0, g.GX_LT, big), (1, g.GX_LT, -1752), (1, g.GX_GT, -(2**31) - 5), (1, g.GX_GE, 5e-324, True)]},
                   plan, False, [1082, 20, 701, 701])
        plan2 = g.make_plan(preds=[(1, g.GX_LE, 1.7976931348623157e308, True)], outer_key_col=0, group_cols=[(1, 0)],
                            aggs=[(g.GX_AGG_COUNT_STAR, [])], est_groups=3)
        case2 = os.path.join(d, "c3b.case")
        write_case(case2, [(LTYPES, lrel), (OTYPES, orel)], {"rti": 1, "attnums": [0, 2]},
                   {"rti": 2, "attnums": [0, 2], "key_col": 0, "payload_cols": [1], "unique": True, "preds": []}, plan2, False, [1082, 20])
        wire2 = ship(case2)
        wire = ship(case)
    toks = wire.strip("()").split(" ")
    assert all(t.lstrip("-").isdigit() for t in toks)     # nothing but integer tokens: nodeRead's number test takes every one
    assert str(big) in toks and str(-big) in toks and str(2**40 + 7) in toks
    bits = lambda x: str(struct.unpack("<q", struct.pack("<d", x))[0])
    for x in (float("inf"), -0.0, 5e-324):
        assert bits(x) in toks
    assert bits(1.7976931348623157e308) in wire2.strip("()").split(" ")
    assert wire2.strip("()").split(" ")[-1] == "3"          # a small est_groups comes back from nodeRead() as a T_Integer node
    nan_tokens = [t for t in toks if t.lstrip("-").isdigit() and abs(int(t)) < 2**63 and
                  math.isnan(struct.unpack("<d", struct.pack("<q", int(t)))[0])]
    assert nan_tokens                                      # the NaN travelled as its bit pattern


def test_q1_partial_plan_with_expression_constants(rels):
    _, lrel = rels
    plan = P.q1_plan(0, 1, 2, 3, 4, 5, 6)                  # (1 - l_discount), (1 + l_tax): float8 constants inside aggregate arguments
    with tempfile.TemporaryDirectory() as d:
        case = os.path.join(d, "q1.case")
        write_case(case, [(LTYPES, lrel)], {"rti": 1, "attnums": [1, 2, 3, 4, 5, 6, 7]}, None, plan, True,
                   [1042, 1042, 701, 701, 701, 701, 1022, 1022, 1022, 20])
        wire = ship(case)
    one = str(struct.unpack("<q", struct.pack("<d", 1.0))[0])
    assert wire.count(one) >= 3


def test_random_plans_round_trip(rels):
    orel, lrel = rels
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as d:
        for it in range(25):
            npred = int(rng.integers(0, 4))
            preds = []
            for _ in range(npred):
                if rng.random() < 0.5:
                    preds.append((0, int(rng.integers(0, 6)), int(rng.integers(-2**63, 2**63 - 1))))
                else:
                    preds.append((1, int(rng.integers(0, 6)), float(np.frombuffer(rng.bytes(8), np.float64)[0]), True))
            plan = g.make_plan(preds=preds, outer_key_col=0, group_cols=[(1, 0)],
                               aggs=[(g.GX_AGG_COUNT_STAR, []), (g.GX_AGG_SUM_F8, [(g.GX_OP_COL, 1, 0)])],
                               est_groups=int(rng.integers(1, 2**62)))
            case = os.path.join(d, f"r{it}.case")
            write_case(case, [(LTYPES, lrel), (OTYPES, orel)], {"rti": 1, "attnums": [0, 2]},
                       {"rti": 2, "attnums": [0, 2], "key_col": 0, "payload_cols": [1], "unique": True,
                        "preds": [(0, g.GX_LE, int(rng.integers(-2**63, 2**63 - 1)))]}, plan, bool(it & 1), [1082, 20, 701])
            ship(case)
