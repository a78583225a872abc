"""Comparison helpers shared by the parity tests: GPU result vs oracle result.

Bar (BASELINE.json north_star): group keys, COUNT and integer SUM bit-exact;
float8 SUM/AVG/MIN/MAX within 1e-9 relative."""
import numpy as np

import opentenbase_b200 as g
import oracle as O

FLOAT_RTOL = 1e-9
INT_FNS = {g.GX_AGG_COUNT_STAR, g.GX_AGG_COUNT, g.GX_AGG_SUM_I4, g.GX_AGG_SUM_I8}


def to_gpu_plan(oplan: O.GxAggPlan) -> g.GxAggPlan:
    """Same POD layout on both sides (include/gpuexec.h)."""
    return g.GxAggPlan.from_buffer_copy(bytes(oplan))


def sort_rows(keys, aggs, nulls):
    n, ng = keys.shape
    if n == 0 or ng == 0:
        return keys, aggs, nulls
    order = np.lexsort([keys[:, c] for c in reversed(range(ng))] + [nulls[:, c] for c in reversed(range(ng))])
    return keys[order], aggs[order], nulls[order]


def assert_agg_equal(plan, got, want: O.AggResult, rtol=FLOAT_RTOL):
    gk, ga, gn = sort_rows(*got)
    w = want.sorted()
    assert gk.shape == w.keys.shape, f"group count: gpu {gk.shape[0]} vs oracle {w.keys.shape[0]}"
    np.testing.assert_array_equal(gn, w.nulls, err_msg="NULL flags differ")
    np.testing.assert_array_equal(gk, w.keys, err_msg="group keys differ")
    for a in range(plan.n_aggs):
        fn = plan.aggs[a].fn
        notnull = w.nulls[:, plan.n_group_cols + a] == 0
        if fn in INT_FNS:
            np.testing.assert_array_equal(ga[:, a].view(np.int64)[notnull], w.aggs[:, a].view(np.int64)[notnull],
                                          err_msg=f"integer aggregate {a} (fn {fn}) not bit-exact")
        else:
            x, y = ga[:, a][notnull], w.aggs[:, a][notnull]
            both_nan = np.isnan(x) & np.isnan(y)
            np.testing.assert_allclose(x[~both_nan], y[~both_nan], rtol=rtol, atol=0,
                                       err_msg=f"float8 aggregate {a} (fn {fn}) beyond {rtol} relative")


def lineitem_rel(cols):
    return O.Rel(g.SCHEMAS[g.T_LINEITEM], cols)


def orders_rel(cols):
    return O.Rel(g.SCHEMAS[g.T_ORDERS], cols)
