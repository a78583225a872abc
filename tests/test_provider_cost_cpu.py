"""The provider's cost model (opentenbase_b200/provider/gpuexec_cost.h) next to the reference's own formulas.

gpuexec_cost.h is pure C; it is compiled here into a tiny shared object and called through ctypes.  The CPU side of
every comparison restates the reference's arithmetic for the same sub-plan:

* cost_seqscan   src/backend/optimizer/path/costsize.c:327-394   seq_page_cost*pages + (cpu_tuple_cost + quals)*tuples
* cost_agg       :2451-2540 (AGG_HASHED)                          input + (transCost + cpu_operator_cost*groupcols)*tuples + cpu_tuple_cost*groups
* final_cost_hashjoin :3788                                       both inputs + cpu_operator_cost per hash clause per tuple + cpu_tuple_cost per joined row

with the planner's default GUCs (seq_page_cost 1, cpu_tuple_cost 0.01, cpu_operator_cost 0.0025, costsize.c:100-112).
What is checked is the DECISION add_path() will take: large scans/joins go to the GPU path, small ones stay on the CPU,
the crossover moves the right way with every rate, and costs are per datanode like the reference's."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "opentenbase_b200", "provider", "gpuexec_cost.h")

SEQ_PAGE, CPU_TUPLE, CPU_OP = 1.0, 0.01, 0.0025


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("seq_page_cost", "cpu_tuple_cost", "cost_unit_us", "host_page_us", "pcie_gb_s", "hbm_gb_s", "startup_us")]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("cost")
    src = d / "cost.c"
    src.write_text('#include "%s"\n'
                   'void defaults(gpuexec_cost_params *p) { gpuexec_default_cost_params(p); }\n'
                   'void cost(const gpuexec_cost_params *p, double pages, double bytes, double groups, double *s, double *t)\n'
                   '{ gpuexec_path_cost(p, pages, bytes, groups, s, t); }\n' % HDR)
    so = d / "cost.so"
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-shared", "-fPIC", "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.cost.argtypes = [C.POINTER(Params), C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.defaults.argtypes = [C.POINTER(Params)]
    return L


def gpu_cost(lib, pages, staged_bytes, groups, **over):
    p = Params()
    lib.defaults(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    s, t = C.c_double(), C.c_double()
    lib.cost(C.byref(p), pages, staged_bytes, groups, C.byref(s), C.byref(t))
    return s.value, t.value


def cpu_seqscan(pages, tuples, nquals=0):
    return SEQ_PAGE * pages + (CPU_TUPLE + nquals * CPU_OP) * tuples


def cpu_hashagg(input_total, tuples, ngroupcols, naggs, groups):
    # transCost.per_tuple: one cpu_operator_cost per transition function call (count_agg_clauses / get_agg_clause_costs)
    startup = input_total + (naggs * CPU_OP) * tuples + (CPU_OP * ngroupcols) * tuples
    return startup, startup + CPU_TUPLE * groups


def cpu_hashjoin(outer_total, inner_total, outer_rows, inner_rows, joined_rows):
    # initial_cost_hashjoin: hash both sides (cpu_operator_cost per clause per tuple) + cpu_tuple_cost per inner tuple inserted;
    # final_cost_hashjoin: ~0.5 bucket comparisons per outer tuple + cpu_tuple_cost per joined row
    return (outer_total + inner_total + CPU_OP * (outer_rows + inner_rows) + CPU_TUPLE * inner_rows +
            CPU_OP * 0.5 * outer_rows + CPU_TUPLE * joined_rows)


# TPC-H shapes per scale factor (lineitem ~ 6 M rows / ~110 k pages at 150 B per row, orders 1.5 M / ~26 k)
def lineitem(sf):
    return 6.0e6 * sf, 110_000.0 * sf


def orders(sf):
    return 1.5e6 * sf, 26_000.0 * sf


def test_defaults_are_the_measured_rates(lib):
    p = Params()
    lib.defaults(C.byref(p))
    assert (p.seq_page_cost, p.cpu_tuple_cost) == (SEQ_PAGE, CPU_TUPLE)
    assert 40 <= p.pcie_gb_s <= 56 and 3000 <= p.hbm_gb_s <= 5800 and p.host_page_us > 0 and p.startup_us > 0


@pytest.mark.parametrize("sf", [1, 10, 100])
def test_config2_shape_goes_to_the_gpu(lib, sf):
    """SUM(l_extendedprice) GROUP BY l_shipdate: the disk term is common, the GPU path drops the per-tuple CPU terms."""
    tuples, pages = lineitem(sf)
    _, cpu_total = cpu_hashagg(cpu_seqscan(pages, tuples), tuples, 1, 1, 2526)
    gs, gt = gpu_cost(lib, pages, tuples * 12, 2526)
    assert gs < gt < cpu_total
    assert gt >= SEQ_PAGE * pages                      # never cheaper than reading the pages
    assert gt - gs == pytest.approx(CPU_TUPLE * 2526)  # a blocking node: only the output rows come after start-up


@pytest.mark.parametrize("sf", [1, 100])
def test_config3_join_shape_goes_to_the_gpu(lib, sf):
    lt, lp = lineitem(sf)
    ot, op = orders(sf)
    join = cpu_hashjoin(cpu_seqscan(lp, lt), cpu_seqscan(op, ot), lt, ot, lt)
    _, cpu_total = cpu_hashagg(join, lt, 1, 2, 2406)
    _, gt = gpu_cost(lib, lp + op, lt * 16 + ot * 12 + ot * 32, 2406)
    assert gt < cpu_total
    assert cpu_total / gt > 1.3                       # the margin is the per-tuple CPU work, not a constant factor


def test_small_relations_stay_on_the_cpu(lib):
    """A 50-page table: the fixed price of a GPU sub-plan (plan compile, launches, result fetch) decides."""
    tuples, pages = 3000.0, 50.0
    _, cpu_total = cpu_hashagg(cpu_seqscan(pages, tuples), tuples, 1, 1, 10)
    _, gt = gpu_cost(lib, pages, tuples * 12, 10)
    assert gt > cpu_total


def test_crossover_moves_with_the_rates(lib):
    def crossover(**over):
        lo, hi = 1.0, 1e7                               # pages; 55 tuples per page
        for _ in range(60):
            mid = (lo * hi) ** 0.5
            t = mid * 55
            cpu = cpu_hashagg(cpu_seqscan(mid, t), t, 1, 1, 100)[1]
            g = gpu_cost(lib, mid, t * 12, 100, **over)[1]
            lo, hi = (mid, hi) if g > cpu else (lo, mid)
        return hi
    base = crossover()
    assert 50 < base < 5000                            # a few hundred pages (a few MB) with the default rates
    assert crossover(startup_us=15000.0) > 5 * base    # a slower start-up keeps more queries on the CPU
    assert crossover(cost_unit_us=100.0) < base        # a slow disk makes everything but the disk term cheap
    assert crossover(host_page_us=6.0) > base          # a slow loader


def test_feed_is_the_slower_of_host_and_pcie_not_their_sum(lib):
    pages = 1e6
    a = gpu_cost(lib, pages, 0, 1, host_page_us=0.1, pcie_gb_s=50.0, startup_us=0.0)[1] - pages
    b = gpu_cost(lib, pages, 0, 1, host_page_us=0.1, pcie_gb_s=25.0, startup_us=0.0)[1] - pages
    c = gpu_cost(lib, pages, 0, 1, host_page_us=0.3, pcie_gb_s=50.0, startup_us=0.0)[1] - pages
    pcie_us = pages * 8192 / 50e3
    assert a == pytest.approx(pcie_us / 10.0 + 0.01, rel=1e-6)        # PCIe-bound: 0.16 us per page > 0.1
    assert b == pytest.approx(2 * pcie_us / 10.0 + 0.01, rel=1e-6)
    assert c == pytest.approx(0.3 * pages / 10.0 + 0.01, rel=1e-6)    # host-bound


def test_costs_are_per_datanode(lib):
    """The hook divides pages and bytes by path_count_datanodes() as cost_seqscan does through PAGES_PER_DN."""
    pages, tuples = 110_000.0 * 100, 6e8
    one = gpu_cost(lib, pages, tuples * 12, 100)[1]
    eight = gpu_cost(lib, pages / 8, tuples * 12 / 8, 100)[1]
    assert eight < one / 7.5
    assert cpu_seqscan(pages / 8, tuples / 8) > eight  # and stays below the reference's per-datanode scan alone
