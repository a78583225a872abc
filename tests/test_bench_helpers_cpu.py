"""bench.py's host-side helpers must not depend on a GPU box being well-behaved: the clock
sampler has to cope with a missing nvidia-smi and with a timed region shorter than its period,
the placement report with unreadable /proc files, the CPU count with cgroup files of either
version."""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
sys.modules["bench_mod"] = bench
spec.loader.exec_module(bench)


class _Proc:
    def terminate(self): pass
    def wait(self, timeout=None): return 0
    def kill(self): pass


def _sampler_with(tmp_path, lines):
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.path = str(tmp_path / "clocks.csv")
    open(s.path, "w").write("\n".join(lines) + "\n")
    s.p = _Proc()
    return s


def _stamp(t):
    lt = time.localtime(t)
    return time.strftime("%Y/%m/%d %H:%M:%S", lt) + ".%03d" % int((t % 1) * 1000)


def test_clock_sampler_picks_samples_inside_the_timed_region(tmp_path):
    t0 = time.time()
    rows = [f"{_stamp(t0 - 1.0)}, 1200, 1965, 300.0, 0x0, Not Active, Not Active, Not Active, Not Active",
            f"{_stamp(t0 + 0.010)}, 1965, 1965, 700.0, 0x4, Not Active, Not Active, Not Active, Active",
            f"{_stamp(t0 + 0.030)}, 1950, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, Not Active",
            "garbage line"]
    out = _sampler_with(tmp_path, rows).stop(t0, t0 + 0.05)
    assert out["samples"] == 2 and out["window"] == "timed region"
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_falls_back_to_the_warm_up_window(tmp_path):
    t0 = time.time()
    rows = [f"{_stamp(t0 - 0.5)}, 1965, 1965, 650.0, 0x0, Not Active, Not Active, Not Active, Not Active"]
    out = _sampler_with(tmp_path, rows).stop(t0, t0 + 0.004)          # region shorter than the sampling period
    assert out["samples"] == 1 and out["window"].startswith("warm-up")


def test_clock_sampler_without_nvidia_smi():
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.p = None
    assert s.stop(0.0, 1.0) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


def test_staging_placement_never_raises():
    info = bench.staging_placement([0x7F0000000000])                  # no such mapping; nvidia-smi may be absent
    assert isinstance(info, dict)


def test_effective_cpus_is_sane():
    n = bench.effective_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_same_result_is_strict_about_integers_and_tolerant_about_floats():
    import numpy as np
    import bench
    import oracle as O
    keys = np.array([[3], [1], [2]], np.int64)
    aggs = np.zeros((3, 2)); aggs[:, 0] = np.array([30, 10, 20], np.int64).view(np.float64); aggs[:, 1] = [3.0, 1.0, 2.0]
    want = O.AggResult(np.array([[1], [2], [3]], np.int64),
                       np.stack([np.array([10, 20, 30], np.int64).view(np.float64), np.array([1.0, 2.0, 3.0 * (1 + 1e-12)])], 1),
                       np.zeros((3, 3), np.uint8))
    ok, _ = bench.same_result((keys, aggs, None), want, int_aggs=(0,))
    assert ok
    bad = aggs.copy(); bad[0, 0] = np.array([31], np.int64).view(np.float64)[0]
    assert not bench.same_result((keys, bad, None), want, int_aggs=(0,))[0]
    bad = aggs.copy(); bad[1, 1] = 1.0 + 1e-6
    assert not bench.same_result((keys, bad, None), want, int_aggs=(0,))[0]
    assert not bench.same_result((keys[:2], aggs[:2], None), want, int_aggs=(0,))[0]
