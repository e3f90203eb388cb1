"""bench.py host logic that needs no GPU: how many batches a launch carries for a given --steps."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_batches_per_launch_times_exactly_k_steps():
    f = _bench().choose_batches_per_launch
    assert f(1024, 16) == 32 and f(512, 16) == 32 and f(256, 16) == 32 and f(64, 16) == 32
    assert f(20, 16) == 10 and f(10, 16) == 5 and f(100, 16) == 25  # the driver's --steps 20: two launches of ten batches
    assert f(7, 16) == 7 and f(1, 16) == 1 and f(37, 16) == 1       # a prime K <= 32 is one launch, beyond that single batches
    assert f(64, 16, requested=4) == 4
    for k in range(1, 300):
        b = f(k, 16)
        assert 1 <= b <= 32 and k % b == 0
