"""Helper (test infrastructure, run as a script by tests/test_bench_on_double.py): bench.py's GPU arm end to end on the
kernel TEST DOUBLE — every host-side line of build_workload / the timed loops / the instrumented pass / the JSON line
executes on a CPU.  The numbers it prints are meaningless (a CPU emulation); only that the line comes out, with the
contract's keys.        python tests/bench_on_double.py --config 2 --size 32 --batch 2 --steps 1 --warmup 1"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import torch  # noqa: E402

import double_plugin  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)

    def synchronize(self):
        pass


def main():
    from unittest import mock
    double_plugin.pytest_configure(None)
    with mock.patch("torch.cuda.Event", _Event), mock.patch.object(torch.Tensor, "pin_memory", lambda self: self):
        from joligen_b200 import lib as L
        L.load().jg_check_device = lambda: 0   # (the compute-capability probe of the real library)
        import bench
        sys.argv = ["bench.py", "--no-cpu-baseline", "--no-incumbent", "--no-graph"] + sys.argv[1:]
        bench.main()


if __name__ == "__main__":
    main()
