"""pytest plugin (opt-in: `python -m pytest tests -m gpu -p double_plugin`, run from tests/ or with PYTHONPATH=tests):
runs the `-m gpu` test files WITHOUT a GPU, on the kernel test double — `.cuda()` becomes the identity,
`torch.cuda.is_available()` answers True, every kernel is the torch-CPU restatement of tests/kernel_double.py.

Purpose: a regression net for HOST-side changes made where no GPU is available (the last part of round 2): the GPU
tests that exercise nets / trainers / samplers end to end then run against their goldens on the CPU.  Tests that check
a KERNEL against torch (tests/test_gpu_ops.py, test_gpu_production_shapes.py, the kernel halves of test_gpu_vid / _gan /
_jit / _widen_cut) compare the double with torch there — they say nothing and are expected to pass trivially or to hit
an op the double does not restate.  Never part of the default suites; never a substitute for the B200 run.
"""
import contextlib

import torch

import kernel_double as KD

_stack = contextlib.ExitStack()


def pytest_configure(config):
    from unittest import mock
    _stack.enter_context(KD.installed())
    ident = lambda self, *a, **k: self  # noqa: E731
    fresh = lambda self, *a, **k: self.clone()  # noqa: E731  (.cuda() makes a new tensor: `x.cuda().requires_grad_()` is a leaf)
    real_device = torch.device

    class _Device:   # torch.device("cuda[:i]") -> the CPU device; isinstance(x, torch.device) keeps working
        def __new__(cls, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                return real_device("cpu")
            return real_device(*a, **k)

        @classmethod
        def __instancecheck__(cls, obj):
            return isinstance(obj, real_device)

    class _Meta(type):
        def __instancecheck__(cls, obj):
            return isinstance(obj, real_device)

    _Dev = _Meta("device", (), {"__new__": staticmethod(lambda cls, *a, **k: real_device("cpu") if (
        a and isinstance(a[0], str) and a[0].startswith("cuda")) else real_device(*a, **k))})
    _stack.enter_context(mock.patch("torch.device", _Dev))
    for target, value in (("torch.cuda.is_available", lambda: True), ("torch.cuda.synchronize", lambda *a, **k: None),
                          ("torch.cuda.set_device", lambda *a, **k: None), ("torch.cuda.device_count", lambda: 1),
                          ("torch.cuda.manual_seed", lambda *a, **k: None),
                          ("torch.cuda.is_current_stream_capturing", lambda: False)):
        _stack.enter_context(mock.patch(target, value))
    _stack.enter_context(mock.patch.object(torch.Tensor, "cuda", fresh))
    _stack.enter_context(mock.patch.object(torch.nn.Module, "cuda", ident))
    real_to = torch.Tensor.to

    def to(self, *a, **k):   # .to("cuda") / .to(device="cuda:0") stay on the CPU
        a = tuple(x for x in a if not (isinstance(x, (str, real_device)) and "cuda" in str(x)))
        if "device" in k and "cuda" in str(k["device"]):
            k = {kk: v for kk, v in k.items() if kk != "device"}
        return real_to(self, *a, **k) if (a or k) else self
    _stack.enter_context(mock.patch.object(torch.Tensor, "to", to))
    real_mto = torch.nn.Module.to

    def mto(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, (str, real_device)) and "cuda" in str(x)))
        if "device" in k and "cuda" in str(k["device"]):
            k = {kk: v for kk, v in k.items() if kk != "device"}
        return real_mto(self, *a, **k) if (a or k) else self
    _stack.enter_context(mock.patch.object(torch.nn.Module, "to", mto))


def pytest_unconfigure(config):
    _stack.close()
