"""Dry run, on CPU, of ``__graft_entry__.smoke()`` -- what the driver runs on the GPU box before the bench: two training steps
per loss through DistributedModelParallel + TrainPipelineSparseDist against the oracle, the bf16 leg (fused towers + in-batch
softmax), exact top-k.  The device entry points are replaced by tests/dryrun_standins.py, ``torch.device("cuda:0")`` resolves
to the CPU and the library's launch counter is a stub (nothing launches here), so what this checks is smoke()'s own flow and
its comparisons against the oracle's numbers, not the kernels.

    python tests/dryrun_smoke.py       (run in a subprocess: the stand-ins are patched process-wide)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dryrun_standins as S  # noqa: E402

S.install()
import oracle  # noqa: E402
import two_tower_recommender_model_b200.functional as Fn  # noqa: E402
from two_tower_recommender_model_b200 import _native as N  # noqa: E402

Fn.score_topk = lambda q, items, k, item_index_base=0, precision="fp32", items_bf16=None: oracle.exact_topk(q, items, k)

_real_device = torch.device


class _DeviceMeta(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, _real_device)


class _Device(metaclass=_DeviceMeta):
    """``torch.device`` whose CUDA devices are the CPU."""

    def __new__(cls, *a, **k):
        d = _real_device(*a, **k)
        return _real_device("cpu") if d.type == "cuda" else d


torch.device = _Device


class _Lib:
    """The loaded library with a launch counter that ticks (no kernel launches in a dry run)."""
    ticks = 0

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def tt_kernel_launch_count(self):
        _Lib.ticks += 1
        return _Lib.ticks


_lib = _Lib(N.load())
N.load = lambda: _lib

import __graft_entry__ as entry  # noqa: E402

entry.build = lambda: None          # the library is built by the suite's own build step; nothing to compile here
entry.smoke()
print("smoke dry run ok")
