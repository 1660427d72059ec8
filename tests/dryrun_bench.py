"""Dry run, on CPU, of bench.py's single-GPU flow -- ``time_block`` (eager variant, then the default one behind
``CudaGraphTrainStep``, whose "replay" here runs the step again) -> ``headline`` -> ``finish`` -- on a tiny configuration, with
the DEVICE ENTRY POINTS replaced by the oracle / plain torch and the CUDA runtime primitives (events, streams, graphs) by
wall-clock / no-op fakes (tests only; the product has no CPU path).  It executes the glue the CPU suite otherwise only sees through fakes of
``time_block``: model construction, the raw-batch path, the pipeline loop, the explanatory passes, the line assembly, every
side block.  The numbers it prints mean nothing; the line's SHAPE is what tests/test_static_checks.py checks.

    python tests/dryrun_bench.py       (run in a subprocess: the stand-ins are patched process-wide)
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import two_tower_recommender_model_b200 as tt  # noqa: E402
import two_tower_recommender_model_b200.functional as Fn  # noqa: E402
from two_tower_recommender_model_b200 import _native as N, retrieval, two_tower as tw_mod  # noqa: E402
from two_tower_recommender_model_b200.modules import embedding_modules, mlp  # noqa: E402
from two_tower_recommender_model_b200.sparse import jagged_tensor as jt  # noqa: E402
from test_reference_boundary import _OracleLookup, _oracle_linear_act  # noqa: E402

import bench  # noqa: E402

cpu = torch.device("cpu")

# ---- device entry points -> oracle / plain torch
N.require_cuda = lambda t, name: None
N.stream_ptr = lambda dev: 0
embedding_modules.EbcLookup = _OracleLookup
mlp.linear_act = _oracle_linear_act


def _softmax_loss(q, c, temperature=1.0, precision="fp32", negatives="local", pg=None):
    logits = q @ c.t() / temperature
    return torch.nn.functional.cross_entropy(logits, torch.arange(q.shape[0])), logits.diagonal().detach()


tw_mod.in_batch_softmax_loss = _softmax_loss


def _from_id_columns(keys, ids, num_embeddings, row_range=None):
    F, B = ids.shape
    ne = torch.as_tensor(num_embeddings).tolist()
    vals, lens = [], []
    for f in range(F):
        keep = ids[f] != 0
        vals.append(ids[f][keep] % ne[f])
        lens.append(keep.to(torch.int32))
    v = torch.cat(vals)
    return jt.KeyedJaggedTensor(keys=list(keys), values=torch.cat([v, torch.zeros(F * B - v.numel(), dtype=torch.int64)]), lengths=torch.cat(lens))


jt.KeyedJaggedTensor.from_id_columns = staticmethod(_from_id_columns)

_live = []
_flat_init = tt.FlatAdam.__init__


def _init(self, *a, **k):
    _flat_init(self, *a, **k)
    _live.append(self)


tt.FlatAdam.__init__ = _init
_calls = {}


def _fake_call(name, *args):
    _calls[name] = _calls.get(name, 0) + 1
    if name == "tt_adam_flat_devstep":
        p, g, m, v, n, lr, b1, b2, eps, step_ptr, stream = args
        o = next(x for x in _live if x.flat_param.data_ptr() == p)
        o.step_dev += 1
        t = float(o.step_dev)
        o.exp_avg.mul_(b1).add_(o.flat_grad, alpha=1 - b1)
        o.exp_avg_sq.mul_(b2).addcmul_(o.flat_grad, o.flat_grad, value=1 - b2)
        o.flat_param.addcdiv_(o.exp_avg / (1 - b1 ** t), (o.exp_avg_sq / (1 - b2 ** t)).sqrt() + eps, value=-lr)
    elif name != "tt_ebc_forward":            # the lookup-alone loop: launches only
        raise AssertionError(f"unexpected device call {name}")


N.call = _fake_call
N.timing_summary = lambda: {"tt_inbatch_softmax_forward_f32": {"ms": 1.0, "calls": 5}, "tt_inbatch_softmax_backward_f32": {"ms": 1.6, "calls": 5},
                            "tt_ebc_forward": {"ms": 0.03, "calls": 5}, "tt_ebc_backward_fused": {"ms": 0.1, "calls": 5}}
retrieval.score_topk = lambda q, items, k, item_index_base=0, precision="fp32", items_bf16=None: oracle.exact_topk(
    q, items if items is not None else items_bf16.float(), k)
Fn.cast_bf16 = lambda x, **kw: x.bfloat16()


# ---- CUDA runtime primitives -> wall clock / no-ops
class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class _Graph:
    """A "captured" step is replayed by running it again: same effect on the static buffers as a real replay."""
    owner = None

    def replay(self):
        _Graph.owner._out = _Graph.owner._step()


class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_graph_init = tt.CudaGraphTrainStep.__init__


def _gs_init(self, *a, **k):
    _graph_init(self, *a, **k)
    _Graph.owner = self


tt.CudaGraphTrainStep.__init__ = _gs_init
torch.cuda.Stream = _Stream
torch.cuda.current_stream = lambda device=None: _Stream()
torch.cuda.stream = _Ctx
torch.cuda.graph = _Ctx
torch.cuda.CUDAGraph = _Graph
torch.cuda.Event = _Event
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda: None
torch.Tensor.pin_memory = lambda self, *a, **k: self

# ---- tiny workload
tiny = dict(rows=[500, 400], dim=16, layers=[32, 16], batch=64, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001, precision="fp32")
bench.CFG1 = dict(rows=[300, 200], dim=16, layers=[32, 16], batch=32, loss="bce", sparse_lr=0.01, dense_lr=0.001)
_probe = bench.retrieval_probe
bench.retrieval_probe = lambda dev, n_items=3000, n_queries=64, d=64, k=100: _probe(dev, min(n_items, 3000), min(n_queries, 64), d, k)
args = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="ours", no_cpu_baseline=False, no_other_configs=True, exchange="peer",
                          parity_only=False, parity_graph=False, no_graph=True)
lib = N.load()
main = bench.time_block(tiny, tiny["batch"], cpu, 0, 1, 0, args, None, None, lib, with_kernels=True)
assert main["explain_error"] is None, main["explain_error"]
assert _calls.get("tt_ebc_forward") == 44 and _calls.get("tt_adam_flat_devstep", 0) >= 2 * (args.steps + args.warmup)
# the default variant: the step behind CudaGraphTrainStep (eager warm-up calls, "capture", replays), as the driver runs it
args.no_graph = False
_calls.clear()
main = bench.time_block(tiny, tiny["batch"], cpu, 0, 1, 0, args, None, None, lib, with_kernels=True)
assert main["explain_error"] is None and main["cuda_graph"] is True and _Graph.owner.captured, main
assert "CudaGraphTrainStep" in main["e2e_api"] and main["last_loss"] > 0
line = bench.headline(args, tiny, main, bench.peaks(), 1, tiny["batch"], "dry run", "strong", [])
bench.finish(args, tiny, cpu, 1, line)
print("bench dry run ok")
