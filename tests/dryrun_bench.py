"""Dry run, on CPU, of bench.py's single-GPU flow -- ``time_block`` (eager variant, then the default one behind
``CudaGraphTrainStep``, whose "replay" here runs the step again) -> ``headline`` -> ``finish`` -- on a tiny configuration, with
the device entry points and the CUDA runtime primitives replaced by tests/dryrun_standins.py (tests only; the product has no
CPU path).  It executes the glue the CPU suite otherwise only sees through fakes of ``time_block``: model construction, the
raw-batch path, the pipeline loop, the explanatory passes, the line assembly, every side block.  The numbers it prints mean
nothing; the line's SHAPE is what tests/test_static_checks.py checks.

    python tests/dryrun_bench.py       (run in a subprocess: the stand-ins are patched process-wide)
"""
import argparse

import torch

import dryrun_standins as S

S.install()
import bench  # noqa: E402
from two_tower_recommender_model_b200 import _native as N  # noqa: E402

cpu = torch.device("cpu")
tiny = dict(rows=[500, 400], dim=16, layers=[32, 16], batch=64, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001, precision="fp32")
bench.CFG1 = dict(rows=[300, 200], dim=16, layers=[32, 16], batch=32, loss="bce", sparse_lr=0.01, dense_lr=0.001)
_probe = bench.retrieval_probe
bench.retrieval_probe = lambda dev, n_items=3000, n_queries=64, d=64, k=100: _probe(dev, min(n_items, 3000), min(n_queries, 64), d, k)
args = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="ours", no_cpu_baseline=False, no_other_configs=True, exchange="peer",
                          parity_only=False, parity_graph=False, no_graph=True)
lib = N.load()
main = bench.time_block(tiny, tiny["batch"], cpu, 0, 1, 0, args, None, None, lib, with_kernels=True)
assert main["explain_error"] is None, main["explain_error"]
assert S.calls.get("tt_ebc_forward") == 44 and S.calls.get("tt_adam_flat_devstep", 0) >= 2 * (args.steps + args.warmup)
# the default variant: the step behind CudaGraphTrainStep (eager warm-up calls, "capture", replays), as the driver runs it
args.no_graph = False
S.calls.clear()
main = bench.time_block(tiny, tiny["batch"], cpu, 0, 1, 0, args, None, None, lib, with_kernels=True)
assert main["explain_error"] is None and main["cuda_graph"] is True and S.Graph.owner.captured, main
assert "CudaGraphTrainStep" in main["e2e_api"] and main["last_loss"] > 0
line = bench.headline(args, tiny, main, bench.peaks(), 1, tiny["batch"], "dry run", "strong", [])
bench.finish(args, tiny, cpu, 1, line)
print("bench dry run ok")
