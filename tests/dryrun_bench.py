"""Dry run, on CPU, of bench.py's single-GPU run: ``run_ours(args)`` itself -- the function the driver's launch reaches -- on a
tiny configuration (the default variant: the step behind ``CudaGraphTrainStep``, whose "replay" here runs the step again), and
``time_block`` once more in its eager variant.  The device entry points and the CUDA runtime primitives are replaced by
tests/dryrun_standins.py (tests only; the product has no CPU path).  It executes the glue the CPU suite otherwise only sees
through fakes of ``time_block``: model construction, the raw-batch path, the pipeline / graph-step loops, the explanatory
passes, the line assembly, every side block.  The numbers it prints mean nothing; the line's SHAPE is what
tests/test_static_checks.py checks.

    python tests/dryrun_bench.py       (run in a subprocess: the stand-ins are patched process-wide)
"""
import argparse

import torch

import dryrun_standins as S

S.install()
import bench  # noqa: E402
from two_tower_recommender_model_b200 import _native as N  # noqa: E402

tiny = dict(rows=[500, 400], dim=16, layers=[32, 16], batch=64, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001)
S.run_bench_on_cpu(bench, tiny)
args = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="ours", no_cpu_baseline=False, no_other_configs=True, exchange="peer",
                          parity_only=False, parity_graph=False, no_graph=True)
# the eager variant (--no-graph): TrainPipelineSparseDist drives the end-to-end loop
main = bench.time_block(dict(tiny, precision="fp32"), tiny["batch"], torch.device("cpu"), 0, 1, 0, args, None, None, N.load(), with_kernels=True)
assert main["explain_error"] is None and main["cuda_graph"] is False, main
assert S.calls.get("tt_ebc_forward") == 44 and S.calls.get("tt_adam_flat_devstep", 0) >= 2 * (args.steps + args.warmup)
assert "TrainPipelineSparseDist" in main["e2e_api"]
# the run as the driver launches it
args.no_graph = False
bench.run_ours(args)
assert S.Graph.owner.captured
print("bench dry run ok")
