"""Column-wise sharding on two GPUs (NCCL exchange): the training worker of tests/test_gpu_multi.py in its ``column_wise``
mode -- every table split by columns over both ranks, per-shard row-wise Adagrad state, losses and gathered tables against
the oracle's column-block update.  Host logic of the same path: tests/test_sharding_gloo.py::test_column_wise_sharding_world2_gloo.
Kept in a file that sorts last: it has not run on GPUs yet (written after the round's GPU budget was spent)."""
import os

import pytest
import torch

from test_gpu_multi import _run_ranks, _worker

pytestmark = pytest.mark.gpu


def test_two_rank_column_wise_training_matches_oracle():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29890 + os.getpid() % 100
    _run_ranks(_worker, lambda r: (r, 2, port, "column_wise"))
