"""Column-wise and data-parallel sharding on two GPUs (NCCL exchange): the training worker of tests/test_gpu_multi.py in its
``column_wise`` mode (every table split by columns over both ranks, per-shard row-wise Adagrad state) and in ``data_parallel``
mode (a replica per rank, dense gradient averaged by ``sync_dense_grads``, row-wise Adagrad applied to the replica) -- losses
of every step and the final tables against the oracle's two-rank step.  Host logic of the same paths:
tests/test_sharding_gloo.py::test_column_wise_sharding_world2_gloo / ::test_data_parallel_tables_world2_gloo.
Kept in a file that sorts last: these modes have not run on GPUs yet (written after the round's GPU budget was spent)."""
import os

import pytest
import torch

from test_gpu_multi import _run_ranks, _worker

pytestmark = pytest.mark.gpu

MODES = ["column_wise", "data_parallel", "data_parallel_dense"]      # *_dense: batches arrive as id columns (from_id_columns)


@pytest.mark.parametrize("sharding", MODES)
def test_two_rank_training_matches_oracle(sharding):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29890 + os.getpid() % 100 + MODES.index(sharding)
    _run_ranks(_worker, lambda r: (r, 2, port, sharding))
