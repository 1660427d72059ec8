"""The bucketize kernel on the vector fbgemm's own test suite holds (added without GPU time: the file sorts last so that a
first run of it cannot hide the tests that have run green on a GPU)."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_block_bucketize_fbgemm_unit_test_vector(cuda):
    """tt_kjt_block_bucketize on the vector fbgemm's own test suite holds for block_bucketize_sparse_features
    (tests/helpers.py: FBGEMM_BUCKETIZE_VECTOR, re-derived by hand in tests/test_oracle_golden.py): T = 4, B = 2, my_size = 2."""
    from two_tower_recommender_model_b200.functional import block_bucketize
    from helpers import FBGEMM_BUCKETIZE_VECTOR as c
    rows = [b * c["my_size"] for b in c["block_sizes"]]
    l = torch.tensor(c["lengths"], dtype=torch.int32)
    v = torch.tensor(c["indices"], dtype=torch.int64)
    off = oracle.lengths_to_offsets(l)
    nl, no, nv, unb = block_bucketize(l.to(cuda), off.to(cuda), v.to(cuda), torch.tensor(rows), 4, c["B"], c["my_size"])
    assert nl.cpu().tolist() == c["new_lengths"] and nv.cpu().tolist() == c["new_indices"]
    assert unb.cpu().tolist() == c["unbucketize_permute"]
    assert torch.equal(no.cpu(), oracle.lengths_to_offsets(torch.tensor(c["new_lengths"], dtype=torch.int32)))
