"""CorpusShardedIndex on one GPU (world size 1: no collective): the shard's first id reaches the kernel as
``item_index_base``, the merge keeps (descending score, ties -> lower id).  The world-2 exchange + merge is covered on CPU
with gloo (tests/test_sharding_gloo.py::test_corpus_sharded_retrieval_world2_gloo)."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_corpus_sharded_index_single_rank(cuda, precision):
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator().manual_seed(1)
    # entries on a 1/8 grid in [-1, 1]: every dot product is exact in fp32 and in bf16 operands, so indices are bit-exact
    q = torch.randint(-8, 9, (64, 32), generator=g).float() / 8
    items = torch.randint(-8, 9, (3000, 32), generator=g).float() / 8
    ws, wi = oracle.exact_topk(q, items, 100)
    index = tt.CorpusShardedIndex(items.to(cuda), first_id=5000, precision=precision)
    s, i = index.search(q.to(cuda), 100)
    assert torch.equal(i.cpu(), wi + 5000) and torch.equal(s.cpu(), ws)
