"""Dry run, on CPU, of the GPU tests that have not run on a GPU yet (tests/test_gpu_zz_*.py): their own host logic -- model
construction, state-dict round trips, FlatAdam's flat buffers, the pipeline, tolerances against the golden fixtures -- is
executed with the DEVICE ENTRY POINTS replaced by the oracle / plain torch (tests only; the product has no CPU path).  What a
dry run cannot show is the kernels' arithmetic: that is what the same tests check on the GPU box.

    python tests/dryrun_zz.py          (run by tests/test_static_checks.py::test_unverified_gpu_tests_dry_run_on_cpu in a
                                        subprocess: the stand-ins are patched into the modules process-wide)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import two_tower_recommender_model_b200 as tt
from two_tower_recommender_model_b200 import _native as N, two_tower as tw_mod, retrieval
from two_tower_recommender_model_b200.modules import embedding_modules, mlp
from two_tower_recommender_model_b200.sparse import jagged_tensor as jt
from test_reference_boundary import _OracleLookup, _oracle_linear_act

N.require_cuda = lambda t, name: None
N.stream_ptr = lambda dev: 0
embedding_modules.EbcLookup = _OracleLookup
mlp.linear_act = _oracle_linear_act
def bce(q, c, labels):
    logits = (q * c).sum(dim=1)
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels.float()), logits.detach()
tw_mod.dot_bce_loss = bce
def from_id_columns(keys, ids, num_embeddings, row_range=None):
    F, B = ids.shape
    ne = torch.as_tensor(num_embeddings).tolist()
    vals, lens = [], []
    for f in range(F):
        col = ids[f]
        keep = col != 0
        vals.append((col[keep] % ne[f] + ne[f]) % ne[f])
        lens.append(keep.to(torch.int32))
    v = torch.cat(vals); pad = torch.zeros(F * B - v.numel(), dtype=torch.int64)
    return jt.KeyedJaggedTensor(keys=list(keys), values=torch.cat([v, pad]), lengths=torch.cat(lens))
jt.KeyedJaggedTensor.from_id_columns = staticmethod(from_id_columns)
live = []
_orig_init = tt.FlatAdam.__init__
def init(self, *a, **k):
    _orig_init(self, *a, **k); live.append(self)
tt.FlatAdam.__init__ = init
def fake_call(name, *args):
    assert name == "tt_adam_flat_devstep", name
    p, g, m, v, n, lr, b1, b2, eps, step_ptr, stream = args
    o = next(x for x in live if x.flat_param.data_ptr() == p)
    o.step_dev += 1; t = float(o.step_dev)
    o.exp_avg.mul_(b1).add_(o.flat_grad, alpha=1 - b1)
    o.exp_avg_sq.mul_(b2).addcmul_(o.flat_grad, o.flat_grad, value=1 - b2)
    o.flat_param.addcdiv_(o.exp_avg / (1 - b1 ** t), (o.exp_avg_sq / (1 - b2 ** t)).sqrt() + eps, value=-lr)
N.call = fake_call
def topk(q, items, k, item_index_base=0, precision="fp32", items_bf16=None):
    s, i = oracle.exact_topk(q, items, k); return s, i + item_index_base
retrieval.score_topk = topk
import two_tower_recommender_model_b200.functional as Fn
Fn.cast_bf16 = lambda x, **kw: x.bfloat16()
cpu = torch.device("cpu")
import test_gpu_zz_checkpoint_resume as t1
t1.test_save_reload_next_step_identical(cpu); print("checkpoint_resume dry run ok")
import test_gpu_zz_corpus_sharded as t2
for prec in ("fp32", "bf16"):
    t2.test_corpus_sharded_index_single_rank(cpu, prec)
print("corpus_sharded dry run ok")
import test_gpu_zz_reference_golden as t3
t3.test_device_batch_construction_equals_the_reference_transform(cpu); print("golden batch construction ok")
t3.test_train_steps_equal_the_reference_bodies(cpu); print("golden train steps ok")
t3.test_corpus_embeddings_and_topk_equal_the_reference_functions(cpu); print("golden corpus ok")
t3.test_ray_tune_towers_equal_the_reference_class(cpu); print("golden raytune ok")


def _bucketize(lengths, offsets, values, num_rows, num_features, batch, world):
    from oracle.kjt import block_bucketize_vectorized
    nl, nv, unb = block_bucketize_vectorized(lengths, values, torch.as_tensor(num_rows).tolist(), world, batch)
    return nl, oracle.lengths_to_offsets(nl).to(torch.int32), nv, unb


Fn.block_bucketize = _bucketize
import test_gpu_zz_fbgemm_vector as t4  # noqa: E402
t4.test_block_bucketize_fbgemm_unit_test_vector(cpu); print("fbgemm vector ok")
