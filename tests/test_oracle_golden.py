"""CPU: pins the oracle against tests/golden/golden.json (hand-computed known answers and
outputs of stock torch ops; see tests/golden/make_golden.py) and against stock torch ops on
random inputs.  The oracle is then trusted as the checker for the CUDA path."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle.ebc import TableSpec
from oracle.kjt import block_bucketize_vectorized
from helpers import FBGEMM_BUCKETIZE_VECTOR, load_reference_golden, random_kjt

with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as f:
    G = json.load(f)
T = torch.tensor


def test_transform_kat():
    c = G["transform_kat"]
    v, l, y = oracle.transform_to_torchrec_batch(c["batch"], c["cat_cols"], c["emb_counts"])
    assert v.tolist() == c["values"] and l.tolist() == c["lengths"] and y.tolist() == c["labels"]
    assert v.dtype == torch.int64 and l.dtype == torch.int32 and y.dtype == torch.int32
    assert oracle.lengths_to_offsets(l).tolist() == c["offsets"]


def test_transform_negative_and_modulo():
    v, l, _ = oracle.transform_to_torchrec_batch({"a": [-3, 7, 0, 5], "label": [0] * 4}, ["a"], [5])
    assert v.tolist() == [2, 2, 0] and l.tolist() == [1, 1, 0, 1]  # Python modulo; 5 % 5 == 0 stays a length-1 bag


def test_transform_row_shard_kat():
    """Hand-computed: rows=10, W=3 -> block 4: rank 0 holds rows 0-3, rank 1 rows 4-7, rank 2 rows 8-9.
    ids [13, 0, 4, 9, 20] -> rows [3, -, 4, 9, 0]."""
    b = {"a": [13, 0, 4, 9, 20], "label": [0] * 5}
    want = {0: ([3, 0], [1, 0, 0, 0, 1]), 1: ([0], [0, 0, 1, 0, 0]), 2: ([1], [0, 0, 0, 1, 0])}
    v_all, l_all, _ = oracle.transform_to_torchrec_batch(b, ["a"], [10])
    nl, nv, _ = oracle.block_bucketize_sparse_features(l_all, v_all, [10], 3, 5)
    off = oracle.lengths_to_offsets(nl).tolist()
    for r in range(3):
        v, l = oracle.transform_row_shard(b, ["a"], [10], 3, r)
        assert (v.tolist(), l.tolist()) == want[r]
        assert l.tolist() == nl[r * 5:(r + 1) * 5].tolist() and v.tolist() == nv[off[r * 5]:off[(r + 1) * 5]].tolist()


def test_permute_kat():
    c = G["permute_kat"]
    ol, ov, _ = oracle.permute_2d_sparse_data(c["permute"], T(c["lengths"], dtype=torch.int32).view(c["T"], c["B"]), T(c["values"]))
    assert ol.reshape(-1).tolist() == c["out_lengths"] and ov.tolist() == c["out_values"]


def test_bucketize_kat():
    c = G["bucketize_kat"]
    for fn in (oracle.block_bucketize_sparse_features, block_bucketize_vectorized):
        nl, nv, unb = fn(T(c["lengths"], dtype=torch.int32), T(c["values"]), c["rows"], c["W"], c["B"])
        assert nl.tolist() == c["new_lengths"] and nv.tolist() == c["new_values"] and unb.tolist() == c["unbucketize"]


def test_bucketize_out_of_range_ids_kat():
    """fbgemm's fallback, hand-computed: rows=10, W=3 -> block 4, block*W = 12.  id 11 is still in the block range
    (bucket 2, local 3); id 12 and 100 are past it: bucket id % 3, local id // 3; id -1 is read as 2^64-1:
    bucket (2^64-1) % 3 = 0, local (2^64-1) // 3 = 6148914691236517205."""
    v = T([11, 12, 100, -1])
    l = T([1, 1, 1, 1], dtype=torch.int32)
    for fn in (oracle.block_bucketize_sparse_features, block_bucketize_vectorized):
        nl, nv, unb = fn(l, v, [10], 3, 4)
        #            w=0: b0 b1 b2 b3 | w=1          | w=2
        assert nl.tolist() == [0, 1, 0, 1, 0, 0, 1, 0, 1, 0, 0, 0]
        assert nv.tolist() == [4, 6148914691236517205, 33, 3]
        assert unb.tolist() == [3, 0, 2, 1]


def test_bucketize_fbgemm_unit_test_vector():
    """The oracle's block_bucketize against the vector fbgemm's own test suite holds (tests/helpers.py: FBGEMM_BUCKETIZE_VECTOR).  By hand:
    bags f0: [] [3 4]; f1: [15] [11 28 29]; f2: [1 10] [11 12 13]; f3: [11 22 20] [20].  bucket = id // block, local = id % block:
    f0 (5): 3, 4 -> bucket 0; f1 (15): 15 -> (1, 0), 11 -> (0, 11), 28 -> (1, 13), 29 -> (1, 14); f2 (10): 1 -> (0, 1), 10 -> (1, 0),
    11 12 13 -> (1, 1 2 3); f3 (20): 11 -> (0, 11), 22 -> (1, 2), 20 -> (1, 0), 20 -> (1, 0).  Output order [bucket][feature][sample]."""
    c = FBGEMM_BUCKETIZE_VECTOR
    rows = [b * c["my_size"] for b in c["block_sizes"]]           # the oracle derives block = ceil(rows / W)
    for fn in (oracle.block_bucketize_sparse_features, block_bucketize_vectorized):
        nl, nv, unb = fn(T(c["lengths"], dtype=torch.int32), T(c["indices"]), rows, c["my_size"], c["B"])
        assert nl.tolist() == c["new_lengths"] and nv.tolist() == c["new_indices"] and unb.tolist() == c["unbucketize_permute"]


@pytest.mark.parametrize("seed", range(5))
def test_bucketize_loop_equals_vectorized(seed):
    F_, B, W = 3, 17, 4
    rows = [50, 7, 1000]
    v, l = random_kjt(["a", "b", "c"], rows, B, 6, seed)
    a = oracle.block_bucketize_sparse_features(l, v, rows, W, B)
    b = block_bucketize_vectorized(l, v, rows, W, B)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_rowwise_adagrad_kat():
    c = G["rowwise_adagrad_kat"]
    w = T(c["weights"]); s = torch.zeros(2)
    g = torch.zeros(2, 2); g.index_add_(0, T(c["ids"]), T(c["grads"]))
    oracle.rowwise_adagrad_dense(w, s, g, lr=c["lr"], eps=c["eps"])
    torch.testing.assert_close(w, T(c["weights_after"])); torch.testing.assert_close(s, T(c["sum_after"]))
    w2 = T(c["weights"]); s2 = torch.zeros(2)
    oracle.rowwise_adagrad_sparse(w2, s2, T([0]), g[:1], lr=c["lr"], eps=c["eps"])
    torch.testing.assert_close(w2, w); torch.testing.assert_close(s2, s)


def test_rowwise_adam_kat():
    c = G["rowwise_adam_kat"]
    w = T(c["weights"]); m = torch.zeros(2, 2); v = torch.zeros(2)
    g = torch.zeros(2, 2); g.index_add_(0, T(c["ids"]), T(c["grads"]))
    oracle.rowwise_adam_sparse(w, m, v, T([0]), g[:1], c["step"], lr=c["lr"], beta1=c["beta1"], beta2=c["beta2"], eps=c["eps"])
    torch.testing.assert_close(w, T(c["weights_after"])); torch.testing.assert_close(m, T(c["m_after"]))
    torch.testing.assert_close(v, T(c["v_after"]))


def test_topk_ties_kat():
    c = G["topk_ties_kat"]
    s, i = oracle.exact_topk(T(c["queries"]), T(c["items"]), c["k"])
    assert i.tolist() == c["indices"] and s.tolist() == c["scores"]


def _specs(c):
    return [TableSpec(t["name"], t["rows"], t["dim"], t["features"], t["pooling"]) for t in c["tables"]]


def test_ebc_forward_and_backward_vs_torch_fixture():
    c, cb = G["ebc_forward_torch"], G["ebc_backward_torch"]
    specs = _specs(c)
    w = [T(x) for x in c["weights"]]
    v, l = T(c["values"]), T(c["lengths"], dtype=torch.int32)
    for fn in (oracle.ebc_forward, oracle.ebc_forward_torch):
        torch.testing.assert_close(fn(specs, w, c["keys"], v, l), T(c["pooled"]), rtol=1e-6, atol=1e-7)
    grads = oracle.ebc_dense_grads(specs, c["keys"], v, l, T(cb["grad_out"]))
    for g, want in zip(grads, cb["grads"]):
        torch.testing.assert_close(g, T(want), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("seed", range(4))
def test_ebc_forward_spec_equals_embedding_bag(seed):
    specs = [TableSpec("t0", 40, 8, ["a"], "sum"), TableSpec("t1", 30, 8, ["b", "c"], "mean")]
    w = [torch.randn(40, 8), torch.randn(30, 8)]
    keys = ["c", "a", "b"]
    v, l = random_kjt(keys, [30, 40, 30], 23, 5, seed)
    torch.testing.assert_close(oracle.ebc_forward(specs, w, keys, v, l), oracle.ebc_forward_torch(specs, w, keys, v, l),
                               rtol=1e-6, atol=1e-6)


def test_mlp_bce_softmax_adam_vs_torch_fixture():
    c = G["mlp_torch"]
    y = oracle.mlp_forward(T(c["x"]), [(T(w), T(b)) for w, b in c["layers"]])
    torch.testing.assert_close(y, T(c["y"]), rtol=1e-6, atol=1e-7)
    c = G["bce_torch"]
    loss, logits = oracle.dot_bce_loss(T(c["q"]), T(c["c"]), T(c["labels"], dtype=torch.int32))
    assert abs(float(loss) - c["loss"]) < 1e-6
    torch.testing.assert_close(logits, T(c["logits"]), rtol=1e-6, atol=1e-7)
    c = G["softmax_torch"]
    loss, _ = oracle.in_batch_softmax_loss(T(c["q"]), T(c["c"]), c["temperature"])
    assert abs(float(loss) - c["loss"]) < 1e-5
    c = G["adam_torch"]
    p = T(c["p0"]); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step, g in enumerate(c["grads"], 1):
        oracle.adam_step(p, T(g), m, v, step, lr=c["lr"])
    torch.testing.assert_close(p, T(c["p3"]), rtol=1e-6, atol=1e-7)


def test_oracle_two_tower_state_dict_names_and_step():
    specs = [TableSpec("t_user_id", 20, 8, ["user_id"]), TableSpec("t_product_id", 30, 8, ["product_id"])]
    m = oracle.OracleTwoTower(specs, [16, 8], seed=0)
    sd = m.torchrec_state_dict()
    assert set(sd) == {"ebc.embedding_bags.t_user_id.weight", "ebc.embedding_bags.t_product_id.weight",
                       "query_proj._mlp.0._linear.weight", "query_proj._mlp.0._linear.bias",
                       "query_proj._mlp.1._linear.weight", "query_proj._mlp.1._linear.bias",
                       "candidate_proj._mlp.0._linear.weight", "candidate_proj._mlp.0._linear.bias",
                       "candidate_proj._mlp.1._linear.weight", "candidate_proj._mlp.1._linear.bias"}
    v, l, y = oracle.transform_to_torchrec_batch({"user_id": [1, 2, 0, 4], "product_id": [3, 3, 9, 0], "label": [1, 0, 1, 0]},
                                                 ["user_id", "product_id"], [20, 30])
    before = sd["ebc.embedding_bags.t_product_id.weight"]
    loss, logits = m.train_step(["user_id", "product_id"], v, l, y)
    after = m.torchrec_state_dict()["ebc.embedding_bags.t_product_id.weight"]
    touched = (after != before).any(dim=1).nonzero().flatten().tolist()
    assert set(touched) <= {3, 9} and logits.shape == (4,) and loss.ndim == 0


def test_retrieval_metrics_known_answer():
    m = oracle.retrieval_metrics([[1, 2, 3, 4]], [[2, 9]], 4)
    import math
    assert abs(m["precision_at_4"] - 0.25) < 1e-9 and abs(m["recall_at_4"] - 0.5) < 1e-9
    assert abs(m["ndcg_at_4"] - (1 / math.log2(3)) / (1 + 1 / math.log2(3))) < 1e-9


def test_chunked_in_batch_softmax_equals_plain():
    """The row-chunked form the CPU baseline uses at B = 65 536 is the same function (loss and both gradients)."""
    g = torch.Generator().manual_seed(3)
    q = torch.randn(300, 16, generator=g, requires_grad=True)
    c = torch.randn(300, 16, generator=g, requires_grad=True)
    l0, d0 = oracle.in_batch_softmax_loss(q, c, 0.7)
    (l0 * 1.5).backward()
    gq, gc = q.grad.clone(), c.grad.clone()
    q.grad = c.grad = None
    l1, d1 = oracle.in_batch_softmax_loss_chunked(q, c, 0.7, chunk=64)
    (l1 * 1.5).backward()
    torch.testing.assert_close(l1, l0, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(d1, d0, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(q.grad, gq, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(c.grad, gc, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("seed", range(3))
def test_gathered_range_shard_is_one_bucket_of_block_bucketize(seed):
    """The sync-free input dist's filter, restated in the oracle, against the oracle's block_bucketize on the
    concatenated batch: rank w's shard == bucket w (lengths and values), for every w."""
    W, B, F = 3, 11, 2
    rows = [40, 9]
    per_rank = [random_kjt(["a", "b"], rows, B, 5, seed * 10 + r) for r in range(W)]
    # concatenated key-major batch: feature f = rank 0's bags, then rank 1's, ...
    cat_len = torch.cat([per_rank[r][1][f * B:(f + 1) * B] for f in range(F) for r in range(W)])
    offs = [oracle.lengths_to_offsets(l).tolist() for _, l in per_rank]
    cat_val = torch.cat([per_rank[r][0][offs[r][f * B]:offs[r][(f + 1) * B]] for f in range(F) for r in range(W)])
    nl, nv, _ = oracle.block_bucketize_sparse_features(cat_len, cat_val, rows, W, W * B)
    noff = oracle.lengths_to_offsets(nl).tolist()
    for w in range(W):
        block = [-(-r // W) for r in rows]
        lo = [w * b for b in block]
        hi = [min((w + 1) * b, r) for b, r in zip(block, rows)]
        v, l = oracle.gathered_range_shard([p[0] for p in per_rank], [p[1] for p in per_rank], lo, hi, B)
        n = F * W * B
        assert l.tolist() == nl[w * n:(w + 1) * n].tolist()
        assert v.tolist() == nv[noff[w * n]:noff[(w + 1) * n]].tolist()


def test_gathered_range_shard_kat():
    """Hand-computed: W = 2 ranks, B = 2, one feature, rows 10, this rank holds rows [5, 10).
    rank 0 bags: [7, 1], [9]; rank 1 bags: [], [5, 4, 6]  ->  global bags [7-5], [9-5], [], [5-5, 6-5]."""
    v, l = oracle.gathered_range_shard([T([7, 1, 9]), T([5, 4, 6])], [T([2, 1], dtype=torch.int32), T([0, 3], dtype=torch.int32)], [5], [10], 2)
    assert v.tolist() == [2, 4, 0, 1] and l.tolist() == [1, 1, 0, 2]


# ---- fixtures made by EXECUTING the reference's own code (tests/golden/make_reference_golden.py) -------------------------
def test_oracle_transform_equals_the_reference_transform_output():
    """oracle.transform_to_torchrec_batch against what the reference's own loop (utils/model_training.py:43-69) produced for
    the same raw columns: ids >= rows (modulo), id 0 (empty bag), bit-exact."""
    G = load_reference_golden()
    cat = ["user_id", "product_id"]
    n_zero = 0
    for i in range(G["steps"] + 2):
        raw = {c: G["z"][f"raw{i}_{c}"].tolist() for c in cat}
        raw["label"] = G["z"][f"batch{i}_labels"].tolist()
        v, l, y = oracle.transform_to_torchrec_batch(raw, cat, G["emb"])
        assert torch.equal(v, G["T"](f"batch{i}_values")) and torch.equal(l, G["T"](f"batch{i}_lengths")) and torch.equal(y, G["T"](f"batch{i}_labels"))
        n_zero += int((l == 0).sum())
    assert n_zero > 0                      # the fixture does exercise empty bags


def test_oracle_train_steps_equal_the_reference_bodies_on_stock_torch():
    """The oracle's whole train step against the reference's TwoTower / TwoTowerTrainTask / train() / evaluate() bodies run
    on stock torch (nn.EmbeddingBag, relu(nn.Linear), BCEWithLogitsLoss, row-wise Adagrad in the backward, Adam): loss and
    logits of every step, final weights, Adagrad accumulators, evaluate()'s average loss.  fp32 on both sides, same ops:
    rtol 1e-5 (summation order inside torch only)."""
    G = load_reference_golden()
    cat = ["user_id", "product_id"]
    specs = [TableSpec(f"t_{c}", G["emb"][i], G["dim"], [c]) for i, c in enumerate(cat)]
    orc = oracle.OracleTwoTower(specs, G["layers"], loss="bce", sparse_lr=G["lr"], dense_lr=G["lr"], seed=0)
    orc.load_torchrec_state_dict(G["init"])
    assert set(orc.torchrec_state_dict()) == set(G["init"])          # TorchRec's key names
    for i in range(G["steps"]):
        loss, logits = orc.train_step(cat, G["T"](f"batch{i}_values"), G["T"](f"batch{i}_lengths"), G["T"](f"batch{i}_labels"))
        torch.testing.assert_close(loss, G["T"](f"step{i}_loss"), rtol=1e-5, atol=1e-7, msg=lambda m: f"step {i} loss: {m}")
        torch.testing.assert_close(logits, G["T"](f"step{i}_logits"), rtol=1e-5, atol=1e-6, msg=lambda m: f"step {i} logits: {m}")
    got = orc.torchrec_state_dict()
    moved = 0.0
    for k, want in G["final"].items():
        torch.testing.assert_close(got[k], want, rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
        moved = max(moved, float((want - G["init"][k]).abs().max()))
    assert moved > 5e-3                    # the steps did move the weights: the comparison is not vacuous
    for c in cat:
        torch.testing.assert_close(orc.sparse_state[f"t_{c}"]["sum"], G["T"](f"sum.t_{c}"), rtol=1e-5, atol=1e-12)
    total = 0.0
    for j in range(2):
        i = G["steps"] + j
        q, cnd = orc.forward(cat, G["T"](f"batch{i}_values"), G["T"](f"batch{i}_lengths"))
        loss, logits = orc.loss(q, cnd, G["T"](f"batch{i}_labels"))
        torch.testing.assert_close(loss.detach(), G["T"](f"eval{j}_loss"), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(logits.detach(), G["T"](f"eval{j}_logits"), rtol=1e-5, atol=1e-6)
        total += float(loss)
    assert abs(total / (2 * G["B"]) - float(G["z"]["eval_average_loss"])) < 1e-8   # U:246 divides by the SAMPLE count


def test_oracle_corpus_embeddings_and_topk_equal_the_reference_functions():
    """oracle towers / oracle.exact_topk against the item and user embeddings the reference's own create_keyed_jagged_tensor
    + process_embeddings (03_model_training.py:1056-1122) computed from the trained model, and their top-100 by torch.sort."""
    G = load_reference_golden()
    cat = ["user_id", "product_id"]
    specs = [TableSpec(f"t_{c}", G["emb"][i], G["dim"], [c]) for i, c in enumerate(cat)]
    orc = oracle.OracleTwoTower(specs, G["layers"], loss="bce", seed=0)
    orc.load_torchrec_state_dict(G["final"])
    embs = {}
    for key, n in (("product_id", G["emb"][1]), ("user_id", G["emb"][0])):
        v, l = G["T"](f"corpus_{key}_values"), G["T"](f"corpus_{key}_lengths")
        assert v.tolist() == list(range(n)) and l.tolist() == ([0] * n + [1] * n if key == "product_id" else [1] * n + [0] * n)
        with torch.no_grad():
            q, c = orc.forward(cat, v, l)
        embs[key] = c if key == "product_id" else q
        torch.testing.assert_close(embs[key], G["T"](f"corpus_{key}_embeddings"), rtol=1e-5, atol=1e-6)
    ws, wi = oracle.exact_topk(embs["user_id"], embs["product_id"], 100)
    torch.testing.assert_close(ws, G["T"]("top100_scores"), rtol=1e-5, atol=1e-6)
    want = G["T"]("top100_ids")
    # ids: equal wherever the neighbouring scores are further apart than fp32 summation noise
    gap = (G["T"]("top100_scores")[:, :-1] - G["T"]("top100_scores")[:, 1:]).abs()
    clear = torch.ones_like(want, dtype=torch.bool)
    clear[:, :-1] &= gap > 1e-6
    clear[:, 1:] &= gap > 1e-6
    clear[:, -1] = False                   # the 100th may swap with the 101st
    assert torch.equal(wi[clear], want[clear]) and float(clear.float().mean()) > 0.9


def test_oracle_ray_tune_towers_equal_the_reference_class():
    """Several features per tower, one layer stack per tower, dense features concatenated to the tower inputs: the oracle
    against the reference's own Ray-Tune TwoTower class (ray_tune_optuna_tuning_alex_test.py:181-306) run on stock torch --
    embeddings, logits, BCE loss and the gradient of every parameter."""
    from helpers import load_raytune_golden
    G = load_raytune_golden()
    keys = list(G["dims"])
    specs = [TableSpec(f"t_{k}", G["rows"][k], G["dims"][k], [k]) for k in keys]
    orc = oracle.OracleTwoTower(specs, G["layers"], loss="bce", query_features=G["feats_u"], candidate_features=G["feats_i"],
                                dense_index=G["dense_index"], dense_dim=G["dense_dim"], seed=0)
    assert set(orc.torchrec_state_dict()) == set(G["weights"])
    orc.load_torchrec_state_dict(G["weights"])
    q, c = orc.forward(keys, G["values"], G["lengths"], G["dense"])
    loss, logits = orc.loss(q, c, G["labels"])
    torch.testing.assert_close(q.detach(), G["q"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(c.detach(), G["c"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits.detach(), G["logits"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss.detach(), G["loss"], rtol=1e-6, atol=1e-7)
    loss.backward()
    got = {f"ebc.embedding_bags.{s.name}.weight": orc.embedding_bags[s.name].weight.grad for s in specs}
    for tower, mods in (("query_proj", orc.query_proj), ("candidate_proj", orc.candidate_proj)):
        for i, lin in enumerate(mods):
            got[f"{tower}._mlp.{i}._linear.weight"], got[f"{tower}._mlp.{i}._linear.bias"] = lin.weight.grad, lin.bias.grad
    assert set(got) == set(G["grads"])
    for k, want in G["grads"].items():
        torch.testing.assert_close(got[k], want, rtol=1e-5, atol=1e-8, msg=lambda m: f"grad of {k}: {m}")
        assert float(want.abs().max()) > 0


def test_oracle_column_blocks_known_answer():
    """Column-wise shards keep one row-wise accumulator PER SHARD.  With one column per shard the mean over the shard's
    columns is g^2 itself, so the first Adagrad step moves every touched element by exactly lr (w -= lr * g / |g|), whereas
    the unsharded table moves a row by lr * g / rms(g) -- the known answer that tells the two apart."""
    cat = ["user_id", "product_id"]
    specs = [TableSpec("t_user_id", 11, 2, ["user_id"]), TableSpec("t_product_id", 7, 2, ["product_id"])]
    lr = 0.1
    a = oracle.OracleTwoTower(specs, [32, 16], loss="bce", sparse_lr=lr, dense_lr=0.0, seed=4, dense_optimizer="sgd",
                              column_blocks={"t_user_id": 2})
    b = oracle.OracleTwoTower(specs, [32, 16], loss="bce", sparse_lr=lr, dense_lr=0.0, seed=4, dense_optimizer="sgd")
    w0 = a.embedding_bags["t_user_id"].weight.detach().clone()
    v = torch.tensor([1, 3, 3, 5, 2, 4, 6, 0])
    l = torch.ones(8, dtype=torch.int32)
    y = torch.tensor([1, 0, 1, 0], dtype=torch.int32)
    a.train_step(cat, v, l, y)
    b.train_step(cat, v, l, y)
    da = a.embedding_bags["t_user_id"].weight.detach() - w0
    db = b.embedding_bags["t_user_id"].weight.detach() - w0
    touched = torch.zeros(11, dtype=torch.bool)
    touched[[1, 3, 5]] = True
    assert (da[~touched] == 0).all() and (db[~touched] == 0).all()
    moved = da[touched].abs()
    torch.testing.assert_close(moved[moved > 0], torch.full_like(moved[moved > 0], lr), rtol=1e-5, atol=1e-7)
    assert float(moved.max()) > 0
    # the unsharded table: one accumulator per row, the row moves by lr * g / rms(g): |step|^2 per row sums to D * lr^2
    torch.testing.assert_close(db[touched].pow(2).sum(dim=1), torch.full((3,), 2 * lr * lr), rtol=1e-4, atol=1e-8)
    # the table without column blocks is updated identically by both
    torch.testing.assert_close(a.embedding_bags["t_product_id"].weight, b.embedding_bags["t_product_id"].weight)


def test_oracle_ndcg_against_sklearn():
    """mlflow's retriever ``ndcg_at_k`` (04_evaluate_retrieval.py:202-226) is sklearn's ``ndcg_score`` over the retrieved ids
    (descending scores, relevance 1 if a target) plus the targets that were not retrieved (relevance 1, lowest score), cut
    at the number of ids retrieved.  The oracle's closed form against that stock implementation on random rows."""
    sklearn_metrics = pytest.importorskip("sklearn.metrics")
    g = torch.Generator().manual_seed(123)
    checked = 0
    for _ in range(200):
        n_pred = int(torch.randint(1, 12, (1,), generator=g))
        k = int(torch.randint(1, 12, (1,), generator=g))
        pred = torch.randperm(25, generator=g)[:n_pred].tolist()
        tgt = torch.randperm(25, generator=g)[:int(torch.randint(1, 8, (1,), generator=g))].tolist()
        retrieved = pred[:k]
        docs = retrieved + [t for t in tgt if t not in retrieved]
        if len(docs) < 2:
            continue                                   # sklearn refuses a single document
        y_true = [1.0 if d in tgt else 0.0 for d in docs]
        y_score = [float(len(docs) - i) for i in range(len(retrieved))] + [0.0] * (len(docs) - len(retrieved))
        want = sklearn_metrics.ndcg_score([y_true], [y_score], k=len(retrieved), ignore_ties=True)
        got = oracle.retrieval_metrics([pred], [tgt], k)[f"ndcg_at_{k}"]
        assert abs(got - want) < 1e-9, (pred, tgt, k, got, want)
        checked += 1
    assert checked > 150


@pytest.mark.parametrize("eps", [1e-10, 1e-8])
def test_rowwise_adagrad_at_dim_1_is_torch_adagrad(eps):
    """A pin that stock torch holds: with one column per row the row-wise mean of g^2 IS g^2, so row-wise Adagrad
    (oracle/ebc.py, restating torchrec.optim.RowWiseAdagrad) must equal ``torch.optim.Adagrad`` step for step --
    accumulate first, eps OUTSIDE the square root, lr_decay = 0.  Both eps values SURVEY 8(c) item 5 names."""
    g0 = torch.Generator().manual_seed(3)
    w = torch.randn(50, 1, generator=g0)
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adagrad([p], lr=0.05, eps=eps, initial_accumulator_value=0.0, lr_decay=0.0)
    s = torch.zeros(50)
    w_sparse, s_sparse = w.clone(), torch.zeros(50)
    for step in range(5):
        grad = torch.randn(50, 1, generator=g0)
        grad[torch.rand(50, generator=g0) < 0.5] = 0.0          # rows without an occurrence this step
        p.grad = grad.clone()
        opt.step()
        oracle.rowwise_adagrad_dense(w, s, grad, lr=0.05, eps=eps)
        rows = grad[:, 0].nonzero().flatten()
        oracle.rowwise_adagrad_sparse(w_sparse, s_sparse, rows, grad[rows], lr=0.05, eps=eps)
        torch.testing.assert_close(w, p.detach(), rtol=1e-6, atol=1e-7, msg=lambda m: f"step {step}: {m}")
        torch.testing.assert_close(s, opt.state[p]["sum"].reshape(-1), rtol=1e-6, atol=1e-12)
        torch.testing.assert_close(w_sparse, w, rtol=1e-6, atol=1e-7)


def test_rowwise_adam_at_dim_1_is_torch_adam():
    """The same pin for the row-wise Adam extension (FBGEMM PARTIAL_ROWWISE_ADAM semantics): at one column per row, with
    every row touched at every step, it must equal ``torch.optim.Adam`` -- first moment, second moment, both bias
    corrections, eps added to sqrt(v_hat)."""
    g0 = torch.Generator().manual_seed(4)
    w = torch.randn(40, 1, generator=g0)
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([p], lr=0.02, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros(40, 1), torch.zeros(40)
    rows = torch.arange(40)
    for step in range(1, 6):
        grad = torch.randn(40, 1, generator=g0)
        p.grad = grad.clone()
        opt.step()
        oracle.rowwise_adam_sparse(w, m, v, rows, grad, step, lr=0.02, beta1=0.9, beta2=0.999, eps=1e-8)
        torch.testing.assert_close(w, p.detach(), rtol=1e-5, atol=1e-7, msg=lambda msg: f"step {step}: {msg}")
        # torch forms the first moment with lerp_ (m + (1 - b1) * (g - m)): same value, another rounding
        torch.testing.assert_close(m, opt.state[p]["exp_avg"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(v, opt.state[p]["exp_avg_sq"].reshape(-1), rtol=1e-5, atol=1e-10)
