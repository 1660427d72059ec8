"""End-to-end: N train steps of the CUDA two-tower vs the CPU oracle on the same
seeded batches, through the reference-facing API (EmbeddingBagCollection on meta,
apply_optimizer_in_backward, DistributedModelParallel, KeyedOptimizerWrapper,
TrainPipelineSparseDist.progress).  fp32: loss rtol 1e-4, weights rtol 1e-4 atol 1e-5."""
import itertools
import os

import pytest
import torch
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward

import oracle
from oracle.ebc import TableSpec

pytestmark = pytest.mark.gpu

CAT = ["user_id", "product_id"]


def make_batches(n, B, emb, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        out.append({"user_id": torch.randint(0, emb[0] * 2, (B,), generator=g).tolist(),
                    "product_id": torch.randint(0, emb[1] * 2, (B,), generator=g).tolist(),
                    "label": torch.randint(0, 2, (B,), generator=g).tolist()})
    return out


def build_models(cuda, emb, dim, layers, loss, sparse_opt, lr, dense="adam"):
    import two_tower_recommender_model_b200 as tt
    specs = [TableSpec(f"t_{c}", emb[i], dim, [c]) for i, c in enumerate(CAT)]
    ref = oracle.OracleTwoTower(specs, layers, loss="bce" if loss == "bce" else "softmax",
                                sparse_optimizer=sparse_opt, sparse_lr=lr, dense_lr=lr, seed=3, dense_optimizer=dense)
    # --- the reference's main() (03_model_training.py:770-829), with our names
    eb_configs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                  for i, c in enumerate(CAT)]
    ebc = tt.EmbeddingBagCollection(tables=eb_configs, device=torch.device("meta"))
    two_tower = tt.TwoTower(embedding_bag_collection=ebc, layer_sizes=layers, device=cuda)
    task = tt.TwoTowerTrainTask(two_tower, loss=loss)
    cls = tt.RowWiseAdagrad if sparse_opt == "rowwise_adagrad" else tt.RowWiseAdam
    apply_optimizer_in_backward(cls, task.two_tower.ebc.parameters(), {"lr": lr})
    model = tt.DistributedModelParallel(module=task, device=cuda)
    model.module.two_tower.load_state_dict(ref.torchrec_state_dict())
    factory = (lambda params: torch.optim.Adam(params, lr=lr)) if dense == "adam" else (lambda params: torch.optim.SGD(params, lr=lr))
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), factory)
    return ref, model, opt


@pytest.mark.parametrize("loss,sparse_opt", [("bce", "rowwise_adagrad"), ("in_batch_softmax", "rowwise_adagrad"),
                                             ("bce", "rowwise_adam"), ("in_batch_softmax", "rowwise_adam")])
def test_train_steps_match_oracle(cuda, loss, sparse_opt):
    import two_tower_recommender_model_b200 as tt
    emb, dim, layers, B, lr = [193, 9740], 64, [128, 64], 1024, 0.01  # workshop/02-mosaic-model-training.py:135-136 sizes
    # softmax + Adam is ill-conditioned for cross-implementation parity (see oracle/two_tower.py)
    ref, model, opt = build_models(cuda, emb, dim, layers, loss, sparse_opt, lr, dense="adam" if loss == "bce" else "sgd")
    batches = make_batches(6, B, emb, seed=11)

    def transform(b):
        v, l, y = oracle.transform_to_torchrec_batch(b, CAT, emb)   # the reference's own host transform
        return tt.Batch(dense_features=torch.zeros(1), sparse_features=tt.KeyedJaggedTensor.from_lengths_sync(CAT, v, l), labels=y)

    pipeline = tt.TrainPipelineSparseDist(model, opt, cuda)
    assert pipeline._model is model and pipeline._device == cuda and pipeline._optimizer.param_groups[0]["lr"] == lr
    pipeline._model.train()
    it = map(transform, iter(batches))
    n = 0
    for b in batches:
        v, l, y = oracle.transform_to_torchrec_batch(b, CAT, emb)
        loss_r, logits_r = ref.train_step(CAT, v, l, y)
        loss_d, logits_d, labels_d = pipeline.progress(it)
        torch.testing.assert_close(loss_d.cpu(), loss_r, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(logits_d.cpu().reshape(-1), logits_r.reshape(-1), rtol=1e-3, atol=1e-5)
        assert torch.equal(labels_d.cpu(), y)
        n += 1
    with pytest.raises(StopIteration):
        pipeline.progress(it)
    assert n == 6
    want = ref.torchrec_state_dict()
    got = model.module.two_tower.state_dict()
    assert set(got.keys()) == set(want.keys())
    for k in want:
        torch.testing.assert_close(got[k].cpu(), want[k], rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")
    # eval mode: no update
    pipeline._model.eval()
    before = {k: v.clone() for k, v in model.module.two_tower.state_dict().items()}
    out = pipeline.progress(map(transform, iter(batches[:1])))
    assert out[0].ndim == 0
    for k, v in model.module.two_tower.state_dict().items():
        assert torch.equal(v, before[k])


def test_multi_feature_towers_mean_pooling(cuda):
    """BASELINE config 3 shape in miniature: mean-pooled history bag + 3 candidate features."""
    import two_tower_recommender_model_b200 as tt
    from helpers import random_kjt
    specs = [TableSpec("t_hist", 500, 32, ["hist"], "mean"), TableSpec("t_product", 400, 32, ["product"]),
             TableSpec("t_aisle", 134, 32, ["aisle"]), TableSpec("t_department", 21, 32, ["department"])]
    keys = ["hist", "product", "aisle", "department"]
    ref = oracle.OracleTwoTower(specs, [64, 32], loss="softmax", sparse_lr=0.02, dense_lr=0.02,
                                query_features=["hist"], candidate_features=["product", "aisle", "department"], seed=5,
                                dense_optimizer="sgd")
    cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=32, num_embeddings=s.num_embeddings, feature_names=list(s.feature_names),
                                  pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in specs]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=cuda)
    tower = tt.TwoTower(ebc, [64, 32], device=cuda, query_features=["hist"], candidate_features=["product", "aisle", "department"])
    task = tt.TwoTowerTrainTask(tower, loss="in_batch_softmax")
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.02})
    tower.load_state_dict(ref.torchrec_state_dict())
    opt = torch.optim.SGD([p for n, p in task.named_parameters() if "embedding_bags" not in n], lr=0.02)
    B = 256
    for step in range(4):
        g = torch.Generator().manual_seed(step)
        lens = torch.cat([torch.randint(0, 21, (B,), generator=g), torch.ones(3 * B, dtype=torch.int64)]).to(torch.int32)
        vals = torch.cat([torch.randint(0, 500, (int(lens[:B].sum()),), generator=g), torch.randint(0, 400, (B,), generator=g),
                          torch.randint(0, 134, (B,), generator=g), torch.randint(0, 21, (B,), generator=g)])
        y = torch.zeros(B, dtype=torch.int32)
        loss_r, _ = ref.train_step(keys, vals, lens, y)
        opt.zero_grad()
        batch = tt.Batch(torch.zeros(1), tt.KeyedJaggedTensor.from_lengths_sync(keys, vals, lens), y).to(cuda)
        loss, _ = task(batch)
        loss.backward()
        opt.step()
        torch.testing.assert_close(loss.detach().cpu(), loss_r, rtol=1e-4, atol=1e-6)
    want = ref.torchrec_state_dict()
    for k, v in tower.state_dict().items():
        torch.testing.assert_close(v.cpu(), want[k], rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")


def test_corpus_embedding_and_retrieval(cuda):
    """03_model_training.py:1056-1122 + 04_evaluate_retrieval.py:134-141 on the device."""
    import two_tower_recommender_model_b200 as tt
    emb, dim, layers = [300, 1200], 32, [64, 32]
    specs = [TableSpec(f"t_{c}", emb[i], dim, [c]) for i, c in enumerate(CAT)]
    ref = oracle.OracleTwoTower(specs, layers, seed=9)
    ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                            for i, c in enumerate(CAT)], device=cuda)
    model = tt.TwoTower(ebc, layers, device=cuda)
    model.load_state_dict(ref.torchrec_state_dict())
    model.eval()
    kjt = tt.create_keyed_jagged_tensor(emb[1], CAT, "product_id", device=cuda)
    assert kjt.length_per_key() == [0, emb[1]] and kjt.keys() == CAT
    items = tt.process_embeddings(model, kjt, "product_id")
    users = tt.embed_corpus(model, CAT, "user_id", emb[0], cuda, chunk=128)
    with torch.no_grad():
        w_items = ref.embedding_bags["t_product_id"].weight
        w_users = ref.embedding_bags["t_user_id"].weight
        it_r, us_r = w_items, w_users
        for lin in ref.candidate_proj:
            it_r = torch.relu(lin(it_r))
        for lin in ref.query_proj:
            us_r = torch.relu(lin(us_r))
    torch.testing.assert_close(items.cpu(), it_r, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(users.cpu(), us_r, rtol=1e-5, atol=1e-5)
    index = tt.BruteForceIndex(items)
    scores, ids = index.search(users, num_results=100)
    ws, wi = oracle.exact_topk(users.cpu(), items.cpu(), 100)
    torch.testing.assert_close(scores.cpu(), ws, rtol=1e-5, atol=1e-5)
    recall = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ids.cpu(), wi)) / wi.numel()
    assert recall >= 0.999
    resp = index.similarity_search(query_vector=users[0].tolist(), columns=["product_id"], num_results=100)
    assert [c["name"] for c in resp["manifest"]["columns"]] == ["product_id", "score"] and len(resp["result"]["data_array"]) == 100
    targets = [wi[i, :10].tolist() for i in range(emb[0])]
    m_dev = tt.retrieval_metrics(ids, targets, 100)
    m_ref = oracle.retrieval_metrics(ids.cpu().tolist(), targets, 100)
    for k in m_ref:
        assert abs(m_dev[k] - m_ref[k]) < 1e-5


def test_reference_utils_file_imports_through_shim(cuda):
    """install_torchrec_shim() makes every torchrec name utils/model_training.py:20-41 imports resolve."""
    import importlib
    import two_tower_recommender_model_b200 as tt
    tt.install_torchrec_shim()
    for mod, names in {
        "torchrec.distributed": ["TrainPipelineSparseDist"],
        "torchrec.distributed.model_parallel": ["DistributedModelParallel", "get_default_sharders"],
        "torchrec.inference.state_dict_transform": ["state_dict_gather", "state_dict_to_device"],
        "torchrec.modules.embedding_configs": ["EmbeddingBagConfig"],
        "torchrec.modules.embedding_modules": ["EmbeddingBagCollection"],
        "torchrec.optim.keyed": ["KeyedOptimizerWrapper"],
        "torchrec.optim.rowwise_adagrad": ["RowWiseAdagrad"],
        "torchrec.sparse.jagged_tensor": ["KeyedJaggedTensor"],
        "torchrec.datasets.utils": ["Batch"],
        "torchrec.modules.mlp": ["MLP"],
        "torchrec.distributed.comm": ["get_local_size"],
        "torchrec.distributed.planner": ["EmbeddingShardingPlanner", "Topology"],
        "torchrec.distributed.planner.storage_reservations": ["HeuristicalStorageReservation"],
    }.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"


def test_bf16_tensor_core_training_tracks_fp32_oracle(cuda):
    """precision="bf16" (tcgen05 towers + in-batch softmax): the loss curve over a few steps stays
    within rtol 1e-2 of the fp32 CPU oracle (tolerance north_star states for bf16 losses)."""
    import two_tower_recommender_model_b200 as tt
    emb, dim, layers, B, lr = [193, 9740], 64, [128, 64], 2048, 0.01
    specs = [TableSpec(f"t_{c}", emb[i], dim, [c]) for i, c in enumerate(CAT)]
    ref = oracle.OracleTwoTower(specs, layers, loss="softmax", sparse_lr=lr, dense_lr=lr, seed=3, dense_optimizer="sgd")
    ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                            for i, c in enumerate(CAT)], device=torch.device("meta"))
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, layers, device=cuda, precision="bf16"), loss="in_batch_softmax", precision="bf16")
    apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": lr})
    model = tt.DistributedModelParallel(module=task, device=cuda)
    model.module.two_tower.load_state_dict(ref.torchrec_state_dict())
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: torch.optim.SGD(p, lr=lr))
    model.train()
    for b in make_batches(5, B, emb, seed=21):
        v, l, y = oracle.transform_to_torchrec_batch(b, CAT, emb)
        loss_r, _ = ref.train_step(CAT, v, l, y)
        batch = tt.Batch(torch.zeros(1), tt.KeyedJaggedTensor.from_lengths_sync(CAT, v, l), y).to(cuda)
        opt.zero_grad()
        loss, _ = model(batch)
        loss.backward()
        opt.step()
        torch.testing.assert_close(loss.detach().cpu(), loss_r, rtol=1e-2, atol=1e-3)


def test_cuda_graph_step_equals_eager(cuda):
    """CudaGraphTrainStep (warm-up eager, then capture + replay) trains exactly like the eager loop.  Run with the
    deterministic (two-pass) softmax backward: the one-pass kernel adds its partial products in L2 in CTA arrival
    order, and bf16 re-rounding downstream amplifies those low-bit differences beyond any useful tolerance."""
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import functional as F
    F.set_deterministic_softmax_backward(True)
    try:
        _graph_vs_eager(cuda, tt)
    finally:
        F.set_deterministic_softmax_backward(False)


def _graph_vs_eager(cuda, tt):
    emb, dim, layers, B, lr = [5000, 3000], 64, [128, 64], 1024, 0.01

    def build():
        torch.manual_seed(0)
        ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                                for i, c in enumerate(CAT)], device=cuda)
        task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, layers, device=cuda, precision="bf16"), loss="in_batch_softmax", precision="bf16")
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": lr})
        opt = tt.KeyedOptimizerWrapper(dict(task.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))
        return task, opt

    g = torch.Generator().manual_seed(5)
    data = [(torch.stack([torch.randint(0, 2 * emb[0], (B,), generator=g), torch.randint(0, 2 * emb[1], (B,), generator=g)]),
             torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)) for _ in range(8)]
    m1, o1 = build()
    m2, o2 = build()
    m2.load_state_dict(m1.state_dict())
    rows = torch.tensor(emb, device=cuda)
    losses1 = []
    for ids, y in data:
        batch = tt.Batch(torch.zeros(1, device=cuda), tt.KeyedJaggedTensor.from_id_columns(CAT, ids.to(cuda), rows), y.to(cuda))
        o1.zero_grad()
        loss, _ = m1(batch)
        loss.backward()
        o1.step()
        losses1.append(float(loss.detach()))
    step = tt.CudaGraphTrainStep(m2, o2, CAT, emb, B, cuda, warmup_steps=3)
    losses2 = [float(step(ids.pin_memory(), y.pin_memory())[0]) for ids, y in data]
    assert step.captured
    assert losses1 == pytest.approx(losses2, rel=1e-6)
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=lambda m: f"{k}: {m}")


def test_ray_tune_variant_towers(cuda):
    from helpers import random_kjt
    """ray_tune_optuna_tuning_alex_test.py:227-306: several features per tower, different layer stacks per tower and
    dense features concatenated to the tower inputs; forward against plain torch on the same weights (fp32 path)."""
    import two_tower_recommender_model_b200 as tt
    feats_u, feats_i = ["u_a", "u_b"], ["i_a"]
    dims = {"u_a": 36, "u_b": 4, "i_a": 36}
    rows = {"u_a": 100, "u_b": 7, "i_a": 90}
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=dims[k], num_embeddings=rows[k], feature_names=[k]) for k in dims]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=cuda)
    model = tt.TwoTower(ebc, [[64, 16], [32, 16]], device=cuda, query_features=feats_u, candidate_features=feats_i,
                        dense_index=3, dense_dim=5)
    task = tt.TwoTowerTrainTask(model)
    B = 97
    g = torch.Generator().manual_seed(0)
    keys = list(dims)
    v, l = random_kjt(keys, [rows[k] for k in keys], B, 1, seed=5)
    dense = torch.randn(B, 5, generator=g)
    labels = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
    batch = tt.Batch(dense.to(cuda), tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda)), labels.to(cuda))
    loss, (_, logits, _) = task(batch)
    # plain torch
    sd = {k: t.detach().cpu() for k, t in model.state_dict().items()}
    offs = oracle.lengths_to_offsets(l).long()
    pooled = {}
    for i, k in enumerate(keys):
        w = sd[f"ebc.embedding_bags.t_{k}.weight"]
        out = torch.zeros(B, dims[k])
        for b in range(B):
            for p in range(int(offs[i * B + b]), int(offs[i * B + b + 1])):
                out[b] += w[v[p]]
        pooled[k] = out

    def mlp(x, prefix, n):
        for j in range(n):
            x = torch.relu(x @ sd[f"{prefix}._mlp.{j}._linear.weight"].t() + sd[f"{prefix}._mlp.{j}._linear.bias"])
        return x
    q = mlp(torch.cat([pooled["u_a"], pooled["u_b"], dense[:, :3]], 1), "query_proj", 2)
    c = mlp(torch.cat([pooled["i_a"], dense[:, 3:]], 1), "candidate_proj", 2)
    want_logits = (q * c).sum(1)
    want = torch.nn.functional.binary_cross_entropy_with_logits(want_logits, labels.float())
    torch.testing.assert_close(logits.cpu(), want_logits, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss.detach().cpu(), want, rtol=1e-5, atol=1e-6)
    loss.backward()
    assert model.query_proj._mlp[0]._linear.weight.grad is not None
    # backward of the same variant: plain-torch autograd on the same weights (tower parameters, dense tables' gradients)
    leaf = {k: t.clone().requires_grad_(True) for k, t in sd.items()}

    def mlp_g(x, prefix, n):
        for j in range(n):
            x = torch.relu(x @ leaf[f"{prefix}._mlp.{j}._linear.weight"].t() + leaf[f"{prefix}._mlp.{j}._linear.bias"])
        return x
    pooled_g = {}
    for i, k in enumerate(keys):
        bag = torch.repeat_interleave(torch.arange(B), l[i * B:(i + 1) * B].long())
        rows_k = v[int(offs[i * B]):int(offs[(i + 1) * B])]
        pooled_g[k] = torch.zeros(B, dims[k]).index_add(0, bag, leaf[f"ebc.embedding_bags.t_{k}.weight"][rows_k])
    qg = mlp_g(torch.cat([pooled_g["u_a"], pooled_g["u_b"], dense[:, :3]], 1), "query_proj", 2)
    cg = mlp_g(torch.cat([pooled_g["i_a"], dense[:, 3:]], 1), "candidate_proj", 2)
    torch.nn.functional.binary_cross_entropy_with_logits((qg * cg).sum(1), labels.float()).backward()
    for k, p in model.state_dict(keep_vars=True).items():
        assert p.grad is not None, k
        torch.testing.assert_close(p.grad.cpu(), leaf[k].grad, rtol=1e-4, atol=1e-6, msg=lambda m: f"grad of {k}: {m}")


def test_cuda_graph_multi_hot_kjt_step_equals_eager(cuda):
    """CudaGraphTrainStep in kjt_capacity mode: multi-hot batches (mean-pooled history + single-id features, the shape of
    BASELINE configs[2]) whose number of ids CHANGES from step to step replay through one captured graph and train
    exactly like the eager loop (deterministic softmax backward, see above)."""
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import functional as F
    F.set_deterministic_softmax_backward(True)
    try:
        keys, rows, D, B, Lmax = ["hist", "item"], [4000, 900], 128, 512, 7

        def build():
            torch.manual_seed(1)
            cfgs = [tt.EmbeddingBagConfig(name="t_hist", embedding_dim=D, num_embeddings=rows[0], feature_names=["hist"], pooling=tt.PoolingType.MEAN),
                    tt.EmbeddingBagConfig(name="t_item", embedding_dim=D, num_embeddings=rows[1], feature_names=["item"])]
            ebc = tt.EmbeddingBagCollection(tables=cfgs, device=cuda)
            tower = tt.TwoTower(ebc, [128, 64], device=cuda, query_features=["hist"], candidate_features=["item"], precision="bf16")
            task = tt.TwoTowerTrainTask(tower, loss="in_batch_softmax", precision="bf16")
            apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.01})
            return task, tt.KeyedOptimizerWrapper(dict(task.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))

        g = torch.Generator().manual_seed(9)
        data = []
        for _ in range(8):
            lens = torch.cat([torch.randint(0, Lmax + 1, (B,), generator=g), torch.ones(B, dtype=torch.int64)]).to(torch.int32)
            n_h = int(lens[:B].sum())
            vals = torch.cat([torch.randint(0, rows[0], (n_h,), generator=g), torch.randint(0, rows[1], (B,), generator=g)])
            data.append((vals, lens, torch.zeros(B, dtype=torch.int32)))
        assert len({v.numel() for v, _, _ in data}) > 1          # the id count really varies
        m1, o1 = build()
        m2, o2 = build()
        m2.load_state_dict(m1.state_dict())
        losses1 = []
        for vals, lens, y in data:
            batch = tt.Batch(torch.zeros(1, device=cuda), tt.KeyedJaggedTensor.from_lengths_sync(keys, vals.to(cuda), lens.to(cuda)), y.to(cuda))
            o1.zero_grad()
            loss, _ = m1(batch)
            loss.backward()
            o1.step()
            losses1.append(float(loss.detach()))
        step = tt.CudaGraphTrainStep(m2, o2, keys, rows, B, cuda, warmup_steps=3, kjt_capacity=B * (Lmax + 1))
        losses2 = [float(step.step_kjt(v.pin_memory(), l.pin_memory(), y.pin_memory())[0]) for v, l, y in data]
        assert step.captured
        assert losses1 == pytest.approx(losses2, rel=1e-6)
        for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=lambda m: f"{k}: {m}")
    finally:
        F.set_deterministic_softmax_backward(False)
