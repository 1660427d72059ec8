"""The CUDA path against fixtures made by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/reference_train.npz, written
by tests/golden/make_reference_golden.py in the build container: /root/reference/utils/model_training.py's
transform_to_torchrec_batch / TwoTower / TwoTowerTrainTask / train() / evaluate() bodies on stock torch).  /root/reference
does not exist on the GPU box; the fixture travels.

Same flow and tolerances as tests/test_gpu_train.py::test_train_steps_match_oracle (fp32 kernels: loss rtol 1e-4,
logits rtol 1e-3 / atol 1e-5, weights rtol 1e-4 / atol 1e-5 -- summation order only), with the reference's numbers in
place of the oracle's.  Integer work (the device batch construction) is bit-exact."""
import pytest
import torch
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward

from helpers import load_reference_golden

pytestmark = pytest.mark.gpu

CAT = ["user_id", "product_id"]


def test_device_batch_construction_equals_the_reference_transform(cuda):
    """KeyedJaggedTensor.from_id_columns (tt_kjt_from_columns) on the raw id columns of the fixture against the values /
    lengths the reference's Python loop (utils/model_training.py:43-69) made of them: modulo, id 0 = empty bag."""
    import two_tower_recommender_model_b200 as tt
    G = load_reference_golden()
    for i in range(G["steps"] + 2):
        ids = torch.stack([G["T"](f"raw{i}_{c}") for c in CAT]).to(cuda)
        kjt = tt.KeyedJaggedTensor.from_id_columns(CAT, ids, torch.tensor(G["emb"]))
        want_v, want_l = G["T"](f"batch{i}_values"), G["T"](f"batch{i}_lengths")
        assert torch.equal(kjt.lengths().cpu(), want_l)
        n = int(kjt.offsets()[-1])
        assert n == want_v.numel() and torch.equal(kjt.values()[:n].cpu(), want_v)


def test_train_steps_equal_the_reference_bodies(cuda):
    import two_tower_recommender_model_b200 as tt
    G = load_reference_golden()
    emb, dim, layers, lr = G["emb"], G["dim"], G["layers"], G["lr"]
    # the reference's main() (03_model_training.py:770-829) with this package's names
    eb_configs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c]) for i, c in enumerate(CAT)]
    ebc = tt.EmbeddingBagCollection(tables=eb_configs, device=torch.device("meta"))
    two_tower = tt.TwoTower(embedding_bag_collection=ebc, layer_sizes=layers, device=cuda)
    task = tt.TwoTowerTrainTask(two_tower)                       # BCE: the reference's loss (utils/model_training.py:128)
    apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": lr})
    model = tt.DistributedModelParallel(module=task, device=cuda)
    model.module.two_tower.load_state_dict(G["init"])
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda params: torch.optim.Adam(params, lr=lr))
    pipeline = tt.TrainPipelineSparseDist(model, opt, cuda)

    def batch(i):
        kjt = tt.KeyedJaggedTensor.from_lengths_sync(CAT, G["T"](f"batch{i}_values"), G["T"](f"batch{i}_lengths"))
        return tt.Batch(dense_features=torch.zeros(1), sparse_features=kjt, labels=G["T"](f"batch{i}_labels"))

    pipeline._model.train()
    it = iter([batch(i) for i in range(G["steps"])])
    for i in range(G["steps"]):
        loss, logits, labels = pipeline.progress(it)
        torch.testing.assert_close(loss.cpu(), G["T"](f"step{i}_loss"), rtol=1e-4, atol=1e-6, msg=lambda m: f"step {i} loss: {m}")
        torch.testing.assert_close(logits.cpu().reshape(-1), G["T"](f"step{i}_logits"), rtol=1e-3, atol=1e-5, msg=lambda m: f"step {i} logits: {m}")
        assert torch.equal(labels.cpu(), G["T"](f"batch{i}_labels"))
    with pytest.raises(StopIteration):
        pipeline.progress(it)
    got = model.module.two_tower.state_dict()
    assert set(got.keys()) == set(G["final"].keys())            # TorchRec's key names, as the reference's checkpoint has them
    for k, want in G["final"].items():
        torch.testing.assert_close(got[k].cpu(), want, rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")
    st = model.module.two_tower.ebc.fused_optimizer_state()
    for c in CAT:
        torch.testing.assert_close(st[f"t_{c}"]["sum"].cpu(), G["T"](f"sum.t_{c}"), rtol=1e-3, atol=1e-8, msg=lambda m: f"sum t_{c}: {m}")
    # evaluate(): eval mode, no update, same (loss, logits, labels) contract (utils/model_training.py:191-253)
    pipeline._model.eval()
    before = {k: v.clone() for k, v in model.module.two_tower.state_dict().items()}
    total = 0.0
    with torch.no_grad():
        ev = iter([batch(G["steps"] + j) for j in range(2)])
        for j in range(2):
            loss, logits, _ = pipeline.progress(ev)
            torch.testing.assert_close(loss.cpu(), G["T"](f"eval{j}_loss"), rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(logits.cpu().reshape(-1), G["T"](f"eval{j}_logits"), rtol=1e-3, atol=1e-5)
            total += float(loss)
    assert abs(total / (2 * G["B"]) - float(G["z"]["eval_average_loss"])) < 1e-6
    for k, v in model.module.two_tower.state_dict().items():
        assert torch.equal(v, before[k])


def test_corpus_embeddings_and_topk_equal_the_reference_functions(cuda):
    """create_keyed_jagged_tensor / process_embeddings (same signatures as 03_model_training.py:1056-1122) and the top-100
    search against what the reference's own functions computed from the trained model (tolerances of
    tests/test_gpu_train.py's retrieval test: embeddings and scores rtol 1e-5 / atol 1e-5, recall >= 0.999)."""
    import two_tower_recommender_model_b200 as tt
    G = load_reference_golden()
    emb, dim, layers = G["emb"], G["dim"], G["layers"]
    ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                            for i, c in enumerate(CAT)], device=cuda)
    model = tt.TwoTower(ebc, layers, device=cuda)
    model.load_state_dict(G["final"])
    model.eval()
    out = {}
    for key, n in (("product_id", emb[1]), ("user_id", emb[0])):
        kjt = tt.create_keyed_jagged_tensor(n, CAT, key, device=cuda)
        assert torch.equal(kjt.values().cpu(), G["T"](f"corpus_{key}_values")) and torch.equal(kjt.lengths().cpu(), G["T"](f"corpus_{key}_lengths"))
        out[key] = tt.process_embeddings(model, kjt, key)
        torch.testing.assert_close(out[key].cpu(), G["T"](f"corpus_{key}_embeddings"), rtol=1e-5, atol=1e-5)
    index = tt.BruteForceIndex(out["product_id"])
    scores, ids = index.search(out["user_id"], num_results=100)
    torch.testing.assert_close(scores.cpu(), G["T"]("top100_scores"), rtol=1e-5, atol=1e-5)
    want = G["T"]("top100_ids")
    recall = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ids.cpu(), want)) / want.numel()
    assert recall >= 0.999


def test_ray_tune_towers_equal_the_reference_class(cuda):
    """Several features per tower, one layer stack per tower, dense features concatenated to the tower inputs
    (ray_tune_optuna_tuning_alex_test.py:181-306): embeddings, logits, loss and every parameter's gradient against what the
    reference's own class computed on stock torch (tests/golden/reference_raytune.npz).  Tolerances of
    tests/test_gpu_train.py::test_ray_tune_variant_towers (fp32 kernels, summation order only)."""
    import two_tower_recommender_model_b200 as tt
    from helpers import load_raytune_golden
    G = load_raytune_golden()
    keys = list(G["dims"])
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=G["dims"][k], num_embeddings=G["rows"][k], feature_names=[k]) for k in keys]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=cuda)
    model = tt.TwoTower(ebc, G["layers"], device=cuda, query_features=G["feats_u"], candidate_features=G["feats_i"],
                        dense_index=G["dense_index"], dense_dim=G["dense_dim"])
    assert set(model.state_dict()) == set(G["weights"])
    model.load_state_dict(G["weights"])
    task = tt.TwoTowerTrainTask(model)
    batch = tt.Batch(G["dense"].to(cuda), tt.KeyedJaggedTensor.from_lengths_sync(keys, G["values"].to(cuda), G["lengths"].to(cuda)),
                     G["labels"].to(cuda))
    with torch.no_grad():
        q, c = model(batch)
    torch.testing.assert_close(q.cpu(), G["q"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(c.cpu(), G["c"], rtol=1e-4, atol=1e-5)
    loss, (_, logits, _) = task(batch)
    torch.testing.assert_close(logits.cpu(), G["logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss.detach().cpu(), G["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    for k, p in model.state_dict(keep_vars=True).items():
        assert p.grad is not None, k
        torch.testing.assert_close(p.grad.cpu(), G["grads"][k], rtol=1e-4, atol=1e-6, msg=lambda m: f"grad of {k}: {m}")
