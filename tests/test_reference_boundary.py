"""Boundary test that EXECUTES the reference: /root/reference/utils/model_training.py is compiled unchanged
after ``install_torchrec_shim()`` and its own ``transform_to_torchrec_batch``, ``TwoTower``,
``TwoTowerTrainTask``, ``train`` and ``evaluate`` bodies are run against this package's torchrec surface
(KeyedJaggedTensor, Batch, EmbeddingBagConfig / EmbeddingBagCollection, MLP, RowWiseAdagrad +
apply_optimizer_in_backward, KeyedOptimizerWrapper, planner, DistributedModelParallel,
TrainPipelineSparseDist).

The reference only exists in the build container (no GPU), the kernels only run on the GPU box (no
reference), so this test replaces the three device entry points the reference's bodies reach --
pooled lookup with its fused row-wise Adagrad backward, ``relu(linear)``, nothing else -- by the oracle
(test infrastructure) and checks what the boundary is about: the reference's code runs unchanged, the
objects it builds have the attributes it reads, the losses and the weights after 3 iterations of ITS
``train()`` loop equal the oracle's.  The kernels themselves are held to the same oracle on the GPU
(tests/test_gpu_*.py).  Non-third-party globals the notebook defines elsewhere (``itertools``,
``cat_cols``) are injected; ``streaming`` / ``torchmetrics`` / ``mlflow`` are
stubbed (out of scope)."""
import itertools
import os
import sys
import types
from functools import partial

import pytest
import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.ebc import TableSpec  # noqa: E402

REF = "/root/reference/utils/model_training.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree only exists in the build container")

CAT, EMB, DIM, LAYERS, B, LR = ["user_id", "product_id"], [97, 53], 16, [32, 16], 24, 0.05


# ------------------------------------------------------------------ oracle stand-ins for the device work
class _OracleLookup(torch.autograd.Function):
    """EbcLookup's contract: pooled [B, sum D]; backward applies row-wise Adagrad in place and the weights get no .grad --
    or, without an in-backward optimizer, hands the tables their dense gradient."""

    @staticmethod
    def forward(ctx, ebc, kjt_keys, values, offsets, batch, *anchors):
        specs = [TableSpec(c.name, c.num_embeddings, c.embedding_dim, list(c.feature_names)) for c in ebc.embedding_bag_configs()]
        ws = [ebc.embedding_bags[s.name].weight.detach() for s in specs]
        lengths = (offsets[1:] - offsets[:-1]).to(torch.int32)
        ctx.ebc, ctx.specs, ctx.keys, ctx.n = ebc, specs, list(kjt_keys), len(anchors)
        ctx.save_for_backward(values, lengths)
        return oracle.ebc_forward(specs, ws, list(kjt_keys), values, lengths)

    @staticmethod
    def backward(ctx, g):
        values, lengths = ctx.saved_tensors
        ebc = ctx.ebc
        grads = oracle.ebc_dense_grads(ctx.specs, ctx.keys, values, lengths, g)
        if ebc._in_backward_kind() is None:
            return (None,) * 5 + tuple(grads)      # no fused optimizer: the tables receive their dense gradient
        for s, gr in zip(ctx.specs, grads):
            w = ebc.embedding_bags[s.name].weight
            cfg = next(c for c in ebc.embedding_bag_configs() if c.name == s.name)
            st = ebc._state_for(cfg, w, ebc._in_backward_kind())["sum"]
            oracle.rowwise_adagrad_dense(w.data, st, gr, lr=w._optimizer_kwargs[0]["lr"])
        return (None,) * 5 + (None,) * ctx.n


def _oracle_linear_act(x, w, b, relu):
    y = torch.nn.functional.linear(x, w, b)
    return torch.relu(y) if relu else y


@pytest.fixture()
def reference(monkeypatch):
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N
    from two_tower_recommender_model_b200.modules import embedding_modules, mlp
    tt.install_torchrec_shim()
    # out-of-scope third-party modules the file imports at the top (utils/model_training.py:8,36)
    streaming = types.ModuleType("streaming")
    streaming.StreamingDataset = type("StreamingDataset", (), {})
    streaming.StreamingDataLoader = type("StreamingDataLoader", (), {})

    class AUROC:                                    # torchmetrics.AUROC(task="binary"): rank statistic
        def __init__(self, task="binary"):
            self.p, self.y = [], []

        def to(self, device):
            return self

        def __call__(self, preds, labels):
            self.p.append(preds.detach().float().cpu().reshape(-1))
            self.y.append(labels.detach().float().cpu().reshape(-1))

        def compute(self):
            p, y = torch.cat(self.p), torch.cat(self.y)
            pos, neg = p[y > 0.5], p[y <= 0.5]
            if pos.numel() == 0 or neg.numel() == 0:
                return torch.tensor(0.0)
            return ((pos[:, None] > neg[None, :]).float().mean() + 0.5 * (pos[:, None] == neg[None, :]).float().mean())

    metrics = types.ModuleType("torchmetrics")
    metrics.AUROC = AUROC
    logged = {}
    mlflow = types.ModuleType("mlflow")
    mlflow.log_metric = lambda k, v: logged.__setitem__(k, v)
    for name, mod in (("streaming", streaming), ("torchmetrics", metrics), ("mlflow", mlflow)):
        monkeypatch.setitem(sys.modules, name, mod)
    # the device work -> oracle (see the module docstring)
    monkeypatch.setattr(embedding_modules, "EbcLookup", _OracleLookup)
    monkeypatch.setattr(mlp, "linear_act", _oracle_linear_act)
    monkeypatch.setattr(N, "require_cuda", lambda t, name: None)
    ns = {"__name__": "reference_model_training", "itertools": itertools, "cat_cols": list(CAT), "mlflow": mlflow}
    with open(REF) as f:
        exec(compile(f.read(), REF, "exec"), ns)
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % (29500 + os.getpid() % 400), rank=0, world_size=1)
    yield types.SimpleNamespace(**ns), logged
    if dist.is_initialized():
        dist.destroy_process_group()


def _raw_batches(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"user_id": torch.randint(0, 2 * EMB[0], (B,), generator=g).tolist(),
             "product_id": torch.randint(0, 2 * EMB[1], (B,), generator=g).tolist(),
             "label": torch.randint(0, 2, (B,), generator=g).tolist()} for _ in range(n)]


def test_reference_transform_builds_our_kjt(reference):
    ref, _ = reference
    import two_tower_recommender_model_b200 as tt
    for raw in _raw_batches(3) + [{"user_id": [0, 0, 5], "product_id": [0, 7, 0], "label": [1, 0, 1]}]:
        batch = ref.transform_to_torchrec_batch(raw, EMB)            # the reference's own loop (U:43-69)
        assert isinstance(batch, tt.Batch) and isinstance(batch.sparse_features, tt.KeyedJaggedTensor)
        v, l, y = oracle.transform_to_torchrec_batch(raw, CAT, EMB)
        kjt = batch.sparse_features
        assert kjt.keys() == CAT
        assert torch.equal(kjt.values(), v) and torch.equal(kjt.lengths(), l) and torch.equal(batch.labels, y)
        assert torch.equal(kjt.offsets().to(torch.int64), oracle.lengths_to_offsets(l).to(torch.int64))
        assert kjt.length_per_key() == [int(l[:len(raw["label"])].sum()), int(l[len(raw["label"]):].sum())]


def _build(ref, device):
    import two_tower_recommender_model_b200 as tt
    from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
    # 03_model_training.py:770-829, with the names the reference file imported through the shim
    eb_configs = [ref.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=DIM, num_embeddings=EMB[i], feature_names=[c])
                  for i, c in enumerate(CAT)]
    ebc = ref.EmbeddingBagCollection(tables=eb_configs, device=torch.device("meta"))
    two_tower = ref.TwoTower(embedding_bag_collection=ebc, layer_sizes=LAYERS, device=device)      # the reference's class
    task = ref.TwoTowerTrainTask(two_tower)                                                       # the reference's class
    apply_optimizer_in_backward(ref.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": LR})
    planner = ref.EmbeddingShardingPlanner(topology=ref.Topology(local_world_size=ref.get_local_size(), world_size=1, compute_device="cpu"),
                                           batch_size=B, storage_reservation=ref.HeuristicalStorageReservation(percentage=0.05))
    plan = planner.collective_plan(task, ref.get_default_sharders(), dist.GroupMember.WORLD)
    model = ref.DistributedModelParallel(module=task, device=device, plan=plan)
    optimizer = ref.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda params: torch.optim.Adam(params, lr=LR))
    assert isinstance(model.module.two_tower.ebc, tt.EmbeddingBagCollection)
    return model, optimizer


def test_reference_train_loop_runs_on_the_shim_and_matches_the_oracle(reference, capsys):
    ref, logged = reference
    device = torch.device("cpu")
    model, optimizer = _build(ref, device)
    # attributes the reference reads (U:88-93, N03:819,1143)
    tw = model.module.two_tower
    assert tw._feature_names_query == ["user_id"] and tw._candidate_feature_names == ["product_id"]
    assert tw.query_proj._mlp[-1]._linear.out_features == LAYERS[-1]
    assert "t_user_id" in str(model._plan.plan)
    specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
    orc = oracle.OracleTwoTower(specs, LAYERS, loss="bce", sparse_lr=LR, dense_lr=LR, seed=11)
    tw.load_state_dict(orc.torchrec_state_dict())

    raws = _raw_batches(3, seed=5)
    transform_partial = partial(ref.transform_to_torchrec_batch, num_embeddings_per_feature=EMB)
    pipeline = ref.TrainPipelineSparseDist(model, optimizer, device)
    # the reference's own train() (U:255-317): 3 iterations, then StopIteration ends the epoch
    ref.train(pipeline, raws, raws, epoch=0, print_lr=True, validation_freq=None, limit_train_batches=None,
              limit_val_batches=None, transform_partial=transform_partial)
    out = capsys.readouterr().out
    assert "Total number of iterations: 3" in out and "lr: 0 0 0.050000" in out
    losses = []
    for raw in raws:
        v, l, y = oracle.transform_to_torchrec_batch(raw, CAT, EMB)
        losses.append(orc.train_step(CAT, v, l, y)[0])
    want = orc.torchrec_state_dict()
    got = tw.state_dict()
    assert set(got) == set(want)
    for k in want:
        torch.testing.assert_close(got[k], want[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
    for name, p in model.named_parameters():
        if "embedding_bags" in name:
            assert p.grad is None            # fused in backward: the tables never see a dense gradient

    # the reference's own evaluate() (U:191-253): eval mode, no update, (loss, logits, labels) contract
    before = {k: v.clone() for k, v in tw.state_dict().items()}
    avg_loss, auroc = ref.evaluate(None, pipeline, raws, "val", transform_partial)
    for k, v in tw.state_dict().items():
        assert torch.equal(v, before[k])
    orc_losses = []
    for raw in raws:
        v, l, y = oracle.transform_to_torchrec_batch(raw, CAT, EMB)
        q, c = orc.forward(CAT, v, l)
        orc_losses.append(float(orc.loss(q, c, y)[0]))
    assert abs(avg_loss - sum(orc_losses) / (3 * B)) < 1e-6          # the reference divides the summed loss by the sample count
    assert 0.0 <= auroc <= 1.0

    # the reference's checkpoint path (U:161-182): every entry is a plain tensor at world size 1
    sd = ref.gather_and_get_state_dict(model.module)
    assert set(sd) == {"two_tower." + k for k in want}


def test_reference_forward_signature(reference):
    """TwoTowerTrainTask.forward returns (loss, (loss.detach(), logits.detach(), labels.detach())) (U:124-143)."""
    ref, _ = reference
    model, _opt = _build(ref, torch.device("cpu"))
    raw = _raw_batches(1, seed=9)[0]
    batch = ref.transform_to_torchrec_batch(raw, EMB)
    loss, (l2, logits, labels) = model(batch)
    assert loss.requires_grad and not l2.requires_grad and logits.shape == (B,) and labels.dtype == torch.int32
    q, c = model.module.two_tower(batch.sparse_features)
    torch.testing.assert_close(logits, (q * c).sum(dim=1).detach())
    assert ref.get_relevant_fields(types.SimpleNamespace(epochs=1, embedding_dim=DIM, layer_sizes=LAYERS, learning_rate=LR, batch_size=B),
                                   CAT, EMB)["cat_cols"] == CAT
    chunks = [list(x) for x in ref.batched(iter(range(5)), 2)]
    assert chunks == [[0, 1], [2, 3], [4]]


def test_reference_train_with_validation_inside_the_epoch(reference, capsys):
    """``validation_freq`` (U:307-317): the reference's train() cuts the epoch into chunks with its own ``batched`` helper,
    runs its evaluate() between them and switches the pipeline's model back to training.  The pipeline here has to follow:
    fresh ``map`` objects on every progress() call, StopIteration at every chunk end, eval forwards without updates."""
    ref, _ = reference
    device = torch.device("cpu")
    model, optimizer = _build(ref, device)
    tw = model.module.two_tower
    specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
    orc = oracle.OracleTwoTower(specs, LAYERS, loss="bce", sparse_lr=LR, dense_lr=LR, seed=12)
    tw.load_state_dict(orc.torchrec_state_dict())
    raws, val = _raw_batches(4, seed=6), _raw_batches(2, seed=7)
    transform_partial = partial(ref.transform_to_torchrec_batch, num_embeddings_per_feature=EMB)
    pipeline = ref.TrainPipelineSparseDist(model, optimizer, device)
    ref.train(pipeline, raws, val, epoch=0, print_lr=False, validation_freq=2, limit_train_batches=None,
              limit_val_batches=None, transform_partial=transform_partial)
    out = capsys.readouterr().out
    assert out.count("Average loss over val set") == 2 and out.count("Total number of iterations:") == 2
    assert model.training                                   # train() switched it back after the last evaluation
    for raw in raws:
        v, l, y = oracle.transform_to_torchrec_batch(raw, CAT, EMB)
        orc.train_step(CAT, v, l, y)
    want = orc.torchrec_state_dict()
    for k, t in tw.state_dict().items():
        torch.testing.assert_close(t, want[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")


def test_reference_train_val_test_and_checkpoint_logging(reference, monkeypatch, capsys):
    """The reference's top-level loop ``train_val_test`` (U:320-372) on the shim: base evaluation, 2 epochs of its train() +
    evaluate(), ``log_state_dict_to_mlflow`` -> ``gather_and_get_state_dict`` (U:161-189) after every epoch, final test
    evaluation.  The logged checkpoints carry TorchRec's key names under ``two_tower.``; stripped of that prefix (as
    03_model_training.py:1026-1052 does on reload) the last one loads into a fresh TwoTower and reproduces the trained model."""
    ref, logged = reference
    import two_tower_recommender_model_b200 as tt
    device = torch.device("cpu")
    model, optimizer = _build(ref, device)
    tw = model.module.two_tower
    specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
    orc = oracle.OracleTwoTower(specs, LAYERS, loss="bce", sparse_lr=LR, dense_lr=LR, seed=13)
    tw.load_state_dict(orc.torchrec_state_dict())
    saved = {}
    ref.mlflow.pytorch = types.SimpleNamespace(
        log_state_dict=lambda sd, artifact_path: saved.__setitem__(artifact_path, {k: v.detach().clone() for k, v in sd.items()}))
    monkeypatch.setenv("RANK", "0")
    train, val, test = _raw_batches(3, seed=8), _raw_batches(2, seed=9), _raw_batches(2, seed=10)
    args = types.SimpleNamespace(epochs=2, print_lr=False, validation_freq=None, limit_train_batches=None, limit_val_batches=None,
                                 limit_test_batches=None)
    transform_partial = partial(ref.transform_to_torchrec_batch, num_embeddings_per_feature=EMB)
    test_auroc = ref.train_val_test(args, model, optimizer, device, train, val, test, transform_partial)
    assert 0.0 <= test_auroc <= 1.0 and {"val_loss", "val_auroc", "test_loss", "test_auroc"} <= set(logged)
    assert sorted(saved) == ["model_state_dict_0", "model_state_dict_1"]
    for _ in range(2):
        for raw in train:
            v, l, y = oracle.transform_to_torchrec_batch(raw, CAT, EMB)
            orc.train_step(CAT, v, l, y)
    want = orc.torchrec_state_dict()
    last = saved["model_state_dict_1"]
    assert set(last) == {"two_tower." + k for k in want}
    fresh_ebc = ref.EmbeddingBagCollection(tables=[ref.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=DIM, num_embeddings=EMB[i], feature_names=[c])
                                                   for i, c in enumerate(CAT)], device=device)
    fresh = ref.TwoTower(embedding_bag_collection=fresh_ebc, layer_sizes=LAYERS, device=device)     # the reference's class
    fresh.load_state_dict({k[len("two_tower."):]: v for k, v in last.items()})
    for k, t in fresh.state_dict().items():
        torch.testing.assert_close(t, want[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
    assert isinstance(fresh.ebc, tt.EmbeddingBagCollection)


def _notebook_functions(path, names, env):
    """Executes ONLY the named top-level function definitions of a notebook source (the rest of the file drives Spark /
    MLflow / a Databricks workspace)."""
    import ast
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in keep) == sorted(names)
    ns = dict(env)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def test_reference_reload_and_corpus_embedding_functions(reference):
    """03_model_training.py:1015-1122 on the shim: the notebook's own ``get_mlflow_model`` (rebuilds EmbeddingBagConfig /
    EmbeddingBagCollection / TwoTower from the logged params, strips ``two_tower.``, load_state_dict), then its
    ``create_keyed_jagged_tensor`` and ``process_embeddings`` on the reloaded model -- item and user embeddings equal the
    oracle's towers."""
    ref, _ = reference
    import two_tower_recommender_model_b200 as tt
    nb_path = "/root/reference/03_model_training.py"
    specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
    orc = oracle.OracleTwoTower(specs, LAYERS, loss="bce", seed=14)
    logged_sd = {"two_tower." + k: v for k, v in orc.torchrec_state_dict().items()}     # what log_state_dict_to_mlflow wrote
    mlflow = sys.modules["mlflow"]
    params = {"cat_cols": repr(CAT), "emb_counts": repr(EMB), "layer_sizes": repr(LAYERS), "embedding_dim": repr(DIM)}
    mlflow.get_run = lambda run_id: types.SimpleNamespace(data=types.SimpleNamespace(params=params))
    mlflow.MlflowClient = lambda: types.SimpleNamespace(download_artifacts=lambda *a, **k: None)
    mlflow.pytorch = types.SimpleNamespace(load_state_dict=lambda path, map_location=None: dict(logged_sd))
    nb = _notebook_functions(nb_path, ["get_mlflow_model", "create_keyed_jagged_tensor", "process_embeddings"],
                             {"torch": torch, "mlflow": mlflow, "EmbeddingBagConfig": ref.EmbeddingBagConfig,
                              "EmbeddingBagCollection": ref.EmbeddingBagCollection, "TwoTower": ref.TwoTower,
                              "KeyedJaggedTensor": ref.KeyedJaggedTensor})
    model, ebc, eb_configs, cat_cols, emb_counts = nb["get_mlflow_model"]("run-1", artifact_path="model_state_dict_1", device="cpu")
    assert isinstance(ebc, tt.EmbeddingBagCollection) and cat_cols == CAT and emb_counts == EMB and len(eb_configs) == 2
    for k, want in orc.torchrec_state_dict().items():
        torch.testing.assert_close(model.state_dict()[k], want, rtol=0, atol=0)
    model.eval()
    with torch.no_grad():
        iv = torch.arange(EMB[1])
        il = torch.cat([torch.zeros(EMB[1], dtype=torch.int32), torch.ones(EMB[1], dtype=torch.int32)])
        _, items_ref = orc.forward(CAT, iv, il)
        uv = torch.arange(EMB[0])
        ul = torch.cat([torch.ones(EMB[0], dtype=torch.int32), torch.zeros(EMB[0], dtype=torch.int32)])
        users_ref, _ = orc.forward(CAT, uv, ul)
    kjt = nb["create_keyed_jagged_tensor"](EMB[1], CAT, "product_id", device="cpu")       # the notebook's own functions
    assert isinstance(kjt, tt.KeyedJaggedTensor) and kjt.length_per_key() == [0, EMB[1]]
    items = nb["process_embeddings"](model, kjt, "product_id")
    users = nb["process_embeddings"](model, nb["create_keyed_jagged_tensor"](EMB[0], CAT, "user_id", device="cpu"), "user_id")
    assert items is not None and users is not None                                        # it returns None on any exception
    torch.testing.assert_close(items, items_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(users, users_ref, rtol=1e-5, atol=1e-6)


def test_reference_ray_tune_two_tower_class_runs_on_the_shim(reference):
    """ray_tune_optuna_tuning_alex_test.py:181-306: the reference's multi-feature / dense-concat ``TwoTower`` class, extracted
    unchanged, built on THIS package's EmbeddingBagCollection / MLP / KeyedJaggedTensor / Batch.  Its embeddings, logits,
    loss and tower gradients equal what the same class computed on stock torch (tests/golden/reference_raytune.npz)."""
    import ast
    from typing import List, Optional, Tuple
    import two_tower_recommender_model_b200 as tt
    from helpers import load_raytune_golden
    ref, _ = reference
    path = "/root/reference/ray_tune_optuna_tuning_alex_test.py"
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TwoTower"]
    ns = {"torch": torch, "nn": nn, "List": List, "Optional": Optional, "Tuple": Tuple, "MLP": ref.MLP, "Batch": ref.Batch,
          "EmbeddingBagCollection": ref.EmbeddingBagCollection}
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    G = load_raytune_golden()
    keys = list(G["dims"])
    ebc = ref.EmbeddingBagCollection(tables=[ref.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=G["dims"][k], num_embeddings=G["rows"][k],
                                                                    feature_names=[k]) for k in keys], device=torch.device("cpu"))
    user_in = sum(G["dims"][k] for k in G["feats_u"]) + G["dense_index"]
    item_in = sum(G["dims"][k] for k in G["feats_i"]) + (G["dense_dim"] - G["dense_index"])
    model = ns["TwoTower"](ebc, G["layers"], [user_in, item_in], G["feats_u"], G["feats_i"], dense_index=G["dense_index"],
                           device=torch.device("cpu"))
    assert isinstance(model.ebc, tt.EmbeddingBagCollection) and isinstance(model.user_proj, tt.MLP)
    rename = {"query_proj": "user_proj", "candidate_proj": "item_proj"}           # the fixture uses this package's tower names
    model.load_state_dict({".".join([rename.get(k.split(".")[0], k.split(".")[0])] + k.split(".")[1:]): v for k, v in G["weights"].items()})
    batch = ref.Batch(dense_features=G["dense"], sparse_features=ref.KeyedJaggedTensor.from_lengths_sync(keys, G["values"], G["lengths"]),
                      labels=G["labels"])
    q, c = model(batch)                                                           # the reference's forward
    torch.testing.assert_close(q.detach(), G["q"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(c.detach(), G["c"], rtol=1e-5, atol=1e-6)
    logits = (q * c).sum(dim=1).squeeze()
    loss = nn.BCEWithLogitsLoss()(logits, G["labels"].float())
    torch.testing.assert_close(loss.detach(), G["loss"], rtol=1e-6, atol=1e-7)
    loss.backward()
    for name, p in model.named_parameters():
        head, rest = name.split(".", 1)
        ours = {"user_proj": "query_proj", "item_proj": "candidate_proj"}.get(head, head) + "." + rest
        torch.testing.assert_close(p.grad, G["grads"][ours], rtol=1e-5, atol=1e-8, msg=lambda m: f"grad of {name}: {m}")


def test_reference_ray_tune_transform_and_train_task_run_on_the_shim(reference):
    """ray_tune_optuna_tuning_alex_test.py:121-153,320-375: the Ray-Tune variant's ``transform_to_torchrec_batch`` (dense
    columns concatenated into ``Batch.dense_features``) and its ``TwoTowerTrainTask`` (``return_sparse``: reads
    ``batch.sparse_features[name].values()``; ``WeightedBCELoss`` on the first dense columns) on this package's types;
    logits / loss against the oracle's towers with the dense features concatenated."""
    import ast
    from dataclasses import dataclass, field
    from typing import List, Optional, Tuple
    ref, _ = reference
    path = "/root/reference/ray_tune_optuna_tuning_alex_test.py"
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    want_names = {"TwoTower", "WeightedBCELoss", "TwoTowerTrainTask", "transform_to_torchrec_batch"}
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in want_names]
    assert {n.name for n in body} == want_names
    ns = {"torch": torch, "nn": nn, "List": List, "Optional": Optional, "Tuple": Tuple, "MLP": ref.MLP, "Batch": ref.Batch,
          "EmbeddingBagCollection": ref.EmbeddingBagCollection, "KeyedJaggedTensor": ref.KeyedJaggedTensor,
          "dataclass": dataclass, "field": field}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)

    cat, emb, dim, Bn = ["u_a", "i_a"], [50, 40], 8, 12
    g = torch.Generator().manual_seed(2)
    raw = {"u_a": torch.randint(0, 120, (Bn,), generator=g).tolist(), "i_a": torch.randint(0, 90, (Bn,), generator=g).tolist(),
           "label": torch.randint(0, 2, (Bn,), generator=g).tolist(),
           "d0": torch.rand(Bn, generator=g), "d1": torch.rand(Bn, 2, generator=g)}      # a 1-d and a 2-d dense column
    batch = ns["transform_to_torchrec_batch"](raw, emb, cat, dense_cols=["d0", "d1"])     # the reference's function
    assert isinstance(batch, ref.Batch) and batch.dense_features.shape == (Bn, 3)
    v, l, y = oracle.transform_to_torchrec_batch(raw, cat, emb)
    assert torch.equal(batch.sparse_features.values(), v) and torch.equal(batch.sparse_features.lengths(), l)

    specs = [TableSpec(f"t_{c}", emb[i], dim, [c]) for i, c in enumerate(cat)]
    orc = oracle.OracleTwoTower(specs, [[16, 4], [8, 4]], loss="bce", seed=3, query_features=["u_a"], candidate_features=["i_a"],
                                dense_index=1, dense_dim=3)
    ebc = ref.EmbeddingBagCollection(tables=[ref.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                             for i, c in enumerate(cat)], device=torch.device("cpu"))
    tower = ns["TwoTower"](ebc, [[16, 4], [8, 4]], [dim + 1, dim + 2], ["u_a"], ["i_a"], dense_index=1, device=torch.device("cpu"))
    rename = {"query_proj": "user_proj", "candidate_proj": "item_proj"}
    tower.load_state_dict({".".join([rename.get(k.split(".")[0], k.split(".")[0])] + k.split(".")[1:]): t
                           for k, t in orc.torchrec_state_dict().items()})
    q_ref, c_ref = orc.forward(cat, v, l, batch.dense_features.detach())
    logits_ref = (q_ref * c_ref).sum(dim=1)

    task = ns["TwoTowerTrainTask"](tower, return_sparse=True, sparse_feature_names=["u_a"])
    loss, (l2, logits, labels, sparse_values) = task(batch)
    torch.testing.assert_close(logits, logits_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss.detach(), torch.nn.functional.binary_cross_entropy_with_logits(logits_ref, y.float()).detach(), rtol=1e-6, atol=1e-7)
    nu = int(l[:Bn].sum())
    assert torch.equal(sparse_values["u_a"], v[:nu]) and torch.equal(labels, y)
    # weighted loss: BCELoss(reduction="none") on the probabilities, weighted by the interaction type in dense[:, :3]
    onehot = torch.zeros(Bn, 3)
    onehot[torch.arange(Bn), torch.randint(0, 3, (Bn,), generator=g)] = 1.0
    wbatch = ref.Batch(dense_features=onehot, sparse_features=batch.sparse_features, labels=batch.labels)
    weights = {(1, 0, 0): 1.0, (0, 1, 0): 2.0, (0, 0, 1): 0.5}
    wtask = ns["TwoTowerTrainTask"](tower, loss_fn=ns["WeightedBCELoss"](weights), return_sparse=False)
    wloss, (_, wlogits, _) = wtask(wbatch)
    q2, c2 = orc.forward(cat, v, l, onehot)
    p2 = torch.sigmoid((q2 * c2).sum(dim=1))
    per = torch.nn.functional.binary_cross_entropy(p2, y.float(), reduction="none")
    wts = torch.tensor([weights[tuple(int(x) for x in row)] for row in onehot.tolist()])
    torch.testing.assert_close(wloss.detach(), (per * wts).mean().detach(), rtol=1e-5, atol=1e-7)


def test_workshop_serving_wrapper_runs_on_the_shim(reference):
    """workshop/02-mosaic-model-training.py:1120-1210: the pyfunc ``TwoTowerWrapper`` (serving is out of scope for speed,
    SURVEY section 2, but it is a caller of the path): its ``_transform_to_torchrec_batch`` keeps id 0 as a real row
    (``is not None`` instead of truthiness) and ``predict`` returns sigmoid(q . c) as a list -- on this package's types,
    against the oracle's towers."""
    import ast
    from typing import Dict, List, Optional
    import numpy as np
    ref, _ = reference
    path = "/root/reference/workshop/02-mosaic-model-training.py"
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TwoTowerWrapper"]
    assert len(cls) == 1
    ns = {"torch": torch, "np": np, "Dict": Dict, "List": List, "Optional": Optional, "PythonModel": object, "Batch": ref.Batch,
          "KeyedJaggedTensor": ref.KeyedJaggedTensor, "cat_cols": list(CAT), "emb_counts": list(EMB)}
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
    orc = oracle.OracleTwoTower(specs, LAYERS, loss="bce", seed=15)
    ebc = ref.EmbeddingBagCollection(tables=[ref.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=DIM, num_embeddings=EMB[i], feature_names=[c])
                                             for i, c in enumerate(CAT)], device=torch.device("cpu"))
    tower = ref.TwoTower(embedding_bag_collection=ebc, layer_sizes=LAYERS, device=torch.device("cpu"))
    tower.load_state_dict(orc.torchrec_state_dict())
    wrapper = ns["TwoTowerWrapper"](tower, torch.device("cpu"))
    users, items = [0, 5, 200, 96], [7, 0, 53, 52]                    # id 0 stays a row here; 200 % 97 = 6, 53 % 53 = 0
    probs = wrapper.predict(None, {"user_id": users, "product_id": items})
    assert isinstance(probs, list) and len(probs) == 4
    v = torch.tensor([u % EMB[0] for u in users] + [i % EMB[1] for i in items])
    with torch.no_grad():
        q, c = orc.forward(CAT, v, torch.ones(8, dtype=torch.int32))
    torch.testing.assert_close(torch.tensor(probs), torch.sigmoid((q * c).sum(dim=1)), rtol=1e-5, atol=1e-6)


def test_reference_extract_columns_reads_our_search_response(reference, monkeypatch):
    """04_evaluate_retrieval.py:117-141: the notebook asks ``index.similarity_search(query_vector=, columns=, num_results=k)``
    and feeds the response to its own ``extract_columns``.  ``BruteForceIndex`` stands in for the remote index; the notebook's
    function, extracted unchanged, reads its response (the scoring kernel is replaced by the oracle here)."""
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import retrieval
    monkeypatch.setattr(retrieval, "score_topk", lambda q, items, k, item_index_base=0, precision="fp32", items_bf16=None:
                        oracle.exact_topk(q, items, k))
    # this notebook holds cell magics, so it does not parse as a module: cut the one function out by its indentation
    lines = open("/root/reference/04_evaluate_retrieval.py").read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith("def extract_columns("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].strip() and not lines[i].startswith((" ", "\t")))
    nb = {}
    exec(compile("\n".join(lines[start:end]), "04_evaluate_retrieval.py:extract_columns", "exec"), nb)
    g = torch.Generator().manual_seed(4)
    items = torch.randn(300, 16, generator=g)
    user = torch.randn(16, generator=g)
    index = tt.BruteForceIndex(items)
    results = index.similarity_search(query_vector=user.tolist(), columns=["product_id", "embeddings"], num_results=100)
    got = nb["extract_columns"](results, columns=["product_id", "score"])             # the notebook's own function
    ws, wi = oracle.exact_topk(user.view(1, -1), items, 100)
    assert got["product_id_pred"] == wi[0].tolist()
    torch.testing.assert_close(torch.tensor(got["score_pred"]), ws[0], rtol=1e-6, atol=1e-6)
    assert list(map(int, got["product_id_pred"])) == got["product_id_pred"]           # N04:163 maps the ids through int()
