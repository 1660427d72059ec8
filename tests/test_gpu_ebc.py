"""EmbeddingBagCollection forward / fused backward vs the oracle.
Tolerance (fp32, summation order only): rtol 1e-5, atol 1e-6."""
import pytest
import torch
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward

import oracle
from oracle.ebc import TableSpec
from helpers import random_kjt

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6

CASES = [
    # dims, rows, pooling, batch, max_len, dup_pool
    ([64, 64], [2000, 500], ["sum", "sum"], 1024, 1, 0),         # BASELINE config 1 shape
    ([128, 128], [5000, 300], ["mean", "sum"], 777, 20, 0),      # config 3 flavour: history bags, mean
    ([36, 36, 4, 4], [100, 90, 7, 3], ["sum"] * 4, 129, 3, 0),   # ray_tune variant dims (RT:533-548)
    ([64, 64], [50, 40], ["sum", "mean"], 4096, 4, 8),           # heavy duplicates
    ([7, 130], [33, 1000], ["sum", "mean"], 65, 5, 0),           # dims not multiple of 4 (scalar path)
    ([256, 512], [300, 200], ["sum", "sum"], 100, 2, 0),         # wide rows (NV=2,4)
    ([64], [10], ["sum"], 3, 0, 0),                              # all bags empty
]


def build(cuda, dims, rows, pooling):
    import two_tower_recommender_model_b200 as tt
    specs = [TableSpec(f"t_f{i}", rows[i], dims[i], [f"f{i}"], pooling[i]) for i in range(len(dims))]
    cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=s.embedding_dim, num_embeddings=s.num_embeddings,
                                  feature_names=list(s.feature_names),
                                  pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in specs]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=cuda)
    weights = [ebc.embedding_bags[s.name].weight.detach().cpu().clone() for s in specs]
    return specs, ebc, weights


@pytest.mark.parametrize("dims,rows,pooling,B,L,dup", CASES)
def test_forward(cuda, dims, rows, pooling, B, L, dup):
    import two_tower_recommender_model_b200 as tt
    specs, ebc, weights = build(cuda, dims, rows, pooling)
    keys = [f"f{i}" for i in range(len(dims))]
    v, l = random_kjt(keys, rows, B, L, seed=B + L, dup_pool=dup)
    want = oracle.ebc_forward(specs, weights, keys, v, l)
    want2 = oracle.ebc_forward_torch(specs, weights, keys, v, l)
    torch.testing.assert_close(want, want2, rtol=RTOL, atol=ATOL)
    kt = ebc(tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda)))
    assert kt.keys() == keys and kt.values().shape == (B, sum(dims))
    torch.testing.assert_close(kt.values().cpu(), want, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(kt[keys[-1]].cpu(), want[:, -dims[-1]:], rtol=RTOL, atol=ATOL)


def test_forward_key_order_independent(cuda):
    """KJT keys in a different order than the tables (and an extra unused key)."""
    import two_tower_recommender_model_b200 as tt
    specs, ebc, weights = build(cuda, [64, 64], [100, 80], ["sum", "sum"])
    B = 50
    v, l = random_kjt(["x", "f1", "f0"], [5, 80, 100], B, 3, seed=3)
    kjt_keys = ["x", "f1", "f0"]
    want = oracle.ebc_forward(specs, weights, kjt_keys, v, l)
    kt = ebc(tt.KeyedJaggedTensor.from_lengths_sync(kjt_keys, v.to(cuda), l.to(cuda)))
    torch.testing.assert_close(kt.values().cpu(), want, rtol=RTOL, atol=ATOL)
    # and the backward ignores the unused key's ids
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.1})
    go = torch.randn(B, 128)
    kt = ebc(tt.KeyedJaggedTensor.from_lengths_sync(kjt_keys, v.to(cuda), l.to(cuda)))
    kt.values().backward(go.to(cuda))
    grads = oracle.ebc_dense_grads(specs, kjt_keys, v, l, go)
    for s, w, g in zip(specs, weights, grads):
        st = torch.zeros(s.num_embeddings)
        oracle.rowwise_adagrad_dense(w, st, g, lr=0.1)
        torch.testing.assert_close(ebc.embedding_bags[s.name].weight.detach().cpu(), w, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("dims,rows,pooling,B,L,dup", CASES)
def test_backward_dense_grad(cuda, dims, rows, pooling, B, L, dup):
    """No in-backward optimizer registered -> dense [R, D] grads, equal to autograd through nn.EmbeddingBag."""
    import two_tower_recommender_model_b200 as tt
    specs, ebc, weights = build(cuda, dims, rows, pooling)
    keys = [f"f{i}" for i in range(len(dims))]
    v, l = random_kjt(keys, rows, B, L, seed=B * 3 + L, dup_pool=dup)
    go = torch.randn(B, sum(dims), generator=torch.Generator().manual_seed(1))
    want = oracle.ebc_dense_grads(specs, keys, v, l, go)
    kt = ebc(tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda)))
    kt.values().backward(go.to(cuda))
    for s, g in zip(specs, want):
        got = ebc.embedding_bags[s.name].weight.grad
        assert got is not None
        torch.testing.assert_close(got.cpu(), g, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("opt", ["adagrad", "adagrad_eps1e-8", "adam", "sgd"])
@pytest.mark.parametrize("dims,rows,pooling,B,L,dup", CASES[:5])
def test_fused_backward_optimizer(cuda, opt, dims, rows, pooling, B, L, dup):
    import two_tower_recommender_model_b200 as tt
    specs, ebc, weights = build(cuda, dims, rows, pooling)
    keys = [f"f{i}" for i in range(len(dims))]
    lr = 0.05
    if opt == "adagrad":
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": lr})
    elif opt == "adagrad_eps1e-8":
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": lr, "eps": 1e-8})
    elif opt == "adam":
        apply_optimizer_in_backward(tt.RowWiseAdam, ebc.parameters(), {"lr": lr})
    else:
        apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": lr})
    state = [dict(sum=torch.zeros(s.num_embeddings), m=torch.zeros(s.num_embeddings, s.embedding_dim),
                  v=torch.zeros(s.num_embeddings)) for s in specs]
    off_all = None
    for step in range(1, 4):
        v, l = random_kjt(keys, rows, B, L, seed=step * 1000 + B, dup_pool=dup)
        go = torch.randn(B, sum(dims), generator=torch.Generator().manual_seed(step))
        kt = ebc(tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda)))
        kt.values().backward(go.to(cuda))
        grads = oracle.ebc_dense_grads(specs, keys, v, l, go)
        offs = oracle.lengths_to_offsets(l).long()
        for i, (s, w, g) in enumerate(zip(specs, weights, grads)):
            ids = v[int(offs[i * B]):int(offs[(i + 1) * B])]
            rows_u = torch.unique(ids, sorted=True)
            if opt.startswith("adagrad"):
                oracle.rowwise_adagrad_dense(w, state[i]["sum"], g, lr=lr, eps=1e-8 if "eps" in opt else 1e-10)
            elif opt == "adam":
                oracle.rowwise_adam_sparse(w, state[i]["m"], state[i]["v"], rows_u, g[rows_u], step, lr=lr)
            else:
                w -= lr * g
            p = ebc.embedding_bags[s.name].weight
            assert p.grad is None  # fused: no dense gradient is ever produced
            torch.testing.assert_close(p.detach().cpu(), w, rtol=2e-5, atol=2e-6)
    st = ebc.fused_optimizer_state()
    for i, s in enumerate(specs):
        if opt.startswith("adagrad"):
            torch.testing.assert_close(st[s.name]["sum"].cpu(), state[i]["sum"], rtol=2e-5, atol=1e-7)
        elif opt == "adam":
            torch.testing.assert_close(st[s.name]["exp_avg"].cpu(), state[i]["m"], rtol=2e-5, atol=1e-7)
            torch.testing.assert_close(st[s.name]["exp_avg_sq"].cpu(), state[i]["v"], rtol=2e-5, atol=1e-9)


def test_sparse_equals_dense_adagrad():
    """Oracle self-check (CPU): sparse-exact row-wise Adagrad == dense form."""
    g = torch.Generator().manual_seed(0)
    R, D = 50, 8
    w1 = torch.randn(R, D, generator=g); w2 = w1.clone()
    s1 = torch.rand(R, generator=g); s2 = s1.clone()
    ids = torch.randint(0, R, (30,), generator=g)
    grad = torch.zeros(R, D); grad.index_add_(0, ids, torch.randn(30, D, generator=g))
    oracle.rowwise_adagrad_dense(w1, s1, grad, lr=0.1)
    rows = torch.unique(ids)
    oracle.rowwise_adagrad_sparse(w2, s2, rows, grad[rows], lr=0.1)
    torch.testing.assert_close(w1, w2); torch.testing.assert_close(s1, s2)
