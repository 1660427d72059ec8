"""Property tests (hypothesis) of the oracle's integer ops and of the host-side KJT / planner logic: the
size-independent invariants the GPU tests then check at BASELINE sizes."""
import os
import sys

import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.kjt import block_bucketize_vectorized  # noqa: E402


@st.composite
def jagged(draw, max_f=4, max_b=9, max_len=4, max_rows=50):
    F = draw(st.integers(1, max_f))
    B = draw(st.integers(1, max_b))
    rows = [draw(st.integers(1, max_rows)) for _ in range(F)]
    lens, vals = [], []
    for f in range(F):
        for _ in range(B):
            n = draw(st.integers(0, max_len))
            lens.append(n)
            vals += [draw(st.integers(0, rows[f] - 1)) for _ in range(n)]
    return F, B, rows, torch.tensor(vals, dtype=torch.int64), torch.tensor(lens, dtype=torch.int32)


@settings(max_examples=60, deadline=None)
@given(jagged(), st.integers(1, 5))
def test_block_bucketize_properties(j, W):
    """Every id lands in exactly one bucket, keeps its bag, local + w * block gives it back, order inside a
    (bucket, feature, bag) is preserved, and the loop form equals the vectorised form."""
    F, B, rows, v, l = j
    nl, nv, unb = oracle.block_bucketize_sparse_features(l, v, rows, W, B)
    nl2, nv2, unb2 = block_bucketize_vectorized(l, v, rows, W, B)
    assert torch.equal(nl, nl2) and torch.equal(nv, nv2) and torch.equal(unb, unb2)
    assert int(nl.sum()) == v.numel()
    assert torch.equal(nl.view(W, F * B).sum(0).to(torch.int32), l)           # bag sizes are conserved
    assert sorted(unb.tolist()) == list(range(v.numel()))                      # a permutation
    off = oracle.lengths_to_offsets(nl).to(torch.int64)
    bag_of_in = torch.repeat_interleave(torch.arange(F * B), l.to(torch.int64))
    for w in range(W):
        for f in range(F):
            block = -(-rows[f] // W)
            for b in range(B):
                s, e = int(off[(w * F + f) * B + b]), int(off[(w * F + f) * B + b + 1])
                local = nv[s:e]
                assert ((local >= 0) & (local < block)).all()
                src = [p for p in range(v.numel()) if int(bag_of_in[p]) == f * B + b and int(v[p]) // block == w]
                assert [int(unb[p]) for p in src] == list(range(s, e))         # stable inside the bag
                assert torch.equal(local + w * block, v[src])


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4), st.integers(1, 3), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_gathered_range_shard_properties(W, F, B, seed):
    """The sync-free input dist's filter (oracle.gathered_range_shard): over the W shards of a block partition every id
    of every rank's batch is kept exactly once, stays in its (feature, rank, sample) bag in source order, and
    local + lo gives it back; shard w equals bucket w of block_bucketize on the concatenated batch."""
    g = torch.Generator().manual_seed(seed)
    rows = [int(torch.randint(1, 40, (1,), generator=g)) for _ in range(F)]
    per_rank = []
    for _ in range(W):
        lens = torch.randint(0, 4, (F * B,), generator=g).to(torch.int32)
        vals = torch.cat([torch.randint(0, rows[f], (int(lens[f * B:(f + 1) * B].sum()),), generator=g) for f in range(F)]) \
            if int(lens.sum()) else torch.zeros(0, dtype=torch.int64)
        per_rank.append((vals, lens))
    total = sum(int(v.numel()) for v, _ in per_rank)
    block = [-(-r // W) for r in rows]
    kept = 0
    cat_len = torch.cat([per_rank[r][1][f * B:(f + 1) * B] for f in range(F) for r in range(W)])
    offs = [oracle.lengths_to_offsets(l).tolist() for _, l in per_rank]
    cat_val = torch.cat([per_rank[r][0][offs[r][f * B]:offs[r][(f + 1) * B]] for f in range(F) for r in range(W)])
    nl, nv, _ = oracle.block_bucketize_sparse_features(cat_len, cat_val, rows, W, W * B)
    noff = oracle.lengths_to_offsets(nl).tolist()
    n = F * W * B
    for w in range(W):
        lo = [w * b for b in block]
        hi = [min((w + 1) * b, r) for b, r in zip(block, rows)]
        v, l = oracle.gathered_range_shard([p[0] for p in per_rank], [p[1] for p in per_rank], lo, hi, B)
        assert int(l.sum()) == v.numel()
        kept += int(v.numel())
        assert (l <= cat_len).all()                                  # a bag only ever loses ids
        o = oracle.lengths_to_offsets(l).tolist()
        for f in range(F):
            seg = v[o[f * W * B]:o[(f + 1) * W * B]]
            assert ((seg >= 0) & (seg < max(hi[f] - lo[f], 0) + (1 if seg.numel() == 0 else 0))).all()
        assert l.tolist() == nl[w * n:(w + 1) * n].tolist() and v.tolist() == nv[noff[w * n]:noff[(w + 1) * n]].tolist()
    assert kept == total


@settings(max_examples=60, deadline=None)
@given(jagged(), st.data())
def test_permute_2d_properties(j, data):
    """Permuting by p then by the inverse of p is the identity; repeats duplicate segments; totals add up."""
    F, B, rows, v, l = j
    perm = data.draw(st.permutations(list(range(F))))
    pl, pv, _ = oracle.permute_2d_sparse_data(perm, l.view(F, B), v)
    inv = [perm.index(i) for i in range(F)]
    bl, bv, _ = oracle.permute_2d_sparse_data(inv, pl, pv)
    assert torch.equal(bl.reshape(-1), l) and torch.equal(bv, v)
    dup = [perm[0]] * 2
    dl, dv, _ = oracle.permute_2d_sparse_data(dup, l.view(F, B), v)
    assert dv.numel() == 2 * int(l.view(F, B)[perm[0]].sum()) and torch.equal(dl[0], dl[1])


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(-5, 40), min_size=1, max_size=30), st.lists(st.integers(-5, 40), min_size=1, max_size=30),
       st.integers(1, 17), st.integers(1, 17))
def test_transform_properties(users, items, ru, ri):
    """transform_to_torchrec_batch (utils/model_training.py:43-69): falsy id -> empty bag; every other id is taken
    modulo the table size (Python modulo: result in [0, R) also for negative ids); key-major layout."""
    B = min(len(users), len(items))
    raw = {"user_id": users[:B], "product_id": items[:B], "label": [0] * B}
    v, l, y = oracle.transform_to_torchrec_batch(raw, ["user_id", "product_id"], [ru, ri])
    assert l.numel() == 2 * B and int(l.sum()) == v.numel() and y.numel() == B
    want = [u % ru for u in users[:B] if u] + [i % ri for i in items[:B] if i]
    assert v.tolist() == want
    assert l.tolist() == [1 if u else 0 for u in users[:B]] + [1 if i else 0 for i in items[:B]]
    assert ((v[:int(l[:B].sum())] >= 0) & (v[:int(l[:B].sum())] < ru)).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.integers(1, 300), st.integers(1, 20), st.integers(0, 2 ** 31 - 1))
def test_topk_oracle_properties(Q, N, k, seed):
    """exact_topk: scores descending, ties by ascending index, the k-th score bounds every item left out."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(-4, 5, (Q, 6), generator=g).float() / 4
    it = torch.randint(-4, 5, (N, 6), generator=g).float() / 4
    s, i = oracle.exact_topk(q, it, k)
    kk = min(k, N)
    full = q.double() @ it.double().t()
    for r in range(Q):
        assert len(set(i[r].tolist())) == kk
        for a in range(kk - 1):
            assert s[r, a] > s[r, a + 1] or (s[r, a] == s[r, a + 1] and i[r, a] < i[r, a + 1])
        rest = torch.ones(N, dtype=torch.bool)
        rest[i[r]] = False
        if rest.any():
            assert float(full[r][rest].max()) <= float(s[r, kk - 1])


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 5000), st.integers(4, 64)), min_size=1, max_size=6), st.integers(1, 8))
def test_planner_properties(tables, W):
    """Every table gets a plan; table-wise owners are valid ranks; row-wise blocks cover all rows; every rank
    computes the same plan (determinism)."""
    import two_tower_recommender_model_b200 as tt
    cfgs = [tt.EmbeddingBagConfig(name=f"t{i}", embedding_dim=d // 4 * 4, num_embeddings=r, feature_names=[f"f{i}"])
            for i, (r, d) in enumerate(tables)]
    mod = torch.nn.ModuleDict({"ebc": tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))})
    plans = [tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=W)).plan(mod) for _ in range(2)]
    assert str(plans[0]) == str(plans[1])
    for c in cfgs:
        ps = plans[0].plan["ebc"][c.name]
        assert ps.sharding_type in ("table_wise", "row_wise")
        if ps.sharding_type == "table_wise":
            assert len(ps.ranks) == 1 and 0 <= ps.ranks[0] < W
        else:
            assert ps.ranks == list(range(W)) and ps.block_size * W >= c.num_embeddings


# ---- floating-point oracle functions: invariants the domain offers ---------------------------------------------------------
from oracle.ebc import TableSpec  # noqa: E402


@settings(max_examples=40, deadline=None)
@given(jagged(max_f=3), st.integers(0, 2 ** 31 - 1), st.sampled_from(["sum", "mean"]))
def test_ebc_forward_properties(j, seed, pooling):
    """ebc_forward == per-feature F.embedding_bag (what unsharded TorchRec runs); linear in the tables; empty bags give zero
    rows; the dense gradient is the adjoint of the lookup: <ebc(W), G> == <W, ebc_dense_grads(G)>."""
    F, B, rows, vals, lens = j
    g = torch.Generator().manual_seed(seed)
    keys = [f"f{i}" for i in range(F)]
    specs = [TableSpec(f"t{i}", rows[i], 4, [keys[i]], pooling) for i in range(F)]
    W1 = [torch.randn(r, 4, generator=g, dtype=torch.float64) for r in rows]
    W2 = [torch.randn(r, 4, generator=g, dtype=torch.float64) for r in rows]
    o1 = oracle.ebc_forward(specs, W1, keys, vals, lens)
    torch.testing.assert_close(o1, oracle.ebc_forward_torch(specs, W1, keys, vals, lens), rtol=1e-12, atol=1e-12)
    o2 = oracle.ebc_forward(specs, W2, keys, vals, lens)
    o12 = oracle.ebc_forward(specs, [2.0 * a - 3.0 * b for a, b in zip(W1, W2)], keys, vals, lens)
    torch.testing.assert_close(o12, 2.0 * o1 - 3.0 * o2, rtol=1e-10, atol=1e-10)
    for f in range(F):
        empty = lens[f * B:(f + 1) * B] == 0
        assert (o1[empty][:, 4 * f:4 * f + 4] == 0).all()
    G = torch.randn(B, 4 * F, generator=g, dtype=torch.float64)
    grads = oracle.ebc_dense_grads(specs, keys, vals, lens, G)
    lhs = float((o1 * G).sum())
    rhs = float(sum((w * gr).sum() for w, gr in zip(W1, grads)))
    assert abs(lhs - rhs) <= 1e-9 * (1.0 + abs(lhs))


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 30), st.integers(1, 6), st.integers(0, 40), st.integers(0, 2 ** 31 - 1))
def test_rowwise_optimizer_properties(R, D, n, seed):
    """Row-wise Adagrad: the sparse-exact form (unique rows + summed gradients) equals the dense form; untouched rows and
    their accumulators do not move; the accumulator never decreases.  Row-wise Adam: only touched rows advance, and a first
    step moves every touched element with a non-zero gradient by lr in magnitude (bias-corrected m / sqrt(v) with v = mean g^2
    per row is +-1 for D = 1)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, R, (n,), generator=g)
    contrib = torch.randn(n, D, generator=g, dtype=torch.float64)
    dense = torch.zeros(R, D, dtype=torch.float64).index_add_(0, ids, contrib)
    w0 = torch.randn(R, D, generator=g, dtype=torch.float64)
    s0 = torch.rand(R, generator=g, dtype=torch.float64)
    wd, sd = w0.clone(), s0.clone()
    oracle.rowwise_adagrad_dense(wd, sd, dense, lr=0.1)
    ws, ss = w0.clone(), s0.clone()
    rows = torch.unique(ids, sorted=True)
    oracle.rowwise_adagrad_sparse(ws, ss, rows, dense[rows], lr=0.1)
    torch.testing.assert_close(ws, wd, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(ss, sd, rtol=1e-12, atol=1e-12)
    untouched = torch.ones(R, dtype=torch.bool)
    untouched[rows] = False
    assert torch.equal(wd[untouched], w0[untouched]) and torch.equal(sd[untouched], s0[untouched])
    assert (sd >= s0).all()
    wa, m, v = w0.clone(), torch.zeros(R, D, dtype=torch.float64), torch.zeros(R, dtype=torch.float64)
    oracle.rowwise_adam_sparse(wa, m, v, rows, dense[rows], step=1, lr=0.05, eps=0.0 if D == 1 else 1e-8)
    assert torch.equal(wa[untouched], w0[untouched]) and (m[untouched] == 0).all() and (v[untouched] == 0).all()
    if D == 1 and rows.numel():
        nz = dense[rows].abs().squeeze(1) > 1e-9
        torch.testing.assert_close((wa[rows] - w0[rows]).abs().squeeze(1)[nz], torch.full((int(nz.sum()),), 0.05, dtype=torch.float64),
                                   rtol=1e-9, atol=1e-12)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 25), min_size=0, max_size=60))
def test_dedup_rows_properties(ids):
    """dedup_rows: keys ascending and unique, counts add up to n, unique[inverse] gives the input back."""
    t = torch.tensor(ids, dtype=torch.int64)
    u, inv, cnt = oracle.dedup_rows(t)
    assert u.tolist() == sorted(set(ids)) and int(cnt.sum()) == len(ids)
    assert torch.equal(u[inv], t)
    assert cnt.tolist() == [ids.count(int(x)) for x in u]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 6), st.integers(1, 12), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_retrieval_metrics_properties(Q, n_pred, k, seed):
    """precision / recall / ndcg at k lie in [0, 1]; retrieving exactly the targets first gives recall = ndcg = 1;
    retrieving nothing relevant gives 0 everywhere."""
    g = torch.Generator().manual_seed(seed)
    pred = [torch.randperm(40, generator=g)[:n_pred].tolist() for _ in range(Q)]
    tgt = [torch.randperm(40, generator=g)[:int(torch.randint(1, 8, (1,), generator=g))].tolist() for _ in range(Q)]
    m = oracle.retrieval_metrics(pred, tgt, k)
    assert all(0.0 <= x <= 1.0 + 1e-12 for x in m.values())
    perfect = [t + [100 + i for i in range(k)] for t in tgt]
    mp = oracle.retrieval_metrics(perfect, tgt, k)
    if all(len(t) <= k for t in tgt):
        assert abs(mp[f"recall_at_{k}"] - 1.0) < 1e-12
    assert abs(mp[f"ndcg_at_{k}"] - 1.0) < 1e-12
    m0 = oracle.retrieval_metrics([[1000 + i for i in range(n_pred)] for _ in range(Q)], tgt, k)
    assert all(x == 0.0 for x in m0.values())


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 24), st.integers(1, 8), st.floats(0.2, 3.0), st.integers(0, 2 ** 31 - 1))
def test_in_batch_softmax_properties(B, d, temp, seed):
    """In-batch softmax: the chunked form equals the full form; d loss / d S = softmax - onehot has zero row sums, so with
    IDENTICAL candidates the loss is log B, dQ = 0 and the candidates' gradients add up to zero (the closed form the GPU
    test checks at B = 65 536)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, d, generator=g, dtype=torch.float64, requires_grad=True)
    c = torch.randn(B, d, generator=g, dtype=torch.float64, requires_grad=True)
    l0, diag0 = oracle.in_batch_softmax_loss(q, c, temp)
    l1, diag1 = oracle.in_batch_softmax_loss_chunked(q, c, temp, chunk=5)
    torch.testing.assert_close(l1, l0, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(diag1, diag0, rtol=1e-12, atol=1e-12)
    l0.backward()
    # d loss / d S has zero row sums (softmax - onehot): dQ = (dS) C / T  =>  for c_j all equal to u, dQ = 0
    same = c.detach()[:1].expand(B, d).clone().requires_grad_(True)
    q2 = q.detach().clone().requires_grad_(True)
    l2, _ = oracle.in_batch_softmax_loss(q2, same, temp)
    l2.backward()
    assert abs(float(l2.detach()) - float(torch.log(torch.tensor(float(B), dtype=torch.float64)))) < 1e-10
    assert float(q2.grad.abs().max()) < 1e-12
    # the candidates' gradients add up to (mean_i q_i - mean_i q_i) = 0 when all candidates are equal
    assert float(same.grad.sum(dim=0).abs().max()) < 1e-10


@settings(max_examples=50, deadline=None)
@given(jagged(), st.data())
def test_kjt_container_matches_the_oracle(j, data):
    """The host-side KeyedJaggedTensor (CPU container path: offsets, per-key slicing, permute with repeats, split) against
    the oracle's integer ops on the same jagged data."""
    import two_tower_recommender_model_b200 as tt
    F, B, rows, vals, lens = j
    keys = [f"f{i}" for i in range(F)]
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(keys, vals, lens)
    offs = oracle.lengths_to_offsets(lens)
    assert kjt.stride() == B and torch.equal(kjt.offsets().to(torch.int64), offs.to(torch.int64))
    assert kjt.length_per_key() == [int(lens[f * B:(f + 1) * B].sum()) for f in range(F)]
    for f, k in enumerate(keys):
        jt = kjt[k]
        assert torch.equal(jt.values(), vals[int(offs[f * B]):int(offs[(f + 1) * B])])
        assert torch.equal(jt.lengths(), lens[f * B:(f + 1) * B])
    perm = data.draw(st.lists(st.integers(0, F - 1), min_size=1, max_size=F + 2))
    p = kjt.permute(perm)
    ol, ov, _ = oracle.permute_2d_sparse_data(perm, lens.view(F, B), vals)
    assert p.keys() == [keys[i] for i in perm]
    assert torch.equal(p.lengths(), ol.reshape(-1)) and torch.equal(p.values(), ov)
    cut = data.draw(st.integers(0, F))
    a, b = kjt.split([cut, F - cut])
    assert a.keys() == keys[:cut] and b.keys() == keys[cut:]
    assert torch.equal(torch.cat([a.values(), b.values()]), vals) and torch.equal(torch.cat([a.lengths(), b.lengths()]), lens)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 7), st.integers(1, 15), st.integers(1, 15), st.integers(0, 2 ** 31 - 1))
def test_product_retrieval_metrics_match_the_oracle(Q, n_pred, k, seed):
    """tt.retrieval_metrics (tensor form; torch ops only, so it runs here) against the oracle's per-row restatement of
    mlflow's retriever metrics (04_evaluate_retrieval.py:202-226): duplicates in the targets, k larger / smaller than the
    list, rows without a hit."""
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator().manual_seed(seed)
    pred = torch.stack([torch.randperm(30, generator=g)[:n_pred] for _ in range(Q)])
    tgt = []
    for _ in range(Q):
        t = torch.randint(0, 30, (int(torch.randint(1, 9, (1,), generator=g)),), generator=g).tolist()   # may repeat ids
        tgt.append(t)
    got = tt.retrieval_metrics(pred, tgt, k)
    want = oracle.retrieval_metrics(pred.tolist(), tgt, k)
    assert set(got) == set(want)
    for key in want:
        assert abs(got[key] - want[key]) < 1e-5, (key, got[key], want[key])
