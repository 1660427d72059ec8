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
