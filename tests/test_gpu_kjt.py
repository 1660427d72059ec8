"""CUDA integer ops vs the oracle: bit-exact."""
import pytest
import torch

import oracle
from oracle.kjt import block_bucketize_vectorized
from helpers import random_kjt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 5, 1023, 1024, 4097, 32768, 32769, 300000, 2_000_003])
def test_lengths_to_offsets(cuda, n):
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator().manual_seed(n)
    lengths = torch.randint(0, 21, (n,), generator=g, dtype=torch.int32)
    want = oracle.lengths_to_offsets(lengths)
    kjt = tt.KeyedJaggedTensor(keys=["a"], values=torch.zeros(int(lengths.sum()), dtype=torch.int64, device=cuda),
                               lengths=lengths.to(cuda))
    got = kjt.offsets()
    assert got.dtype == torch.int32 and torch.equal(got.cpu(), want)


def test_from_id_columns_matches_reference_loop(cuda):
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator().manual_seed(7)
    B = 5000
    cols = {"user_id": torch.randint(-50, 400, (B,), generator=g), "product_id": torch.randint(0, 3, (B,), generator=g) * 977,
            "label": torch.randint(0, 2, (B,), generator=g)}
    emb = [193, 9740]
    v, l, _ = oracle.transform_to_torchrec_batch({k: t.tolist() for k, t in cols.items()}, ["user_id", "product_id"], emb)
    ids = torch.stack([cols["user_id"], cols["product_id"]]).to(cuda)
    kjt = tt.KeyedJaggedTensor.from_id_columns(["user_id", "product_id"], ids, torch.tensor(emb))
    assert torch.equal(kjt.lengths().cpu(), l)
    n = int(kjt.offsets()[-1])
    assert n == v.numel() and torch.equal(kjt.values()[:n].cpu(), v)


@pytest.mark.parametrize("F,B,L,perm", [(3, 7, 3, [2, 0, 1]), (4, 1000, 5, [3, 3, 0]), (2, 65536, 1, [1, 0]), (5, 33, 0, [4, 1])])
def test_permute_2d(cuda, F, B, L, perm):
    import two_tower_recommender_model_b200 as tt
    keys = [f"f{i}" for i in range(F)]
    v, l = random_kjt(keys, [1000] * F, B, L, seed=F * 131 + B)
    ol, ov, _ = oracle.permute_2d_sparse_data(perm, l.view(F, B), v)
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda))
    out = kjt.permute(perm)
    assert out.keys() == [keys[i] for i in perm]
    assert torch.equal(out.lengths().cpu(), ol.reshape(-1)) and torch.equal(out.values().cpu(), ov)
    assert torch.equal(out.offsets().cpu(), oracle.lengths_to_offsets(ol.reshape(-1)))
    # permute o inverse permute = identity (for true permutations)
    if sorted(perm) == list(range(F)):
        inv = [perm.index(i) for i in range(F)]
        back = out.permute(inv)
        assert torch.equal(back.values().cpu(), v) and torch.equal(back.lengths().cpu(), l)


@pytest.mark.parametrize("F,B,L,W,rows", [(1, 9, 4, 2, [10]), (3, 257, 6, 4, [1000, 7, 123457]), (2, 4096, 20, 8, [100_000_000, 50]), (2, 50, 0, 3, [5, 5])])
def test_block_bucketize(cuda, F, B, L, W, rows):
    from two_tower_recommender_model_b200.functional import block_bucketize
    keys = [f"f{i}" for i in range(F)]
    v, l = random_kjt(keys, rows, B, L, seed=W * 17 + B)
    want = block_bucketize_vectorized(l, v, rows, W, B)
    if v.numel() <= 5000:
        loop = oracle.block_bucketize_sparse_features(l, v, rows, W, B)
        assert all(torch.equal(a, b) for a, b in zip(want, loop))
    off = oracle.lengths_to_offsets(l)
    nl, no, nv, unb = block_bucketize(l.to(cuda), off.to(cuda), v.to(cuda), torch.tensor(rows), F, B, W)
    assert torch.equal(nl.cpu(), want[0]) and torch.equal(nv.cpu(), want[1]) and torch.equal(unb.cpu(), want[2])
    assert torch.equal(no.cpu(), oracle.lengths_to_offsets(want[0]))
    # un-bucketise: local id + bucket*block == original id
    blocks = torch.tensor([-(-r // W) for r in rows])
    bag = torch.repeat_interleave(torch.arange(W * F * B), want[0].long())
    w_of = bag // (F * B)
    f_of = (bag // B) % F
    assert torch.equal((nv.cpu() + w_of * blocks[f_of])[unb.cpu()], v)


def test_block_bucketize_out_of_range_ids(cuda):
    """Ids outside [0, block*W) (a KJT built without the reference's modulo, negative ids) take fbgemm's fallback
    bucket = id % W, local = id / W on unsigned ids: no out-of-bounds write, same answer as the oracle."""
    from two_tower_recommender_model_b200.functional import block_bucketize
    F, B, W, rows = 2, 300, 4, [1000, 37]
    v, l = random_kjt(["a", "b"], rows, B, 5, seed=3)
    g = torch.Generator().manual_seed(4)
    bad = torch.randperm(v.numel(), generator=g)[: v.numel() // 3]
    v[bad] = torch.randint(-5000, 5000, (bad.numel(),), generator=g)
    v[bad[:4]] = torch.tensor([-1, -(2 ** 63), 2 ** 63 - 1, 4000])
    want = block_bucketize_vectorized(l, v, rows, W, B)
    loop = oracle.block_bucketize_sparse_features(l, v, rows, W, B)
    assert all(torch.equal(a, b) for a, b in zip(want, loop))
    off = oracle.lengths_to_offsets(l)
    guard = torch.full((W * F * B + 64,), -7, dtype=torch.int32, device=cuda)     # nothing may land past new_lengths
    nl, no, nv, unb = block_bucketize(l.to(cuda), off.to(cuda), v.to(cuda), torch.tensor(rows), F, B, W)
    assert torch.equal(nl.cpu(), want[0]) and torch.equal(nv.cpu(), want[1]) and torch.equal(unb.cpu(), want[2])
    assert int(nl.sum()) == v.numel() and bool((guard == -7).all())


@pytest.mark.parametrize("F,B,W,rows", [(2, 1000, 4, [1500, 900]), (1, 9, 2, [10]), (3, 4097, 8, [100_000_000, 50, 12345])])
def test_from_id_columns_range_is_one_bucket_of_block_bucketize(cuda, F, B, W, rows):
    """tt_kjt_from_columns_range (row-wise shard of a dense id-column batch) == the reference transform
    (utils/model_training.py:43-61) followed by bucket `w` of fbgemm::block_bucketize_sparse_features."""
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator().manual_seed(F * B + W)
    ids = torch.stack([torch.randint(0, 3 * r, (B,), generator=g) for r in rows])
    ids[:, ::5] = 0
    raw = {f"f{i}": ids[i].tolist() for i in range(F)}
    raw["label"] = [0] * B
    keys = [f"f{i}" for i in range(F)]
    v, l, _ = oracle.transform_to_torchrec_batch(raw, keys, rows)
    nl, nv, _ = block_bucketize_vectorized(l, v, rows, W, B)
    noff = oracle.lengths_to_offsets(nl).long()
    rows_dev = torch.tensor(rows, device=cuda)
    for w in range(W):
        blocks = [-(-r // W) for r in rows]
        lo = torch.tensor([w * b for b in blocks], device=cuda)
        hi = torch.tensor([min((w + 1) * b, r) for b, r in zip(blocks, rows)], device=cuda)
        kjt = tt.KeyedJaggedTensor.from_id_columns(keys, ids.to(cuda), rows_dev, row_range=(lo, hi))
        want_l = nl[w * F * B:(w + 1) * F * B]
        want_v = nv[int(noff[w * F * B]):int(noff[(w + 1) * F * B])]
        assert torch.equal(kjt.lengths().cpu(), want_l)
        assert torch.equal(kjt.offsets().cpu().long(), oracle.lengths_to_offsets(want_l).long())
        assert torch.equal(kjt.values()[:want_v.numel()].cpu(), want_v)


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 5), (2048, 16), (2049, 25), (131072, 25), (1_300_000, 28), (70000, 32)])
def test_radix_sort_stable(cuda, n, bits):
    from two_tower_recommender_model_b200.functional import sort_pairs
    g = torch.Generator().manual_seed(n)
    hi = (1 << bits) - 1 if bits < 32 else (1 << 31) - 1
    keys = torch.randint(0, min(hi, max(n // 3, 1)) + 1, (n,), generator=g, dtype=torch.int64)
    if bits == 32:
        keys[::3] += (1 << 31)  # exercise the top bit
    vals = torch.arange(n, dtype=torch.int64)
    order = torch.sort(keys, stable=True).indices
    k32 = (keys & 0xFFFFFFFF).to(torch.int64)
    k_dev = torch.where(k32 >= (1 << 31), k32 - (1 << 32), k32).to(torch.int32).to(cuda)
    ko, vo = sort_pairs(k_dev, vals.to(torch.int32).to(cuda), bits)
    assert torch.equal(vo.cpu().long(), order)
    assert torch.equal((ko.cpu().long() & 0xFFFFFFFF), keys[order])


@pytest.mark.parametrize("W,F,B,L,rows", [(2, 1, 2, 3, [10]), (4, 3, 257, 6, [1000, 7, 123457]), (8, 2, 1024, 20, [100_000_000, 50]), (3, 2, 50, 0, [5, 5])])
def test_gathered_range_matches_oracle(cuda, W, F, B, L, rows):
    """tt_kjt_gathered_range (sync-free row-wise / table-wise input dist for multi-hot KJTs) against the oracle's filter,
    bit-exact, for every rank's row range, with the gathered values padded to a common capacity."""
    from two_tower_recommender_model_b200.functional import kjt_gathered_range
    keys = [f"f{i}" for i in range(F)]
    per_rank = [random_kjt(keys, rows, B, L, seed=W * 31 + r) for r in range(W)]
    cap = max(int(v.numel()) for v, _ in per_rank) + 5
    g_vals = torch.full((W, cap), -77, dtype=torch.int64)
    g_offs = torch.zeros(W, F * B + 1, dtype=torch.int32)
    for r, (v, l) in enumerate(per_rank):
        g_vals[r, :v.numel()] = v
        g_offs[r] = oracle.lengths_to_offsets(l).to(torch.int32)
    for w in ([0, W - 1] if W > 2 else range(W)):
        block = [-(-r // W) for r in rows]
        lo = [w * b for b in block]
        hi = [min((w + 1) * b, r) for b, r in zip(block, rows)]
        if w == 0 and F > 1:
            lo[1], hi[1] = 0, rows[1]          # a table-wise feature owned by this rank: all of its rows
        want_v, want_l = oracle.gathered_range_shard([p[0] for p in per_rank], [p[1] for p in per_rank], lo, hi, B) \
            if g_vals.numel() < 200_000 else _vectorised_shard(per_rank, lo, hi, B)
        ov, ol, oo = kjt_gathered_range(g_vals.to(cuda).view(-1), cap, g_offs.to(cuda).view(-1), torch.tensor(lo, device=cuda),
                                        torch.tensor(hi, device=cuda), W, F, B)
        assert torch.equal(ol.cpu(), want_l)
        assert torch.equal(oo.cpu(), oracle.lengths_to_offsets(want_l).to(torch.int32))
        assert torch.equal(ov.cpu()[:want_v.numel()], want_v)


def _vectorised_shard(per_rank, lo, hi, B):
    """The oracle's filter, vectorised for the large case (checked equal to the loop on the small ones above)."""
    W, F = len(per_rank), len(lo)
    vals, lens = [], []
    for f in range(F):
        for r in range(W):
            v, l = per_rank[r]
            off = oracle.lengths_to_offsets(l).long()
            seg = v[int(off[f * B]):int(off[(f + 1) * B])]
            bag = torch.repeat_interleave(torch.arange(B), l[f * B:(f + 1) * B].long())
            keep = (seg >= lo[f]) & (seg < hi[f])
            vals.append(seg[keep] - lo[f])
            lens.append(torch.bincount(bag[keep], minlength=B).to(torch.int32))
    return torch.cat(vals), torch.cat(lens)
