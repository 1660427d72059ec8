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
        torch.testing.assert_close(got.cpu(), g, rtol=1e-4, atol=1e-4)   # hot rows: thousands of terms summed segment-wise, not in the oracle's order


@pytest.mark.parametrize("opt", ["adagrad", "adagrad_eps1e-8", "adam", "sgd"])
@pytest.mark.parametrize("dims,rows,pooling,B,L,dup", CASES[:5])
def test_fused_backward_optimizer(cuda, opt, dims, rows, pooling, B, L, dup):
    import two_tower_recommender_model_b200 as tt
    specs, ebc, weights = build(cuda, dims, rows, pooling)
    keys = [f"f{i}" for i in range(len(dims))]
    lr = 0.05
    # hot rows (dup pool: every row is hit by hundreds of ids) are summed segment by segment and combined with
    # red.global.add, not in the oracle's id order: the sums of thousands of O(1) terms differ by ~n*eps
    hot = 50.0 if dup else 1.0
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
            torch.testing.assert_close(p.detach().cpu(), w, rtol=2e-5 * hot, atol=2e-6 * hot)
    st = ebc.fused_optimizer_state()
    for i, s in enumerate(specs):
        if opt.startswith("adagrad"):
            torch.testing.assert_close(st[s.name]["sum"].cpu(), state[i]["sum"], rtol=2e-5 * hot, atol=1e-7 * hot)
        elif opt == "adam":
            torch.testing.assert_close(st[s.name]["exp_avg"].cpu(), state[i]["m"], rtol=2e-5 * hot, atol=1e-7 * hot * 2)
            torch.testing.assert_close(st[s.name]["exp_avg_sq"].cpu(), state[i]["v"], rtol=2e-5 * hot, atol=1e-9 * hot)


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


# ------------------------------------------------------------------ peer-memory entry points on ONE GPU
# tt_ebc_forward_peer / tt_ebc_backward_fused_peer only see pointers: three buffers on the same device stand in
# for the ranks of a box, which checks the (rank, local row) addressing without a multi-GPU machine
# (tests/test_gpu_multi.py runs the real thing over NVLink).
def _peer_struct(N, bufs, rows_per_peer, flags=0):
    pb = N.PeerBuffers()
    pb.world, pb.rows_per_peer, pb.flags = len(bufs), rows_per_peer, flags
    for i, b in enumerate(bufs):
        pb.ptr[i] = b.data_ptr()
    return pb


@pytest.mark.parametrize("dims,rows,pooling,Bl,L", [([64, 64], [2000, 500], ["sum", "sum"], 300, 1),
                                                     ([128, 36], [700, 90], ["mean", "sum"], 129, 6)])
def test_peer_forward_and_backward_virtual_ranks(cuda, dims, rows, pooling, Bl, L):
    from ctypes import byref
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N
    W = 3
    B = W * Bl                                            # global batch, rank s owns rows [s*Bl, (s+1)*Bl)
    specs, ebc, weights = build(cuda, dims, rows, pooling)
    keys = [f"f{i}" for i in range(len(dims))]
    v, l = random_kjt(keys, rows, B, L, seed=11 + L)
    want = oracle.ebc_forward(specs, weights, keys, v, l)
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda))
    vals, offs = kjt.values().contiguous(), kjt.offsets().to(torch.int32).contiguous()
    D = sum(dims)
    stride = D + 8                                        # wider buffer: columns start at 4
    layout = (stride, {k: 4 + sum(dims[:i]) for i, k in enumerate(keys)})
    bufs = [torch.full((Bl, stride), float("nan"), device=cuda) for _ in range(W)]
    plan, _ = ebc._build_plan(tuple(keys), B, with_state=False, out_layout=layout)
    pb = _peer_struct(N, bufs, Bl)
    N.call("tt_ebc_forward_peer", byref(plan), N.ptr(vals), N.ptr(offs), byref(pb), N.stream_ptr(cuda))
    got = torch.cat([b[:, 4:4 + D] for b in bufs]).cpu()
    torch.testing.assert_close(got, want, rtol=RTOL, atol=ATOL)
    assert all(bool(torch.isnan(b[:, :4]).all()) and bool(torch.isnan(b[:, 4 + D:]).all()) for b in bufs)   # nothing else touched

    # backward: gradients read from the virtual ranks == gradients read from one matrix (same kernel, same order: exact)
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    g = torch.randn(B, stride, device=cuda)
    gbufs = [g[s * Bl:(s + 1) * Bl].clone() for s in range(W)]

    def run(peer):
        for i, s in enumerate(specs):
            ebc.embedding_bags[s.name].weight.data.copy_(weights[i])
        ebc._fused_state.clear()
        spec = ebc._sparse_optimizer_spec(advance_step=True)
        plan, _ = ebc._build_plan(tuple(keys), B, with_state=True, out_layout=layout)
        n = vals.numel()
        ws = N.workspace(N.load().tt_ebc_backward_workspace_bytes(n), cuda)
        if peer:
            pg = _peer_struct(N, gbufs, Bl)
            N.call("tt_ebc_backward_fused_peer", byref(plan), byref(spec), N.ptr(vals), n, N.ptr(offs), byref(pg), N.ptr(ws), ws.numel(), N.stream_ptr(cuda))
        else:
            N.call("tt_ebc_backward_fused", byref(plan), byref(spec), N.ptr(vals), n, N.ptr(offs), N.ptr(g), N.ptr(ws), ws.numel(), N.stream_ptr(cuda))
        return [ebc.embedding_bags[s.name].weight.detach().clone() for s in specs]

    for a, b in zip(run(True), run(False)):
        assert torch.equal(a, b)


def test_peer_scatter_add_row_shards(cuda):
    """Row-wise sharding over peer memory on one GPU: two 'ranks' hold row ranges of the tables, each looks up the ids of
    ITS range for the global batch (tt_kjt_from_columns_range) and ADDS its rows into the zeroed per-rank buffers
    (TT_PEER_SCATTER_ADD).  The union must equal the unsharded lookup; ids 0 are empty bags, ids >= rows wrap."""
    from ctypes import byref
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N
    rows, D, Bl, W = [1001, 333], 64, 257, 2
    B = W * Bl
    specs, ebc, weights = build(cuda, [D, D], rows, ["sum", "mean"])
    keys = ["f0", "f1"]
    g = torch.Generator().manual_seed(5)
    ids = torch.stack([torch.randint(0, 2 * r, (B,), generator=g) for r in rows])
    ids[:, ::9] = 0
    # oracle on the reference's transform (id 0 -> empty, else id % rows)
    lens = (ids != 0).to(torch.int32).reshape(-1)
    vals = torch.cat([(ids[f][ids[f] != 0] % rows[f]) for f in range(2)])
    want = oracle.ebc_forward(specs, weights, keys, vals, lens)
    bufs = [torch.zeros(Bl, 2 * D, device=cuda) for _ in range(W)]
    pb = _peer_struct(N, bufs, Bl, N.TT_PEER_SCATTER_ADD)
    rows_dev = torch.tensor(rows, device=cuda)
    for s in range(W):                                    # shard s: rows [s*block, (s+1)*block) of every table
        block = [-(-r // W) for r in rows]
        lo = torch.tensor([s * b for b in block], device=cuda)
        hi = torch.tensor([min((s + 1) * b, r) for b, r in zip(block, rows)], device=cuda)
        kjt = tt.KeyedJaggedTensor.from_id_columns(keys, ids.to(cuda), rows_dev, row_range=(lo, hi))
        n_live = int(kjt.offsets()[-1])
        assert bool((kjt.values()[:n_live] >= 0).all()) and bool((kjt.values()[:n_live] < (hi - lo).max()).all())
        shard = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=sp.name, embedding_dim=D, num_embeddings=int(hi[i] - lo[i]),
                                                                        feature_names=list(sp.feature_names)) for i, sp in enumerate(specs)], device=cuda)
        for i, sp in enumerate(specs):
            shard.embedding_bags[sp.name].weight.data.copy_(weights[i][int(lo[i]):int(hi[i])])
        plan, _ = shard._build_plan(tuple(keys), B, with_state=False)
        N.call("tt_ebc_forward_peer", byref(plan), N.ptr(kjt.values()), N.ptr(kjt.offsets().to(torch.int32).contiguous()), byref(pb), N.stream_ptr(cuda))
    got = torch.cat(bufs).cpu()
    torch.testing.assert_close(got, want, rtol=RTOL, atol=ATOL)


def test_full_size_gather_and_update_checksum(cuda):
    """BASELINE configs[1] size: one 10M x 64 table, 65536 one-id bags.  Forward must equal a plain row gather
    (bit-exact); the fused SGD update with an all-ones gradient must lower the checksum of the table by exactly
    lr * (#ids) * D, and touch nothing but the looked-up rows."""
    import two_tower_recommender_model_b200 as tt
    R, D, B, lr = 10_000_000, 64, 65536, 0.5
    ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name="t", embedding_dim=D, num_embeddings=R, feature_names=["f"])], device=cuda)
    apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": lr})
    w = ebc.embedding_bags["t"].weight
    g = torch.Generator(device=cuda).manual_seed(9)
    ids = torch.randint(1, R, (1, B), device=cuda, generator=g)
    ids[0, :1000] = ids[0, 1000:2000]                                  # duplicates: their gradients must add up
    kjt = tt.KeyedJaggedTensor.from_id_columns(["f"], ids, torch.tensor([R]))
    before = w.detach().double().sum()
    rows_before = w.detach()[ids[0]].clone()
    out = ebc(kjt).values()
    assert torch.equal(out, rows_before)
    out.backward(torch.ones_like(out))
    after = w.detach().double().sum()
    assert abs(float(before - after) - lr * B * D) < 1e-6 * lr * B * D
    uniq, counts = torch.unique(ids[0], return_counts=True)
    torch.testing.assert_close(w.detach()[uniq], rows_before_unique(rows_before, ids[0], uniq) - lr * counts.unsqueeze(1).float(), rtol=0, atol=1e-6)


def rows_before_unique(rows_before, ids, uniq):
    first = torch.full((int(ids.max()) + 1,), -1, dtype=torch.long, device=ids.device)
    first[ids] = torch.arange(ids.numel(), device=ids.device)
    return rows_before[first[uniq]]


@pytest.mark.parametrize("rows,B,L,dup", [([2000, 500], 1024, 1, 0), ([50, 40], 4096, 4, 8), ([10_000_000, 10_000_000], 65536, 1, 0),
                                          ([100, 90, 7], 129, 3, 0), ([33], 5, 0, 0)])
def test_dedup_matches_torch_unique(cuda, rows, B, L, dup):
    """north_star lists dedup among the BIT-EXACT ops: the unique (table,row) keys, their counts and the inverse
    map produced by the fused backward's own key construction + radix sort (tt_ebc_dedup) must equal
    oracle.dedup_rows (= torch.unique(sorted, return_inverse, return_counts)) on the linearised keys."""
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200.functional import dedup_rows
    keys = [f"f{i}" for i in range(len(rows))]
    cfgs = [tt.EmbeddingBagConfig(name=f"t_f{i}", embedding_dim=4, num_embeddings=rows[i], feature_names=[keys[i]]) for i in range(len(rows))]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))        # no table memory needed: keys only
    v, l = random_kjt(keys, rows, B, L, seed=B + L + len(rows), dup_pool=dup)
    off = oracle.lengths_to_offsets(l).to(torch.int64)
    base, acc = [], 0
    for r in rows:
        base.append(acc)
        acc += r
    lin = torch.cat([v[int(off[f * B]):int(off[(f + 1) * B])] + base[f] for f in range(len(rows))]) if v.numel() else v
    want_u, want_inv, want_c = oracle.dedup_rows(lin)
    for bag in ebc.embedding_bags.values():          # _build_plan reads data_ptr(): give the meta tables a dummy allocation
        bag.weight = torch.nn.Parameter(torch.zeros(1, 4, device=cuda))
    got_u, got_inv, got_c = dedup_rows(ebc, tt.KeyedJaggedTensor.from_lengths_sync(keys, v.to(cuda), l.to(cuda)))
    assert torch.equal(got_u.cpu(), want_u) and torch.equal(got_c.cpu(), want_c) and torch.equal(got_inv.cpu(), want_inv)


@pytest.mark.parametrize("W,Bl,rows,D", [(8, 256, 1777, 64), (8, 256, 3001, 64), (4, 256, 1777, 64), (8, 512, 90, 64), (8, 300, 50000, 128)])
def test_table_wise_owner_at_world_8_virtual_ranks(cuda, W, Bl, rows, D):
    """What the owner of ONE table does in an 8-rank table-wise job, on one GPU: the key-major KJT over the GLOBAL batch
    (W * Bl bags, built on the device from id columns, so `values` is over-allocated and the live count sits in the
    offsets) -> fused backward reading every sample's gradient row from the buffer of the rank that owns the sample
    (W virtual peers) -> row-wise Adagrad with grad_scale 1/W.  Three steps against the oracle's dense formula.
    (bench.py's N = 8 parity check once reported 8e-4 on the gathered tables where N = 2 / 4 gave 3e-8.)"""
    from ctypes import byref
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N
    B = W * Bl
    specs, ebc, weights = build(cuda, [D], [rows], ["sum"])
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    ebc._grad_scale = 1.0 / W
    w_ref = weights[0].clone()
    state = torch.zeros(rows)
    rows_dev = torch.tensor([rows], device=cuda)
    for step in range(3):
        g = torch.Generator().manual_seed(100 * step + W + rows)
        ids = torch.randint(0 if step == 1 else 1, 2 * rows, (1, B), generator=g)      # step 1: id 0 -> empty bags too
        go = torch.randn(B, D, generator=g)
        kjt = tt.KeyedJaggedTensor.from_id_columns(["f0"], ids.to(cuda), rows_dev)
        vals, offs = kjt.values().contiguous(), kjt.offsets().to(torch.int32).contiguous()
        gbufs = [go[s * Bl:(s + 1) * Bl].clone().to(cuda) for s in range(W)]
        spec = ebc._sparse_optimizer_spec(advance_step=True)
        plan, _ = ebc._build_plan(("f0",), B, with_state=True)
        n = vals.numel()
        ws = N.workspace(N.load().tt_ebc_backward_workspace_bytes(n), cuda)
        pg = _peer_struct(N, gbufs, Bl)
        N.call("tt_ebc_backward_fused_peer", byref(plan), byref(spec), N.ptr(vals), n, N.ptr(offs), byref(pg), N.ptr(ws), ws.numel(), N.stream_ptr(cuda))
        # oracle: dense gradient of the table / W, row-wise Adagrad
        v, l, _ = oracle.transform_to_torchrec_batch({"f0": ids[0].tolist(), "label": [0] * B}, ["f0"], [rows])
        dense = oracle.ebc_dense_grads(specs, ["f0"], v, l, go)[0] / W
        oracle.rowwise_adagrad_dense(w_ref, state, dense, lr=0.05, eps=1e-10)
        got = ebc.embedding_bags["t_f0"].weight.detach().cpu()
        torch.testing.assert_close(got, w_ref, rtol=2e-5, atol=2e-6, msg=lambda m: f"step {step}: {m}")


@pytest.mark.parametrize("B", [256, 1024, 2048])
@pytest.mark.parametrize("mode", ["dense_grad", "adagrad"])
def test_short_runs_across_segment_boundaries(cuda, B, mode):
    """One id per bag, two small tables (the shape of bench.py's sharded parity check): with 2*B keys over 3001 + 1777 rows
    many rows are hit 2-7 times and some of those short runs straddle a multiple of 128 in the sorted key array, so
    their sum is formed by two groups through the hot-row scratch.  Dense gradients and the fused Adagrad update must
    still equal the oracle's to rounding."""
    import two_tower_recommender_model_b200 as tt
    rows, D = [3001, 1777], 64
    specs, ebc, weights = build(cuda, [D, D], rows, ["sum", "sum"])
    keys = ["f0", "f1"]
    if mode == "adagrad":
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    state = [torch.zeros(r) for r in rows]
    rows_dev = torch.tensor(rows, device=cuda)
    for step in range(3):
        g = torch.Generator().manual_seed(7000 + 100 * step + B)
        ids = torch.stack([torch.randint(1, r, (B,), generator=g) for r in rows])
        go = torch.randn(B, 2 * D, generator=g)
        v, l, _ = oracle.transform_to_torchrec_batch({"f0": ids[0].tolist(), "f1": ids[1].tolist(), "label": [0] * B}, keys, rows)
        want = oracle.ebc_dense_grads(specs, keys, v, l, go)
        for p in ebc.parameters():
            p.grad = None
        kt = ebc(tt.KeyedJaggedTensor.from_id_columns(keys, ids.to(cuda), rows_dev))
        kt.values().backward(go.to(cuda))
        for i, s in enumerate(specs):
            p = ebc.embedding_bags[s.name].weight
            if mode == "dense_grad":
                torch.testing.assert_close(p.grad.cpu(), want[i], rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} table {i}: {m}")
            else:
                oracle.rowwise_adagrad_dense(weights[i], state[i], want[i], lr=0.05, eps=1e-10)
                torch.testing.assert_close(p.detach().cpu(), weights[i], rtol=2e-5, atol=2e-6, msg=lambda m: f"step {step} table {i}: {m}")
