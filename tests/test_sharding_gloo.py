"""CPU, world_size 2, gloo: the host-side logic of table-wise / row-wise sharding -- routing,
split sizes, the (source rank, feature) -> key-major permutation, output assembly, the reverse
exchanges in backward, ShardedTensor state dicts.  The device work is replaced by the oracle
(tests only): sharded(W=2) must equal the unsharded oracle on the same global batch."""
import os
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle.ebc import TableSpec  # noqa: E402
from oracle.kjt import block_bucketize_vectorized  # noqa: E402


# ------------------------------------------------------------------ oracle-backed stand-ins for the CUDA work
class _OracleLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, keys, values, lengths, *weights):
        ctx.mod, ctx.keys = mod, keys
        ctx.save_for_backward(values, lengths)
        return oracle.ebc_forward(mod.specs, [w.detach() for w in weights], keys, values, lengths)

    @staticmethod
    def backward(ctx, g):
        values, lengths = ctx.saved_tensors
        mod = ctx.mod
        grads = oracle.ebc_dense_grads(mod.specs, ctx.keys, values, lengths, g * getattr(mod, "_grad_scale", 1.0))
        if not all(hasattr(mod.embedding_bags[s.name].weight, "_optimizer_kwargs") for s in mod.specs):
            return (None, None, None, None) + tuple(grads)      # no in-backward optimizer: dense [R, D] gradients (OPT_DENSE_GRAD)
        for s, gr in zip(mod.specs, grads):  # "fused": row-wise Adagrad applied in backward
            w = mod.embedding_bags[s.name].weight
            oracle.rowwise_adagrad_dense(w.data, mod.state[s.name], gr, lr=w._optimizer_kwargs[0]["lr"])
        return (None, None, None, None) + (None,) * len(mod.specs)


class _Holder(nn.Module):
    def __init__(self, rows, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(rows, dim))


class OracleLocalEbc(nn.Module):
    def __init__(self, tables, device):
        super().__init__()
        import two_tower_recommender_model_b200 as tt
        self.specs = [TableSpec(c.name, c.num_embeddings, c.embedding_dim, list(c.feature_names),
                                "mean" if c.pooling == tt.PoolingType.MEAN else "sum") for c in tables]
        self.embedding_bags = nn.ModuleDict({s.name: _Holder(s.num_embeddings, s.embedding_dim) for s in self.specs})
        self.state = {s.name: torch.zeros(s.num_embeddings) for s in self.specs}
        self._features = [f for s in self.specs for f in s.feature_names]
        self._dims = [s.embedding_dim for s in self.specs for _ in s.feature_names]
        self._cfgs = list(tables)

    # what the sharded module's optimizer-state checkpoint asks of a local collection
    def embedding_bag_configs(self):
        return self._cfgs

    def _in_backward_kind(self):
        from two_tower_recommender_model_b200 import _native as N
        tagged = all(hasattr(self.embedding_bags[s.name].weight, "_optimizer_kwargs") for s in self.specs)
        return N.OPT_ROWWISE_ADAGRAD if tagged else None      # the stand-in applies row-wise Adagrad whatever the tag says

    def _state_for(self, cfg, w, kind):
        return {"sum": self.state[cfg.name]}

    def forward(self, kjt):
        import two_tower_recommender_model_b200 as tt
        w = [self.embedding_bags[s.name].weight for s in self.specs]
        out = _OracleLookup.apply(self, list(kjt.keys()), kjt.values(), kjt.lengths(), *w)
        return tt.KeyedTensor(self._features, self._dims, out)


def oracle_gather_range(g_vals, cap, g_offs, lo, hi, W, F, B):
    """Stand-in for tt_kjt_gathered_range: the oracle's filter over the W gathered (padded) KJTs."""
    offs = g_offs.view(W, F * B + 1)
    lens = [(offs[r, 1:] - offs[r, :-1]).to(torch.int32) for r in range(W)]
    vals = [g_vals.view(W, cap)[r, :int(offs[r, -1])] for r in range(W)]
    v, l = oracle.gathered_range_shard(vals, lens, lo.tolist(), hi.tolist(), B)
    out_v = torch.zeros(W * cap, dtype=torch.int64)
    out_v[:v.numel()] = v
    return out_v, l, oracle.lengths_to_offsets(l).to(torch.int32)


def oracle_bucketize(lengths, offsets, values, rows, F, B, W):
    nl, nv, unb = block_bucketize_vectorized(lengths, values, rows.tolist(), W, B)
    return nl, oracle.lengths_to_offsets(nl), nv, unb


# ------------------------------------------------------------------ worker
SPECS = [TableSpec("t_a", 40, 8, ["a"], "sum"), TableSpec("t_b", 30, 8, ["b1", "b2"], "sum"),
         TableSpec("t_c", 101, 4, ["c"], "mean"), TableSpec("t_d", 7, 4, ["d"], "sum")]
KEYS = ["c", "a", "b2", "d", "b1"]          # KJT key order differs from table order on purpose
B, LR = 6, 0.1


def _batch(rank):
    from helpers import random_kjt
    rows = {"a": 40, "b1": 30, "b2": 30, "c": 101, "d": 7}
    return random_kjt(KEYS, [rows[k] for k in KEYS], B, 4, seed=100 + rank)


def _worker(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

        g = torch.Generator().manual_seed(7)
        full = {s.name: torch.randn(s.num_embeddings, s.embedding_dim, generator=g) for s in SPECS}
        cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=s.embedding_dim, num_embeddings=s.num_embeddings,
                                      feature_names=list(s.feature_names),
                                      pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in SPECS]
        ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": LR})
        holder = nn.ModuleDict({"ebc": ebc})
        planner = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                              constraints={"t_c": ParameterConstraints(sharding_types=["row_wise"]),
                                                           "t_d": ParameterConstraints(sharding_types=["row_wise"])})
        plan = planner.collective_plan(holder, tt.get_default_sharders(), dist.GroupMember.WORLD)
        p = plan.plan["ebc"]
        assert p["t_a"].sharding_type == "table_wise" and p["t_b"].sharding_type == "table_wise"
        assert {p["t_a"].ranks[0], p["t_b"].ranks[0]} == {0, 1}
        assert p["t_c"].sharding_type == "row_wise" and p["t_c"].block_size == 51 and p["t_d"].block_size == 4
        model = tt.DistributedModelParallel(module=holder, device=torch.device("cpu"), plan=plan,
                                            sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize,
                                                                 gather_range_fn=oracle_gather_range))
        assert model._plan is plan and "t_c" in str(model._plan)
        sharded = model.module["ebc"]
        assert sharded.tw_ebc._grad_scale == 0.5 and sharded.rw_ebc._grad_scale == 0.5
        sharded.load_state_dict({f"embedding_bags.{k}.weight": v for k, v in full.items()})

        # ---- forward: every rank's output == unsharded lookup of ITS batch
        values, lengths = _batch(rank)
        kjt = tt.KeyedJaggedTensor.from_lengths_sync(KEYS, values, lengths)
        kt = sharded(kjt)
        want = oracle.ebc_forward(SPECS, [full[s.name] for s in SPECS], KEYS, values, lengths)
        assert kt.keys() == ["a", "b1", "b2", "c", "d"]
        torch.testing.assert_close(kt.values(), want, rtol=1e-6, atol=1e-6)

        # ---- backward: weights after the fused update == unsharded update with the GLOBAL batch's gradient
        gout = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(50 + rank))
        (kt.values() * gout).sum().backward()
        ref = {k: v.clone() for k, v in full.items()}
        dense = [torch.zeros_like(ref[s.name]) for s in SPECS]
        for r in range(world):
            v_r, l_r = _batch(r)
            g_r = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(50 + r))
            for acc, gr in zip(dense, oracle.ebc_dense_grads(SPECS, KEYS, v_r, l_r, g_r)):
                acc += gr
        for s, gr in zip(SPECS, dense):
            # TorchRec's gradient division: the tables see sum_r grad_r / W
            oracle.rowwise_adagrad_dense(ref[s.name], torch.zeros(s.num_embeddings), gr / world, lr=LR)

        # ---- state dict: ShardedTensor per table, gathered as utils/model_training.py:161-182 does
        sd = model.state_dict()
        assert set(sd) == {f"ebc.embedding_bags.{s.name}.weight" for s in SPECS}
        for s in SPECS:
            t = sd[f"ebc.embedding_bags.{s.name}.weight"]
            assert isinstance(t, ShardedTensor) and tuple(t.size()) == (s.num_embeddings, s.embedding_dim)
            full_t = torch.zeros(t.size()) if rank == 0 else None
            t.gather(0, full_t)
            if rank == 0:
                torch.testing.assert_close(full_t, ref[s.name], rtol=1e-5, atol=1e-6, msg=lambda m: f"{s.name}: {m}")

        # ---- the pipeline's prefetch hook produces the same result
        v2, l2 = _batch(rank + 10)
        kjt2 = tt.KeyedJaggedTensor.from_lengths_sync(KEYS, v2, l2)
        batch = tt.Batch(torch.zeros(1), kjt2, torch.zeros(B, dtype=torch.int32))
        model.start_sparse_data_dist(batch, None)
        with torch.no_grad():
            kt2 = sharded(kjt2)
        cur = {}
        for s in SPECS:
            t = model.state_dict()[f"ebc.embedding_bags.{s.name}.weight"]
            ft = torch.zeros(t.size())
            outs = [torch.zeros(t.size()) for _ in range(world)] if False else None
            t.gather(0, ft if rank == 0 else None)
            lst = [ft]
            dist.broadcast_object_list(lst, src=0)
            cur[s.name] = lst[0]
        want2 = oracle.ebc_forward(SPECS, [cur[s.name] for s in SPECS], KEYS, v2, l2)
        torch.testing.assert_close(kt2.values(), want2, rtol=1e-5, atol=1e-6)

        # ---- sync-free input dist for fixed-capacity (padded) multi-hot KJTs: all-gather + row-range filter.  What each
        # rank ends up holding == the oracle's filter with lo / hi worked out from the plan by hand.
        cap = 5 * B * 4 + 3          # >= F * B * L ids, the same on every rank
        per_rank = [_batch(r + 20) for r in range(world)]
        v3, l3 = per_rank[rank]
        padded = torch.full((cap,), -5, dtype=torch.int64)
        padded[:v3.numel()] = v3
        kjt3 = tt.KeyedJaggedTensor(keys=KEYS, values=padded, lengths=l3)
        kjt3._values_padded = True
        owner = {"a": p["t_a"].ranks[0], "b1": p["t_b"].ranks[0], "b2": p["t_b"].ranks[0]}
        rows = {"a": 40, "b1": 30, "b2": 30, "c": 101, "d": 7}
        for grp, want_lohi in ((sharded._rw, {"c": (51 * rank, min(51 * (rank + 1), 101)), "d": (4 * rank, min(4 * rank + 4, 7))}),
                               (sharded._tw, {k: (0, rows[k] if owner[k] == rank else 0) for k in owner})):
            got = sharded._dist_kjt_gather(grp, kjt3, B)
            lo = [want_lohi.get(k, (0, 0))[0] for k in KEYS]
            hi = [want_lohi.get(k, (0, 0))[1] for k in KEYS]
            wv, wl = oracle.gathered_range_shard([q[0] for q in per_rank], [q[1] for q in per_rank], lo, hi, B)
            assert got.keys() == KEYS and got.stride() == world * B
            assert torch.equal(got.lengths(), wl) and torch.equal(got.values()[:wv.numel()], wv)
        sharded._gathered = None

        # ---- dense gradient sync: one all-reduce, mean over ranks
        from two_tower_recommender_model_b200.distributed.sharding import DenseGradSync
        lin = nn.Linear(3, 2)
        sync = DenseGradSync(nn.ModuleDict({"l": lin}), None)
        w0 = lin.weight.detach().clone()
        lst = [w0]
        dist.broadcast_object_list(lst, src=0)
        assert torch.equal(lst[0], w0)  # broadcast from rank 0 made them identical
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
        sync.all_reduce()
        assert torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)) and torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 15.0))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def test_sharded_equals_unsharded_world2_gloo():
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ corpus-sharded retrieval (SURVEY 8(e) fallback)
def _worker_corpus_sharded(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from two_tower_recommender_model_b200 import retrieval
        # the device kernel -> oracle (tests only): same contract, global ids through item_index_base
        def topk(q, items, k, item_index_base=0, precision="fp32", items_bf16=None):
            s, i = oracle.exact_topk(q, items, k)
            return s, i + item_index_base
        retrieval.score_topk = topk
        g = torch.Generator().manual_seed(3)
        n_items, d, Q = 157, 8, 11
        # entries on a 1/4 grid: every dot product is exact in fp32 and TIES ACROSS SHARDS are common
        items = torch.randint(-4, 5, (n_items, d), generator=g).float() / 4
        items[100] = items[3]                       # the same vector in shard 0 and shard 1: a guaranteed cross-shard tie
        queries = [torch.randint(-4, 5, (Q, d), generator=g).float() / 4 for _ in range(world)]
        per = -(-n_items // world)
        lo, hi = rank * per, min((rank + 1) * per, n_items)
        for k in (1, 10, 100, 200):                 # 100 > one shard's share of the top, 200 > the whole corpus
            index = tt.CorpusShardedIndex(items[lo:hi], first_id=lo)
            s, i = index.search(queries[rank], k)
            ws, wi = oracle.exact_topk(queries[rank], items, min(k, n_items))
            kk = min(k, n_items)
            assert s.shape == (Q, k) and i.shape == (Q, k)
            assert torch.equal(s[:, :kk], ws) and torch.equal(i[:, :kk], wi), f"k={k}"
            if k > n_items:                         # more results asked for than items exist: the tail is padding
                assert torch.isinf(s[:, kk:]).all()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def test_corpus_sharded_retrieval_world2_gloo():
    """Each rank holds half of the corpus, the queries are all-gathered, per-shard top-k lists return to the query's rank in
    one all-to-all and are merged: ids and scores bit-exact against the oracle's top-k over the whole corpus, ties across
    shards resolved to the lower id."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_worker_corpus_sharded, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ column-wise sharding (the knob that spreads a table-wise owner's work)
CW_SPECS = [TableSpec("t_a", 40, 8, ["a"], "sum"), TableSpec("t_e", 20, 8, ["e1", "e2"], "mean"), TableSpec("t_c", 33, 4, ["c"], "sum")]
CW_KEYS = ["e2", "c", "a", "e1"]


def _cw_batch(rank):
    from helpers import random_kjt
    rows = {"a": 40, "e1": 20, "e2": 20, "c": 33}
    return random_kjt(CW_KEYS, [rows[k] for k in CW_KEYS], B, 3, seed=300 + rank)


def _worker_column_wise(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

        g = torch.Generator().manual_seed(11)
        full = {s.name: torch.randn(s.num_embeddings, s.embedding_dim, generator=g) for s in CW_SPECS}
        cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=s.embedding_dim, num_embeddings=s.num_embeddings, feature_names=list(s.feature_names),
                                      pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in CW_SPECS]
        ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": LR})
        holder = nn.ModuleDict({"ebc": ebc})
        planner = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                              constraints={"t_e": ParameterConstraints(sharding_types=["column_wise"]),
                                                           "t_a": ParameterConstraints(sharding_types=["table_wise"]),
                                                           "t_c": ParameterConstraints(sharding_types=["row_wise"])})
        plan = planner.collective_plan(holder, tt.get_default_sharders(), dist.GroupMember.WORLD)
        p = plan.plan["ebc"]
        assert p["t_e"].sharding_type == "column_wise" and p["t_e"].ranks == [0, 1]
        model = tt.DistributedModelParallel(module=holder, device=torch.device("cpu"), plan=plan,
                                            sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize))
        sharded = model.module["ebc"]
        assert sharded.shard_info()["t_e"] == ("column_wise", 0, 20) and sharded._col_info["t_e"] == (4 * rank, 4)
        # load from FULL tensors: every rank keeps its columns
        sharded.load_state_dict({f"embedding_bags.{k}.weight": v for k, v in full.items()})
        torch.testing.assert_close(sharded.tw_ebc.embedding_bags["t_e"].weight.detach(), full["t_e"][:, 4 * rank:4 * rank + 4])

        values, lengths = _cw_batch(rank)
        kjt = tt.KeyedJaggedTensor.from_lengths_sync(CW_KEYS, values, lengths)
        kt = sharded(kjt)
        want = oracle.ebc_forward(CW_SPECS, [full[s.name] for s in CW_SPECS], CW_KEYS, values, lengths)
        assert kt.keys() == ["a", "e1", "e2", "c"] and kt.length_per_key() == [8, 8, 8, 4]
        torch.testing.assert_close(kt.values(), want, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(kt["e2"], want[:, 16:24], rtol=1e-6, atol=1e-6)

        # backward: the tables see sum_r grad_r / W; a column-wise table applies row-wise Adagrad PER COLUMN SHARD
        gout = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(70 + rank))
        (kt.values() * gout).sum().backward()
        ref = {k: v.clone() for k, v in full.items()}
        dense = [torch.zeros_like(ref[s.name]) for s in CW_SPECS]
        for r in range(world):
            v_r, l_r = _cw_batch(r)
            g_r = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(70 + r))
            for acc, gr in zip(dense, oracle.ebc_dense_grads(CW_SPECS, CW_KEYS, v_r, l_r, g_r)):
                acc += gr
        for s, gr in zip(CW_SPECS, dense):
            if s.name == "t_e":
                for j in range(world):
                    blk = ref[s.name][:, 4 * j:4 * j + 4].clone()
                    oracle.rowwise_adagrad_dense(blk, torch.zeros(s.num_embeddings), gr[:, 4 * j:4 * j + 4] / world, lr=LR)
                    ref[s.name][:, 4 * j:4 * j + 4] = blk
            else:
                oracle.rowwise_adagrad_dense(ref[s.name], torch.zeros(s.num_embeddings), gr / world, lr=LR)

        # state dict: the column-wise table is a ShardedTensor with column shards; gathered as utils/model_training.py:161-182 does
        sd = model.state_dict()
        for s in CW_SPECS:
            t = sd[f"ebc.embedding_bags.{s.name}.weight"]
            assert isinstance(t, ShardedTensor) and tuple(t.size()) == (s.num_embeddings, s.embedding_dim)
            full_t = torch.zeros(t.size()) if rank == 0 else None
            t.gather(0, full_t)
            if rank == 0:
                torch.testing.assert_close(full_t, ref[s.name], rtol=1e-5, atol=1e-6, msg=lambda m: f"{s.name}: {m}")
        md = sd["ebc.embedding_bags.t_e.weight"].metadata().shards_metadata
        assert sorted((m.shard_offsets, m.shard_sizes) for m in md) == [([0, 0], [20, 4]), ([0, 4], [20, 4])]

        # resume from the module's OWN sharded state dict (no gather): weights survive a round trip through zeros
        keep = sharded.tw_ebc.embedding_bags["t_e"].weight.detach().clone()
        own = sharded.state_dict()
        own = {k: v for k, v in own.items()}
        saved = {k: [sh.tensor.clone() for sh in v.local_shards()] for k, v in own.items()}
        with torch.no_grad():
            sharded.tw_ebc.embedding_bags["t_e"].weight.zero_()
        for k, v in own.items():                      # the shards alias the live weights: put the saved values back
            for sh, t0 in zip(v.local_shards(), saved[k]):
                sh.tensor.copy_(t0)
        sharded.load_state_dict(own)
        torch.testing.assert_close(sharded.tw_ebc.embedding_bags["t_e"].weight.detach(), keep)

        # the optimizer-state checkpoint refuses column-wise tables instead of writing something ambiguous
        sharded.include_optimizer_state(True)
        try:
            sharded.state_dict()
            raise AssertionError("include_optimizer_state with a column-wise table must raise")
        except NotImplementedError:
            pass
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def test_column_wise_sharding_world2_gloo():
    """A column-wise table (two features, mean pooling) next to a table-wise and a row-wise one: forward equals the unsharded
    lookup, the fused update equals row-wise Adagrad per column shard on the global batch's gradient / W, the state dict holds
    column shards that ShardedTensor.gather reassembles, and the module reloads both full tensors and its own shards."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_worker_column_wise, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ data-parallel tables (the fourth sharding type the planner honours)
DP_SPECS = [TableSpec("t_a", 40, 8, ["a"], "sum"), TableSpec("t_s", 9, 4, ["s1", "s2"], "mean"), TableSpec("t_c", 101, 4, ["c"], "sum")]
DP_KEYS = ["c", "s2", "a", "s1"]


def _dp_batch(rank, step=0):
    from helpers import random_kjt
    rows = {"a": 40, "s1": 9, "s2": 9, "c": 101}
    return random_kjt(DP_KEYS, [rows[k] for k in DP_KEYS], B, 3, seed=300 + 10 * step + rank)


def _worker_data_parallel(rank, world, port, optimizer, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

        g = torch.Generator().manual_seed(11)
        full = {s.name: torch.randn(s.num_embeddings, s.embedding_dim, generator=g) for s in DP_SPECS}
        cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=s.embedding_dim, num_embeddings=s.num_embeddings,
                                      feature_names=list(s.feature_names),
                                      pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in DP_SPECS]
        ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
        opt_cls = {"adagrad": tt.RowWiseAdagrad, "adam": tt.RowWiseAdam, "sgd": torch.optim.SGD}[optimizer]
        apply_optimizer_in_backward(opt_cls, ebc.parameters(), {"lr": LR})
        holder = nn.ModuleDict({"ebc": ebc})
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                           constraints={"t_s": ParameterConstraints(sharding_types=["data_parallel"]),
                                                        "t_c": ParameterConstraints(sharding_types=["row_wise"])}
                                           ).collective_plan(holder, tt.get_default_sharders(), dist.GroupMember.WORLD)
        assert plan.plan["ebc"]["t_s"].sharding_type == "data_parallel" and plan.plan["ebc"]["t_s"].ranks == [0, 1]
        model = tt.DistributedModelParallel(module=holder, device=torch.device("cpu"), plan=plan,
                                            sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize))
        sharded = model.module["ebc"]
        assert sharded.shard_info()["t_s"] == ("data_parallel", 0, 9) and sharded.dp_ebc is not None
        rep = sharded.dp_ebc.embedding_bags["t_s"].weight
        assert not hasattr(rep, "_optimizer_classes")                       # the replica's backward yields the dense gradient
        # construction broadcast rank 0's replica: identical start on every rank (then the known weights are loaded)
        lst = [rep.detach().clone()]
        dist.broadcast_object_list(lst, src=0)
        assert torch.equal(lst[0], rep.detach())
        sharded.load_state_dict({f"embedding_bags.{k}.weight": v for k, v in full.items()})
        torch.testing.assert_close(rep.detach(), full["t_s"])

        # reference: the unsharded update on the GLOBAL batch's gradient / W (dense forms; untouched rows have zero gradient)
        ref = {k: v.clone() for k, v in full.items()}
        st_sum = torch.zeros(9)
        st_m, st_v = torch.zeros(9, 4), torch.zeros(9)
        n_steps = 3
        for step in range(n_steps):
            values, lengths = _dp_batch(rank, step)
            kjt = tt.KeyedJaggedTensor.from_lengths_sync(DP_KEYS, values, lengths)
            kt = sharded(kjt)
            want = oracle.ebc_forward(DP_SPECS, [ref[s.name] for s in DP_SPECS], DP_KEYS, values, lengths)
            assert kt.keys() == ["a", "s1", "s2", "c"] and kt.length_per_key() == [8, 4, 4, 4]
            torch.testing.assert_close(kt.values(), want, rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} forward: {m}")
            gout = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(900 + 10 * step + rank))
            (kt.values() * gout).sum().backward()
            model.sync_dense_grads()                 # what TrainPipelineSparseDist / CudaGraphTrainStep call after the backward
            assert rep.grad is None
            dense = [torch.zeros_like(ref[s.name]) for s in DP_SPECS]
            for r in range(world):
                v_r, l_r = _dp_batch(r, step)
                g_r = torch.randn(B, want.shape[1], generator=torch.Generator().manual_seed(900 + 10 * step + r))
                for acc, gr in zip(dense, oracle.ebc_dense_grads(DP_SPECS, DP_KEYS, v_r, l_r, g_r)):
                    acc += gr
            gs = dense[1] / world
            if optimizer == "adagrad":
                oracle.rowwise_adagrad_dense(ref["t_s"], st_sum, gs, lr=LR)
            elif optimizer == "sgd":
                ref["t_s"] -= LR * gs
            else:                                    # row-wise Adam, touched rows only (oracle's sparse form on the rows with gradient)
                hit = (gs != 0).any(dim=1).nonzero().flatten()
                oracle.rowwise_adam_sparse(ref["t_s"], st_m, st_v, hit, gs[hit], step=step + 1, lr=LR)
            # the sharded tables next to it keep their fused path (the stand-in applies row-wise Adagrad whatever the tag)
            oracle.rowwise_adagrad_dense(ref["t_a"], sharded_state("t_a", ref), dense[0] / world, lr=LR)
            oracle.rowwise_adagrad_dense(ref["t_c"], sharded_state("t_c", ref), dense[2] / world, lr=LR)
            torch.testing.assert_close(rep.detach(), ref["t_s"], rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} replica: {m}")
        # replicas stayed identical
        lst = [rep.detach().clone()]
        dist.broadcast_object_list(lst, src=0)
        assert torch.equal(lst[0], rep.detach())

        # state dict: the replicated table is a PLAIN tensor (utils/model_training.py:178-180 keeps rank 0's copy), the others ShardedTensors
        sd = model.state_dict()
        t = sd["ebc.embedding_bags.t_s.weight"]
        assert isinstance(t, torch.Tensor) and not isinstance(t, ShardedTensor)
        torch.testing.assert_close(t, ref["t_s"], rtol=1e-5, atol=1e-6)
        for name in ("t_a", "t_c"):
            t = sd[f"ebc.embedding_bags.{name}.weight"]
            assert isinstance(t, ShardedTensor)
            full_t = torch.zeros(t.size()) if rank == 0 else None
            t.gather(0, full_t)
            if rank == 0:
                torch.testing.assert_close(full_t, ref[name], rtol=1e-5, atol=1e-6, msg=lambda m: f"{name}: {m}")
        # optimizer state of the replica travels behind include_optimizer_state and reloads
        sharded.include_optimizer_state(True)
        sd2 = sharded.state_dict()
        if optimizer == "adagrad":
            torch.testing.assert_close(sd2["embedding_bags.t_s.sum"], st_sum, rtol=1e-5, atol=1e-7)
        elif optimizer == "adam":
            torch.testing.assert_close(sd2["embedding_bags.t_s.exp_avg"], st_m, rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(sd2["embedding_bags.t_s.exp_avg_sq"], st_v, rtol=1e-5, atol=1e-7)
            assert float(sd2["fused_optimizer_step"]) == n_steps
        # (the replica's entries only: the stand-in of the sharded tables keeps no reloadable state; their round trip is
        # test_sharded_checkpoint_with_optimizer_state_world2_gloo's subject)
        keep = {k: v.clone() for k, v in sd2.items() if ".t_s." in k or k == "fused_optimizer_step"}
        sharded._dp_state.clear()
        sharded._dp_step = 0
        with torch.no_grad():
            rep.zero_()
        sharded.load_state_dict(keep, strict=False)
        torch.testing.assert_close(rep.detach(), ref["t_s"], rtol=1e-5, atol=1e-6)
        if optimizer == "adagrad":
            torch.testing.assert_close(sharded._dp_state["t_s"]["sum"], st_sum, rtol=1e-5, atol=1e-7)
        if optimizer == "adam":
            assert sharded._dp_step == n_steps
        # an evaluation forward (no backward, no sync) leaves everything as it is
        with torch.no_grad():
            values, lengths = _dp_batch(rank, 7)
            kt = sharded(tt.KeyedJaggedTensor.from_lengths_sync(DP_KEYS, values, lengths))
        torch.testing.assert_close(kt["s2"], oracle.ebc_forward(DP_SPECS, [ref[s.name] for s in DP_SPECS], DP_KEYS, values, lengths)[:, 12:16],
                                   rtol=1e-5, atol=1e-6)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


_SHARDED_STATE = {}


def sharded_state(name, ref):
    """Row-wise Adagrad accumulator of a sharded (non-replicated) table of the reference run."""
    return _SHARDED_STATE.setdefault(name, torch.zeros(ref[name].shape[0]))


def _worker_data_parallel_all(rank, world, port, errq):
    """One pair of processes takes the three optimizers in turn (a process start costs more than the test itself)."""
    for i, optimizer in enumerate(("adagrad", "adam", "sgd")):
        _SHARDED_STATE.clear()
        _worker_data_parallel(rank, world, port + i, optimizer, errq)


def test_data_parallel_tables_world2_gloo():
    """A data_parallel table (two features, mean pooling) next to a table-wise and a row-wise one: the replica is looked up on
    the rank's own batch, `sync_dense_grads` averages its dense gradient over the ranks and applies the tagged optimizer
    (row-wise Adagrad / row-wise Adam / SGD) identically on every rank -- equal to the unsharded update on the global batch's
    gradient / W over three steps; the state dict holds it as a plain tensor; its optimizer state reloads."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_data_parallel_all, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ corpus embedding against sharded tables + both multi-rank retrieval layouts
def _worker_sharded_retrieval(rank, world, port, sharding, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from two_tower_recommender_model_b200 import _native as N, retrieval
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
        from two_tower_recommender_model_b200.modules import mlp
        # device work -> oracle / plain torch (tests only)
        mlp.linear_act = lambda x, w, b, relu: torch.relu(torch.nn.functional.linear(x, w, b)) if relu else torch.nn.functional.linear(x, w, b)
        N.require_cuda = lambda t, name: None

        def topk(q, items, k, item_index_base=0, precision="fp32", items_bf16=None):
            s, i = oracle.exact_topk(q, items, k)
            return s, i + item_index_base
        retrieval.score_topk = topk

        cat, emb, dim, layers = ["user_id", "product_id"], [13, 37], 8, [8, 4]      # 37 items: not a multiple of the world size
        specs = [TableSpec(f"t_{c}", emb[i], dim, [c]) for i, c in enumerate(cat)]
        orc = oracle.OracleTwoTower(specs, layers, loss="bce", seed=21)
        ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                                for i, c in enumerate(cat)], device=torch.device("meta"))
        tower = tt.TwoTower(ebc, layers, device=torch.device("cpu"))
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                           constraints={f"t_{c}": ParameterConstraints(sharding_types=[sharding]) for c in cat}
                                           ).collective_plan(tower, tt.get_default_sharders(), dist.GroupMember.WORLD)
        model = tt.DistributedModelParallel(module=tower, device=torch.device("cpu"), plan=plan,
                                            sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize))
        model.load_state_dict(orc.torchrec_state_dict())
        tw = model.module
        tw.eval()
        with torch.no_grad():
            iv = torch.arange(emb[1])
            il = torch.cat([torch.zeros(emb[1], dtype=torch.int32), torch.ones(emb[1], dtype=torch.int32)])
            _, items_ref = orc.forward(cat, iv, il)
            uv = torch.arange(emb[0])
            ul = torch.cat([torch.ones(emb[0], dtype=torch.int32), torch.zeros(emb[0], dtype=torch.int32)])
            users_ref, _ = orc.forward(cat, uv, ul)
        per = -(-emb[1] // world)
        for chunk in (5, 7, 19, 1 << 18):                 # chunks that do and do not divide a rank's share
            local, first = tt.embed_corpus_sharded(tw, cat, "product_id", emb[1], torch.device("cpu"), chunk=chunk)
            lo, hi = rank * per, min((rank + 1) * per, emb[1])
            assert first == lo and local.shape == (hi - lo, layers[-1]), (chunk, local.shape)
            torch.testing.assert_close(local, items_ref[lo:hi], rtol=1e-5, atol=1e-6, msg=lambda m: f"chunk {chunk}: {m}")
        # every rank asks for ITS block of users (6 + 6 of the 13, the last one left out): same count on both ranks
        q = users_ref[rank * 6:(rank + 1) * 6]
        ws, wi = oracle.exact_topk(q, items_ref, 10)
        gathered = tt.BruteForceIndex.from_sharded(local, precision="fp32")           # corpus all-gathered, queries sharded
        s1, i1 = gathered.search(q, 10)
        kept = tt.CorpusShardedIndex(local, first_id=first)                           # corpus stays sharded, lists merged
        s2, i2 = kept.search(q, 10)
        for s_, i_ in ((s1, i1), (s2, i2)):
            torch.testing.assert_close(s_, ws, rtol=1e-5, atol=1e-6)
            assert float((i_ == wi).float().mean()) >= 0.98                           # near-ties may swap under another summation order
        assert torch.equal(i1, i2)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def _worker_sharded_retrieval_all(rank, world, port, errq):
    for i, sharding in enumerate(("table_wise", "row_wise")):          # one pair of processes, both shardings in turn
        _worker_sharded_retrieval(rank, world, port + i, sharding, errq)


def test_sharded_corpus_embedding_and_retrieval_world2_gloo():
    """embed_corpus_sharded against table-wise / row-wise sharded tables (a corpus size the world size does not divide, chunk
    sizes that do not divide a rank's share), then both multi-rank retrieval layouts -- BruteForceIndex.from_sharded and
    CorpusShardedIndex -- against the oracle's towers and exact top-k."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 30010 + os.getpid() % 50
    procs = [ctx.Process(target=_worker_sharded_retrieval_all, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ sharded checkpoint with the fused optimizer state
def _worker_sharded_resume(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import two_tower_recommender_model_b200 as tt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        from two_tower_recommender_model_b200 import _native as N
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
        from two_tower_recommender_model_b200.modules import embedding_modules
        from test_reference_boundary import _OracleLookup as _FusedOracleLookup
        # the REAL local EmbeddingBagCollection (its state_dict / optimizer-state bookkeeping is what is under test);
        # only its device entry point is replaced by the oracle
        embedding_modules.EbcLookup = _FusedOracleLookup
        N.require_cuda = lambda t, name: None
        keys, rows, dim = ["a", "c"], [40, 33], 8

        def build(seed):
            torch.manual_seed(seed)
            cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=dim, num_embeddings=rows[j], feature_names=[k]) for j, k in enumerate(keys)]
            ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
            apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": LR})
            holder = nn.ModuleDict({"ebc": ebc})
            plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                               constraints={"t_a": ParameterConstraints(sharding_types=["table_wise"]),
                                                            "t_c": ParameterConstraints(sharding_types=["row_wise"])}
                                               ).collective_plan(holder, tt.get_default_sharders(), dist.GroupMember.WORLD)
            model = tt.DistributedModelParallel(module=holder, device=torch.device("cpu"), plan=plan,
                                                sharding_kwargs=dict(bucketize_fn=oracle_bucketize))
            return model.module["ebc"]

        def step(m, s):
            from helpers import random_kjt
            v, l = random_kjt(keys, rows, B, 3, seed=500 + 10 * s + rank, dup_pool=30)
            out = m(tt.KeyedJaggedTensor.from_lengths_sync(keys, v, l)).values()
            g = torch.randn(out.shape, generator=torch.Generator().manual_seed(600 + 10 * s + rank))
            (out * g).sum().backward()

        def local_state(m):
            out = {}
            for local in (m.tw_ebc, m.rw_ebc):
                if local is not None:
                    for name, bag in local.embedding_bags.items():
                        out[name + ".weight"] = bag.weight.detach().clone()
                        out[name + ".sum"] = local.fused_optimizer_state()[name]["sum"].clone()
            return out

        a = build(0)
        for s in range(2):
            step(a, s)
        plain = a.state_dict()
        assert set(plain) == {"embedding_bags.t_a.weight", "embedding_bags.t_c.weight"}
        full = a.include_optimizer_state(True).state_dict()
        assert set(full) == set(plain) | {"embedding_bags.t_a.sum", "embedding_bags.t_c.sum", "fused_optimizer_step"}
        assert isinstance(full["embedding_bags.t_c.sum"], ShardedTensor) and tuple(full["embedding_bags.t_c.sum"].size()) == (33,)
        # the shards alias the live tensors: a checkpoint writer would serialise them here; keep copies instead
        frozen = {k: ([sh.tensor.clone() for sh in v.local_shards()] if isinstance(v, ShardedTensor) else v.clone()) for k, v in full.items()}
        b = build(1)                           # resumes from the sharded checkpoint, no gather
        for k, v in full.items():              # hand b a checkpoint whose shards hold the frozen values
            if isinstance(v, ShardedTensor):
                for sh, t0 in zip(v.local_shards(), frozen[k]):
                    assert torch.equal(sh.tensor, t0)
        b.load_state_dict(full)
        # gathered (full-tensor) form of the same checkpoint, as utils/model_training.py:161-182 writes it, weights + state
        gathered = {}
        for k, v in full.items():
            if isinstance(v, ShardedTensor):
                out = torch.zeros(v.size()) if rank == 0 else None
                v.gather(0, out)
                lst = [out]
                dist.broadcast_object_list(lst, src=0)
                gathered[k] = lst[0]
            else:
                gathered[k] = v
        c = build(2)
        c.load_state_dict(gathered)
        d = build(3)                           # weights only: the accumulators restart
        d.load_state_dict({k: v for k, v in gathered.items() if k.endswith(".weight")})
        for m in (a, b, c, d):
            step(m, 2)
        sa, sb, sc, sd_ = local_state(a), local_state(b), local_state(c), local_state(d)
        assert sa, "every rank holds a row-wise shard"
        for k in sa:
            assert torch.equal(sa[k], sb[k]), f"own-shard resume differs at {k}"
            assert torch.equal(sa[k], sc[k]), f"gathered resume differs at {k}"
        differs = torch.tensor([1.0 if any(not torch.equal(sa[k], sd_[k]) for k in sa if k.endswith(".weight")) else 0.0])
        dist.all_reduce(differs, op=dist.ReduceOp.MAX)          # a rank whose shard saw no repeated row cannot tell
        assert float(differs) == 1.0, "the weights-only resume took the same step: the accumulators did not matter?"
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def test_sharded_checkpoint_with_optimizer_state_world2_gloo():
    """Two ranks, one table-wise and one row-wise table: after include_optimizer_state(True) the sharded state dict carries
    the fused row-wise accumulators as ShardedTensors; a module resumed from its own shards, and one resumed from the gathered
    full tensors, take the same third step as the module that kept training; a weights-only resume does not."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 30140 + os.getpid() % 50
    procs = [ctx.Process(target=_worker_sharded_resume, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


# ------------------------------------------------------------------ random MIXED plans: all four sharding types in one collection
def _worker_mixed_plans(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import random
        import two_tower_recommender_model_b200 as tt
        from helpers import random_kjt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

        for seed in range(int(os.environ.get("TT_MIXED_PLAN_SEEDS", "8"))):
            rnd = random.Random(1000 + seed)                      # the same draw on every rank
            n_tables = rnd.randint(2, 5)
            kinds = [rnd.choice(["table_wise", "row_wise", "column_wise", "data_parallel"]) for _ in range(n_tables)]
            specs, feat_rows = [], {}
            for t, kind in enumerate(kinds):
                dim = rnd.choice([4, 8, 12])
                rows = rnd.choice([1, 2, 3]) if rnd.random() < 0.25 else rnd.randint(5, 60)      # 1 row over 2 ranks: a zero-row shard
                feats = [f"f{t}_{j}" for j in range(rnd.randint(1, 2))]
                specs.append(TableSpec(f"t{t}", rows, dim, feats, rnd.choice(["sum", "mean"])))
                for f in feats:
                    feat_rows[f] = rows
            keys = [f for s in specs for f in s.feature_names]
            rnd.shuffle(keys)                                     # KJT key order unrelated to table order
            Bm = rnd.randint(3, 7)
            g = torch.Generator().manual_seed(seed)
            full = {s.name: torch.randn(s.num_embeddings, s.embedding_dim, generator=g) for s in specs}
            cfgs = [tt.EmbeddingBagConfig(name=s.name, embedding_dim=s.embedding_dim, num_embeddings=s.num_embeddings,
                                          feature_names=list(s.feature_names),
                                          pooling=tt.PoolingType.MEAN if s.pooling == "mean" else tt.PoolingType.SUM) for s in specs]
            ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
            apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": LR})
            holder = nn.ModuleDict({"ebc": ebc})
            plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world, compute_device="cpu"),
                                               constraints={s.name: ParameterConstraints(sharding_types=[k]) for s, k in zip(specs, kinds)}
                                               ).collective_plan(holder, tt.get_default_sharders(), dist.GroupMember.WORLD)
            assert [plan.plan["ebc"][s.name].sharding_type for s in specs] == kinds
            model = tt.DistributedModelParallel(module=holder, device=torch.device("cpu"), plan=plan,
                                                sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize))
            sharded = model.module["ebc"]
            sharded.load_state_dict({f"embedding_bags.{k}.weight": v for k, v in full.items()})
            tag = f"seed {seed} kinds {kinds} keys {keys}"

            def batch(r):
                return random_kjt(keys, [feat_rows[k] for k in keys], Bm, 3, seed=7000 + 10 * seed + r)

            values, lengths = batch(rank)
            kt = sharded(tt.KeyedJaggedTensor.from_lengths_sync(keys, values, lengths))
            want = oracle.ebc_forward(specs, [full[s.name] for s in specs], keys, values, lengths)
            assert kt.keys() == [f for s in specs for f in s.feature_names], tag
            torch.testing.assert_close(kt.values(), want, rtol=1e-5, atol=1e-6, msg=lambda m: f"{tag} forward: {m}")
            gout = torch.randn(Bm, want.shape[1], generator=torch.Generator().manual_seed(8000 + 10 * seed + rank))
            (kt.values() * gout).sum().backward()
            model.sync_dense_grads()
            # reference: row-wise Adagrad on the global batch's gradient / W; column-wise tables per column shard
            dense = [torch.zeros_like(full[s.name]) for s in specs]
            for r in range(world):
                v_r, l_r = batch(r)
                g_r = torch.randn(Bm, want.shape[1], generator=torch.Generator().manual_seed(8000 + 10 * seed + r))
                for acc, gr in zip(dense, oracle.ebc_dense_grads(specs, keys, v_r, l_r, g_r)):
                    acc += gr
            ref = {k: v.clone() for k, v in full.items()}
            for s, kind, gr in zip(specs, kinds, dense):
                if kind == "column_wise":
                    shards = len(plan.plan["ebc"][s.name].ranks)
                    dw = s.embedding_dim // shards
                    for j in range(shards):
                        blk = ref[s.name][:, j * dw:(j + 1) * dw].clone()
                        oracle.rowwise_adagrad_dense(blk, torch.zeros(s.num_embeddings), gr[:, j * dw:(j + 1) * dw] / world, lr=LR)
                        ref[s.name][:, j * dw:(j + 1) * dw] = blk
                else:
                    oracle.rowwise_adagrad_dense(ref[s.name], torch.zeros(s.num_embeddings), gr / world, lr=LR)
            sd = model.state_dict()
            for s, kind in zip(specs, kinds):
                t = sd[f"ebc.embedding_bags.{s.name}.weight"]
                if kind == "data_parallel":
                    assert not isinstance(t, ShardedTensor), tag
                    got = t
                else:
                    assert isinstance(t, ShardedTensor) and tuple(t.size()) == (s.num_embeddings, s.embedding_dim), tag
                    got = torch.zeros(t.size()) if rank == 0 else None
                    t.gather(0, got)
                if rank == 0 or kind == "data_parallel":
                    torch.testing.assert_close(got, ref[s.name], rtol=1e-5, atol=1e-6, msg=lambda m: f"{tag} table {s.name} ({kind}): {m}")
            # the pipeline's prefetch hook (input dist issued one batch ahead) on the updated tables, under no_grad (an evaluation pass)
            v2, l2 = random_kjt(keys, [feat_rows[k] for k in keys], Bm, 3, seed=9000 + 10 * seed + rank)
            kjt2 = tt.KeyedJaggedTensor.from_lengths_sync(keys, v2, l2)
            model.start_sparse_data_dist(tt.Batch(torch.zeros(1), kjt2, torch.zeros(Bm, dtype=torch.int32)), None)
            with torch.no_grad():
                kt2 = sharded(kjt2)
            lst = [ref]
            dist.broadcast_object_list(lst, src=0)            # rank 0's reference (identical on every rank by construction)
            want2 = oracle.ebc_forward(specs, [lst[0][s.name] for s in specs], keys, v2, l2)
            torch.testing.assert_close(kt2.values(), want2, rtol=1e-5, atol=1e-6, msg=lambda m: f"{tag} prefetched forward: {m}")
            # resume: a second module built from the same plan loads the first one's OWN state dict (row shards, column shards,
            # replicated tensors -- no gather) and gives the same lookup
            ebc_b = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
            apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc_b.parameters(), {"lr": LR})
            model_b = tt.DistributedModelParallel(module=nn.ModuleDict({"ebc": ebc_b}), device=torch.device("cpu"), plan=plan,
                                                  sharding_kwargs=dict(local_ebc_factory=OracleLocalEbc, bucketize_fn=oracle_bucketize))
            own = {k: (v.clone() if not isinstance(v, ShardedTensor) else v) for k, v in sharded.state_dict().items()}
            res = model_b.module["ebc"].load_state_dict(own)
            assert not res.missing_keys, tag
            with torch.no_grad():
                kt3 = model_b.module["ebc"](tt.KeyedJaggedTensor.from_lengths_sync(keys, v2, l2))
            torch.testing.assert_close(kt3.values(), kt2.values(), rtol=0, atol=0, msg=lambda m: f"{tag} reloaded module: {m}")
            dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def test_random_mixed_sharding_plans_world2_gloo():
    """Eight random collections (2-5 tables, 1-2 features each, sum / mean pooling, shuffled KJT key order) whose tables draw their
    sharding type from all four -- table-wise, row-wise, column-wise, data-parallel -- IN ONE collection: forward equals the
    unsharded lookup of the rank's batch, and after one fused step every table equals row-wise Adagrad on the global batch's
    gradient / W (per column shard for column-wise tables), gathered as utils/model_training.py:161-182 does."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29400 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_mixed_plans, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)
