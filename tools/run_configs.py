"""Runs the BASELINE.json configurations that are NOT the bench headline, at full size on one B200, a
few steps each, and prints one JSON line per configuration (timing with CUDA events, eager loop).

    python tools/run_configs.py [3] [4] [5]

  3: multi-hot user history (L=20, mean) + product/aisle/department tables, 100M rows x 128 (51 GB each)
  4: batch 262144, bf16 towers 128->1024->512->256, fused row-wise Adam, in-batch softmax d=256
  5: retrieval: item-tower corpus of 10M items, top-100 for 131072 queries (1/8 of the 1M: one GPU's share)

`config3_sharded` / `config4_sharded` (configs[2] row-wise sharded, configs[3] under the planner's sharding, over the ranks of a
multi-GPU run) are called by `bench.py --gpus N`.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward  # noqa: E402

import two_tower_recommender_model_b200 as tt  # noqa: E402
from two_tower_recommender_model_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")


def timed_steps(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def _peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


def _instrumented(step, n=2):
    """Per-call device times (CUDA events around every library call) over n steps: name -> (ms per call, calls per step)."""
    N.enable_timing(True)
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    per = {k: (v["ms"], v["calls"] / n) for k, v in N.timing_summary().items()}
    N.enable_timing(False)
    return per


def config3(steps=5, warmup=2):
    """configs[2] on one GPU: multi-hot user history (pooling factor 20, mean) + product / aisle / department features,
    two 100M-row x 128 fp32 tables (51 GB each) + two tiny ones, B = 65536, towers 128 -> [128, 64], in-batch softmax."""
    B, D, L = 65536, 128, 20
    hbm, tf, src = _peaks()
    rows = {"hist": 100_000_000, "product": 100_000_000, "aisle": 134, "department": 21}
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=D, num_embeddings=r, feature_names=[k],
                                  pooling=tt.PoolingType.MEAN if k == "hist" else tt.PoolingType.SUM) for k, r in rows.items()]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=dev)
    tower = tt.TwoTower(ebc, [128, 64], device=dev, query_features=["hist"], candidate_features=["product", "aisle", "department"],
                        precision="bf16")
    task = tt.TwoTowerTrainTask(tower, loss="in_batch_softmax", precision="bf16")
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.01})
    opt = tt.KeyedOptimizerWrapper(dict(task.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))
    keys = list(rows)
    g = torch.Generator(device=dev).manual_seed(3)
    batches, uniq = [], 0
    for _ in range(3):
        lens = torch.cat([torch.full((B,), L, dtype=torch.int32, device=dev), torch.ones(3 * B, dtype=torch.int32, device=dev)])
        cols = [torch.randint(0, rows["hist"], (B * L,), device=dev, generator=g), torch.randint(0, rows["product"], (B,), device=dev, generator=g),
                torch.randint(0, 134, (B,), device=dev, generator=g), torch.randint(0, 21, (B,), device=dev, generator=g)]
        uniq = sum(int(torch.unique(c).numel()) for c in cols)
        batches.append(tt.Batch(torch.zeros(1, device=dev), tt.KeyedJaggedTensor.from_lengths_sync(keys, torch.cat(cols), lens),
                                torch.zeros(B, dtype=torch.int32, device=dev)))
    i = [0]

    def step():
        opt.zero_grad()
        loss, _ = task(batches[i[0] % 3]); i[0] += 1
        loss.backward()
        opt.step()
    ms_eager = timed_steps(step, steps=steps, warmup=warmup)
    per = _instrumented(step)
    # the same step replayed as ONE CUDA graph (multi-hot KJT in a fixed-capacity buffer, offsets scanned on the device)
    gstep = tt.CudaGraphTrainStep(task, opt, keys, list(rows.values()), B, dev, warmup_steps=2, kjt_capacity=B * L + 3 * B)
    parts = [(b.sparse_features.values(), b.sparse_features.lengths(), b.labels) for b in batches]
    ms = timed_steps(lambda: gstep.step_kjt(*parts[i[0] % 3]), steps=steps, warmup=4)
    nnz = B * L + 3 * B
    fwd_bytes = nnz * (8 + 4 * D) + 4 * (4 * B) + 4 * B * D * 4           # ids + rows, offsets, pooled output
    bwd_bytes = 4 * B * D * 4 + 8 * nnz + uniq * (8 * D + 8)               # 4BD + 8BL + U(8D + 8) per feature
    fwd_ms, bwd_ms = per["tt_ebc_forward"][0], per["tt_ebc_backward_fused"][0]
    sm_ms = sum(v[0] * v[1] for k, v in per.items() if "softmax" in k)
    out = {"config": 3, "what": "configs[2] on 1 GPU: L=20 mean history + product/aisle/department, 2 x 100M x 128 fp32 tables (102 GB), "
                                "B=65536, towers 128-[128,64], in-batch softmax, fused row-wise Adagrad; whole step replayed as one CUDA graph",
           "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3), "ms_per_step_eager": round(ms_eager, 3), "cuda_graph": gstep.captured,
           "calls_ms": {k: round(v[0] * v[1], 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1][0] * kv[1][1])},
           "ebc_lookup": {"ms": round(fwd_ms, 4), "bytes": fwd_bytes, "gbs": round(fwd_bytes / fwd_ms / 1e6, 1), "peak_gbs": hbm,
                          "frac": round(fwd_bytes / fwd_ms / 1e6 / hbm, 4)},
           "ebc_update": {"ms": round(bwd_ms, 4), "bytes": bwd_bytes, "unique_rows": uniq, "gbs": round(bwd_bytes / bwd_ms / 1e6, 1),
                          "peak_gbs": hbm, "frac": round(bwd_bytes / bwd_ms / 1e6 / hbm, 4)},
           "softmax": {"ms": round(sm_ms, 4), "tflops_credited": round(6.0 * B * B * 64 / sm_ms / 1e9, 1), "peak": tf,
                       "frac": round(6.0 * B * B * 64 / sm_ms / 1e9 / tf, 4)},
           "peak_source": src, "hbm_gb_allocated": round(torch.cuda.memory_allocated() / 1e9, 1)}
    del task, tower, ebc, opt, batches, gstep, parts
    torch.cuda.empty_cache()
    return out


def config3_sharded(device, rank, world, steps=5, warmup=3, batch=65536, big_rows=100_000_000, D=128, L=20, layers=(128, 64),
                    precision="bf16", peer_exchange=True):
    """configs[2] AS STATED on `world` GPUs (called by every rank of a `bench.py --gpus N` run): the four tables ROW-WISE sharded
    (100M x 128 fp32 hist / product: 51 GB each, 51 / world per rank; aisle / department alongside), GLOBAL batch 65536, mean-
    pooled history of L = 20, in-batch softmax with per-rank negatives, fused row-wise Adagrad, Adam on the towers.  The
    per-rank KJT sits in a fixed-capacity buffer, so the input dist is the sync-free all-gather + tt_kjt_gathered_range route,
    the partial sums of the shards are ADDED into the sample's rank over NVLink peer memory, and the whole step replays as ONE
    CUDA graph (the combination tests/test_gpu_multi.py::test_two_rank_multi_hot_sync_free_input_dist_and_graph checks
    against the oracle at small size).  Timed with CUDA events, max over ranks; value = global batch / step time.
    (`peer_exchange=False`, `precision="fp32"`, small sizes: the CPU dry run of tests/dryrun_bench_world2.py.)"""
    import torch.distributed as dist
    from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
    Br = batch // world
    rows = {"hist": big_rows, "product": big_rows, "aisle": 134, "department": 21}
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=D, num_embeddings=r, feature_names=[k],
                                  pooling=tt.PoolingType.MEAN if k == "hist" else tt.PoolingType.SUM) for k, r in rows.items()]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
    tower = tt.TwoTower(ebc, list(layers), device=device, query_features=["hist"], candidate_features=["product", "aisle", "department"],
                        precision=precision)
    task = tt.TwoTowerTrainTask(tower, loss="in_batch_softmax", precision=precision)
    apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": 0.01})
    cons = {f"t_{k}": ParameterConstraints(sharding_types=["row_wise"]) for k in rows}
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world), constraints=cons
                                       ).collective_plan(task, tt.get_default_sharders(), dist.GroupMember.WORLD)
    model = tt.DistributedModelParallel(module=task, device=device, plan=plan, sharding_kwargs={"peer_exchange": True} if peer_exchange else None)
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))
    model.train()
    keys = list(rows)
    g = torch.Generator(device=device).manual_seed(300 + rank)
    parts = []
    for _ in range(3):
        lens = torch.cat([torch.full((Br,), L, dtype=torch.int32, device=device), torch.ones(3 * Br, dtype=torch.int32, device=device)])
        cols = [torch.randint(0, rows["hist"], (Br * L,), device=device, generator=g), torch.randint(0, rows["product"], (Br,), device=device, generator=g),
                torch.randint(0, 134, (Br,), device=device, generator=g), torch.randint(0, 21, (Br,), device=device, generator=g)]
        parts.append((torch.cat(cols), lens, torch.zeros(Br, dtype=torch.int32, device=device)))
    gstep = tt.CudaGraphTrainStep(model, opt, keys, list(rows.values()), Br, device, warmup_steps=2, kjt_capacity=Br * L + 3 * Br)
    loss = None
    for i in range(2 + 1 + max(warmup, 1)):            # eager warm-up steps, the capture, replays
        loss = gstep.step_kjt(*parts[i % 3])[0]
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = gstep.step_kjt(*parts[i % 3])[0]
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    hbm, _tf, src = _peaks()
    nnz = Br * L + 3 * Br                                 # ids a rank contributes; it looks up ~ world * nnz / world of the global batch
    fwd_bytes = nnz * (8 + 4 * D) + 4 * (4 * Br) + 4 * Br * D * 4
    out = {"config": 3, "what": "configs[2] on %d GPUs: 4 tables row-wise sharded (2 x 100M x 128 fp32 = 102 GB over the ranks), GLOBAL batch %d "
                                "(per-rank %d), L=20 mean history, in-batch softmax (per-rank negatives), fused row-wise Adagrad, Adam; sync-free "
                                "multi-hot input dist, peer-memory exchange, whole step as one CUDA graph" % (world, batch, Br),
           "value": round(batch / ms * 1e3, 1), "unit": "samples/s", "ms_per_step": round(ms, 4), "global_batch": batch, "per_rank_batch": Br,
           "sharding": sorted({ps.sharding_type for tables in plan.plan.values() for ps in tables.values()}), "cuda_graph": gstep.captured,
           "last_loss": float(loss), "ids_per_rank_per_step": nnz, "lookup_bytes_per_rank": fwd_bytes, "hbm_peak_gbs": hbm, "peak_source": src,
           "hbm_gb_allocated_per_rank": round(torch.cuda.memory_allocated() / 1e9, 1) if torch.cuda.is_available() else None}
    del model, task, tower, ebc, opt, gstep, parts
    torch.cuda.empty_cache()
    return out


def config4(steps=3, warmup=1):
    """configs[3] on one GPU: batch 262144, bf16 towers 128 -> 1024 -> 512 -> 256 (both towers), in-batch softmax over
    d = 256, fused row-wise Adam on two 10M x 128 tables."""
    B, D = 262144, 128
    hbm, tf, src = _peaks()
    rows = [10_000_000, 10_000_000]
    cat = ["user_id", "product_id"]
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=D, num_embeddings=rows[i], feature_names=[c]) for i, c in enumerate(cat)]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=dev)
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, [1024, 512, 256], device=dev, precision="bf16"), loss="in_batch_softmax", precision="bf16")
    apply_optimizer_in_backward(tt.RowWiseAdam, ebc.parameters(), {"lr": 0.01})
    opt = tt.KeyedOptimizerWrapper(dict(task.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))
    rows_dev = torch.tensor(rows, device=dev)
    g = torch.Generator(device=dev).manual_seed(4)
    batches = []
    for _ in range(2):
        ids = torch.stack([torch.randint(1, r, (B,), device=dev, generator=g) for r in rows])
        batches.append(tt.Batch(torch.zeros(1, device=dev), tt.KeyedJaggedTensor.from_id_columns(cat, ids, rows_dev),
                                torch.zeros(B, dtype=torch.int32, device=dev)))
    i = [0]

    def step():
        opt.zero_grad()
        loss, _ = task(batches[i[0] % 2]); i[0] += 1
        loss.backward()
        opt.step()
    ms_eager = timed_steps(step, steps=steps, warmup=warmup)
    per = _instrumented(step)
    gstep = tt.CudaGraphTrainStep(task, opt, cat, rows, B, dev, warmup_steps=1)
    raw = [(b.sparse_features._id_columns[0], b.labels) for b in batches]
    ms = timed_steps(lambda: gstep(*raw[i[0] % 2]), steps=steps, warmup=3)
    sm = {k: v for k, v in per.items() if "softmax" in k}
    sm_ms = sum(v[0] * v[1] for v in sm.values())
    fwd_ms = sum(v[0] * v[1] for k, v in sm.items() if "forward" in k)
    bwd_ms = sum(v[0] * v[1] for k, v in sm.items() if "backward" in k)
    gemm_ms = sum(v[0] * v[1] for k, v in per.items() if "gemm" in k or "towers" in k)
    layer_flops = 2.0 * B * (128 * 1024 + 1024 * 512 + 512 * 256) * 2           # both towers, one GEMM per layer
    tower_flops = 3 * layer_flops                                               # y, dX (the embeddings train), dW per layer
    logit_flops = 6.0 * B * B * 256
    out = {"config": 4, "what": "configs[3] on 1 GPU: B=262144, bf16 towers 128-1024-512-256, in-batch softmax d=256, 2 x 10M x 128 tables, "
                                "fused row-wise Adam; whole step replayed as one CUDA graph",
           "ms_per_step": round(ms, 2), "samples_per_s": round(B / ms * 1e3), "ms_per_step_eager": round(ms_eager, 2), "cuda_graph": gstep.captured,
           "calls_ms": {k: round(v[0] * v[1], 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1][0] * kv[1][1])},
           "roofline": {"kernel": "in-batch softmax d=256: tc_softmax_fwd_kernel<4> + 2 x tc_softmax_bwd_wide_kernel<4> (tcgen05)",
                        "bound": "tensor", "ms": round(sm_ms, 3), "achieved": round(logit_flops / sm_ms / 1e9, 1), "unit": "TFLOP/s",
                        "peak": tf, "frac": round(logit_flops / sm_ms / 1e9 / tf, 4),
                        "executed_tflops": round(10.0 * B * B * 256 / sm_ms / 1e9, 1), "executed_frac": round(10.0 * B * B * 256 / sm_ms / 1e9 / tf, 4),
                        "forward_tflops": round(2.0 * B * B * 256 / max(fwd_ms, 1e-9) / 1e9, 1), "backward_tflops_executed": round(8.0 * B * B * 256 / max(bwd_ms, 1e-9) / 1e9, 1),
                        "flops_credited": "6*B*B*d; executed 10*B*B*d (S recomputed by both backward passes: accX + accY of a one-pass "
                                          "scheme need 2 x 256 TMEM columns, which leaves none for S)"},
           "tower_gemms": {"ms": round(gemm_ms, 3), "tflops": round(tower_flops / max(gemm_ms, 1e-9) / 1e9, 1), "peak": tf,
                           "frac": round(tower_flops / max(gemm_ms, 1e-9) / 1e9 / tf, 4), "flops": "3 GEMMs (y, dX, dW) x 3 layers x 2 towers"},
           "whole_step_tflops_credited": round((logit_flops + tower_flops) / (ms * 1e-3) / 1e12, 1), "peak_source": src}
    del task, ebc, opt, batches, gstep, raw
    torch.cuda.empty_cache()
    return out


def config4_sharded(device, rank, world, steps=5, warmup=3, batch=262144, rows=(10_000_000, 10_000_000), D=128, layers=(1024, 512, 256),
                    precision="bf16", peer_exchange=True):
    """configs[3] AS STATED on `world` GPUs (called by every rank of a `bench.py --gpus N` run): GLOBAL batch 262144 (per-rank
    262144 / world), bf16 towers 128 -> 1024 -> 512 -> 256, in-batch softmax over d = 256 (per-rank negatives), two 10M x 128
    tables with fused row-wise ADAM, sharded as the planner decides (table-wise while the tables can occupy the ranks, row-wise
    beyond: the same rule the weak-scaling block of bench.py runs under), id-column batches (one fixed-size exchange, no host
    sync), peer-memory output exchange, the whole step as ONE CUDA graph.  Timed with CUDA events, max over ranks."""
    import torch.distributed as dist
    Br = batch // world
    cat = ["user_id", "product_id"]
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=D, num_embeddings=rows[i], feature_names=[c]) for i, c in enumerate(cat)]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, list(layers), device=device, precision=precision), loss="in_batch_softmax", precision=precision)
    apply_optimizer_in_backward(tt.RowWiseAdam, task.two_tower.ebc.parameters(), {"lr": 0.01})
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world)).collective_plan(task, tt.get_default_sharders(), dist.GroupMember.WORLD)
    model = tt.DistributedModelParallel(module=task, device=device, plan=plan, sharding_kwargs={"peer_exchange": True} if peer_exchange else None)
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-3))
    model.train()
    g = torch.Generator(device=device).manual_seed(400 + rank)
    raw = [(torch.stack([torch.randint(1, r, (Br,), device=device, generator=g) for r in rows]), torch.zeros(Br, dtype=torch.int32, device=device))
           for _ in range(2)]
    gstep = tt.CudaGraphTrainStep(model, opt, cat, list(rows), Br, device, warmup_steps=2)
    loss = None
    for i in range(2 + 1 + max(warmup, 1)):            # eager warm-up steps, the capture, replays
        loss = gstep(*raw[i % 2])[0]
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = gstep(*raw[i % 2])[0]
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    _hbm, tf, src = _peaks()
    d_out = layers[-1]
    logit_flops = 6.0 * Br * Br * d_out                                                   # per rank, per-rank negatives
    tower_flops = 3 * 2.0 * Br * sum(a_ * b_ for a_, b_ in zip((D,) + tuple(layers[:-1]), layers)) * 2
    out = {"config": 4, "what": "configs[3] on %d GPUs: GLOBAL batch %d (per-rank %d), bf16 towers %d-%s, in-batch softmax d=%d (per-rank "
                                "negatives), 2 x 10M x %d tables, fused row-wise Adam, planner-chosen sharding, peer-memory exchange, whole step "
                                "as one CUDA graph" % (world, batch, Br, D, "-".join(str(x) for x in layers), d_out, D),
           "value": round(batch / ms * 1e3, 1), "unit": "samples/s", "ms_per_step": round(ms, 4), "global_batch": batch, "per_rank_batch": Br,
           "sharding": sorted({ps.sharding_type for tables in plan.plan.values() for ps in tables.values()}), "cuda_graph": gstep.captured,
           "last_loss": float(loss), "per_rank_tflops_credited": round((logit_flops + tower_flops) / (ms * 1e-3) / 1e12, 1),
           "frac_of_sustained_bf16_peak": round((logit_flops + tower_flops) / (ms * 1e-3) / 1e12 / tf, 4), "peak_source": src,
           "flops_credited": "per rank: 6*Br*Br*d logits + 3 GEMMs x layers x 2 towers; communication and the embedding update are inside the time"}
    del model, task, ebc, opt, gstep, raw
    torch.cuda.empty_cache()
    return out


def config5():
    n_items, n_queries, d, k = 10_000_000, 131072, 64, 100
    cat = ["user_id", "product_id"]
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=d, num_embeddings=n_items, feature_names=[c]) for c in cat]
    model = tt.TwoTower(tt.EmbeddingBagCollection(tables=cfgs, device=dev), [128, 64], device=dev, precision="bf16")
    model.eval()
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    items = tt.embed_corpus(model, cat, "product_id", n_items, dev, chunk=1 << 20)
    b.record()
    users = tt.embed_corpus(model, cat, "user_id", n_queries, dev, chunk=1 << 20)
    index = tt.BruteForceIndex(items, precision="bf16")
    torch.cuda.synchronize()
    index.search(users[:4096], k)
    torch.cuda.synchronize()
    b.record()
    scores, ids = index.search(users, k)
    c.record()
    torch.cuda.synchronize()
    ms_search = b.elapsed_time(c)
    return {"config": 5, "what": "item-tower corpus of 10M items (bf16 resident), top-100 for 131072 queries (one GPU's eighth of 1M)",
            "search_ms": round(ms_search, 1), "queries_per_s": round(n_queries / ms_search * 1e3),
            "scoring_tflops": round(2.0 * n_queries * n_items * d / (ms_search * 1e-3) / 1e12, 1),
            "top1_score_mean": round(float(scores[:, 0].mean()), 6)}


if __name__ == "__main__":
    which = [int(x) for x in sys.argv[1:]] or [3, 4, 5]
    for w in which:
        out = {3: config3, 4: config4, 5: config5}[w]()
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()
