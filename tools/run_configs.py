"""Runs the BASELINE.json configurations that are NOT the bench headline, at full size on one B200, a
few steps each, and prints one JSON line per configuration (timing with CUDA events, eager loop).

    python tools/run_configs.py [3] [4] [5]

  3: multi-hot user history (L=20, mean) + product/aisle/department tables, 100M rows x 128 (51 GB each)
  4: batch 262144, bf16 towers 128->1024->512->256, fused row-wise Adam, in-batch softmax d=256
  5: retrieval: item-tower corpus of 10M items, top-100 for 131072 queries (1/8 of the 1M: one GPU's share)
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


def config3():
    B, D, L = 65536, 128, 20
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
    batches = []
    for _ in range(3):
        lens = torch.cat([torch.full((B,), L, dtype=torch.int32, device=dev), torch.ones(3 * B, dtype=torch.int32, device=dev)])
        vals = torch.cat([torch.randint(0, rows["hist"], (B * L,), device=dev, generator=g), torch.randint(0, rows["product"], (B,), device=dev, generator=g),
                          torch.randint(0, 134, (B,), device=dev, generator=g), torch.randint(0, 21, (B,), device=dev, generator=g)])
        batches.append(tt.Batch(torch.zeros(1, device=dev), tt.KeyedJaggedTensor.from_lengths_sync(keys, vals, lens),
                                torch.zeros(B, dtype=torch.int32, device=dev)))
    i = [0]

    def step():
        opt.zero_grad()
        loss, _ = task(batches[i[0] % 3]); i[0] += 1
        loss.backward()
        opt.step()
    ms = timed_steps(step)
    N.enable_timing(True)
    step(); step()
    torch.cuda.synchronize()
    per = {k: round(v["ms"], 4) for k, v in N.timing_summary().items() if k.startswith("tt_ebc")}
    N.enable_timing(False)
    fwd_bytes = B * L * (8 + 4 * D) + 4 * B + 4 * B * D + 3 * (B * (8 + 4 * D) + 4 * B + 4 * B * D)
    return {"config": 3, "what": "L=20 mean history + product/aisle/department, 2 x 100M x 128 fp32 tables (102 GB), B=65536, one GPU",
            "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3), "ebc_ms": per,
            "ebc_forward_gbs": round(fwd_bytes / (per["tt_ebc_forward"] * 1e-3) / 1e9, 1), "hbm_gb_allocated": round(torch.cuda.memory_allocated() / 1e9, 1)}


def config4():
    B, D = 262144, 128
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
    ms = timed_steps(step, steps=3, warmup=1)
    flops = 6.0 * B * B * 256 + 3 * 2 * (128 * 1024 + 1024 * 512 + 512 * 256) * B * 2
    return {"config": 4, "what": "B=262144, bf16 towers 128-1024-512-256, in-batch softmax d=256, fused row-wise Adam, one GPU",
            "ms_per_step": round(ms, 2), "samples_per_s": round(B / ms * 1e3), "credited_tflops": round(flops / (ms * 1e-3) / 1e12, 1)}


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
