"""Retrieval micro-benchmark: tt_score_topk_bf16 on random-normal (trained-like) and degenerate corpora.
    python tools/topk_bench.py [items queries]...   (default: 2M x 16384, 10M x 131072)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import two_tower_recommender_model_b200 as tt


def run(n_items, n_queries, d=64, k=100, kind="randn", reps=2):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    if kind == "randn":
        items = torch.randn(n_items, d, device=dev, generator=g)
        queries = torch.randn(n_queries, d, device=dev, generator=g)
    else:   # untrained-tower-like: every score nearly equal
        items = torch.full((n_items, d), 0.0125, device=dev) + 1e-4 * torch.randn(n_items, d, device=dev, generator=g)
        queries = torch.full((n_queries, d), 0.0125, device=dev)
    index = tt.BruteForceIndex(items, precision="bf16")
    del items
    index.search(queries[:512], k)
    index.search(queries, k, query_chunk=1 << 17)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s, i = index.search(queries, k, query_chunk=1 << 17)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    prof = None
    if os.environ.get("TT_TOPK_PROFILE"):
        import ctypes
        from two_tower_recommender_model_b200 import _native as N
        buf = (ctypes.c_uint64 * 16)()
        N.load().tt_debug_read_counters(buf, 16, 1)
        index.search(queries, k, query_chunk=1 << 17)
        torch.cuda.synchronize()
        N.load().tt_debug_read_counters(buf, 16, 1)
        c = list(buf)
        ep, mm, tm = max(c[4], 1), max(c[11], 1), max(c[13], 1)
        prof = {"epilogue": {"wait_scores": round(c[0] / ep, 3), "wait_tmem_ld": round(c[1] / ep, 3), "compaction": round(c[3] / ep, 3)},
                "issuer": {"wait_items": round(c[8] / mm, 3), "wait_free_tmem": round(c[9] / mm, 3)},
                "producer": {"wait_free_smem": round(c[12] / tm, 3)}, "issuer_clk": c[11], "epilogue_clk_per_warp": c[4]}
    return {"profile": prof, "items": n_items, "queries": n_queries, "k": k, "d": d, "data": kind, "ms": round(best, 3),
            "queries_per_s": round(n_queries / best * 1e3, 1), "tflops": round(2.0 * n_queries * n_items * d / best / 1e9, 1),
            "top1_score_mean": float(s[:, 0].mean())}


if __name__ == "__main__":
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(2_000_000, 16384), (10_000_000, 131072)]
    for n, q in shapes:
        for kind in os.environ.get("TT_TOPK_KINDS", "randn,flat").split(","):
            print(json.dumps(run(n, q, kind=kind)), flush=True)
