"""Oracle: exact dot-product top-k and retriever metrics.  Test infrastructure only.

The reference obtains top-k from a remote Databricks Vector Search index
(/root/reference/04_evaluate_retrieval.py:134-141, k=100); the service is not
reproducible, so the oracle is the exact answer with an explicit tie rule.
"""
from typing import Dict, List, Sequence, Tuple

import torch


def exact_topk(queries: torch.Tensor, items: torch.Tensor, k: int, chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """Scores ``queries @ items.T`` accumulated in float64; order = descending
    score, ties broken by LOWER item index.  Returns ``(scores f32 [Q,k],
    indices int64 [Q,k])``."""
    Q = queries.shape[0]
    N = items.shape[0]
    k = min(k, N)
    out_s = torch.empty(Q, k, dtype=torch.float32)
    out_i = torch.empty(Q, k, dtype=torch.int64)
    it64 = items.to(torch.float64)
    for s in range(0, Q, chunk):
        sc = queries[s:s + chunk].to(torch.float64) @ it64.t()
        # stable sort on descending score keeps lower index first among ties
        order = torch.sort(sc, dim=1, descending=True, stable=True).indices[:, :k]
        out_i[s:s + chunk] = order
        out_s[s:s + chunk] = torch.gather(sc, 1, order).to(torch.float32)
    return out_s, out_i


def retrieval_metrics(pred: Sequence[Sequence[int]], targets: Sequence[Sequence[int]], k: int) -> Dict[str, float]:
    """``mlflow.evaluate(model_type="retriever", evaluator_config={"retriever_k": k})``
    as called at 04_evaluate_retrieval.py:202-210: per-row precision_at_k /
    recall_at_k / ndcg_at_k over the first ``k`` retrieved ids, then the mean.
    (mlflow semantics: precision = hits / len(retrieved[:k]); recall = hits /
    len(set(targets)); ndcg with binary relevance, ideal = all targets first, cut at the length of the retrieved list.)"""
    import math
    P: List[float] = []
    R: List[float] = []
    N: List[float] = []
    for p, t in zip(pred, targets):
        p = list(p)[:k]
        tset = set(int(x) for x in t)
        hits = [1.0 if int(x) in tset else 0.0 for x in p]
        nh = sum(hits)
        P.append(nh / len(p) if p else 0.0)
        R.append(nh / len(tset) if tset else 0.0)
        dcg = sum(h / math.log2(i + 2) for i, h in enumerate(hits))
        # a list shorter than k: mlflow hands sklearn's ndcg_score k = len(retrieved[:k]), so the ideal ranking is cut at the
        # number of ids actually retrieved (the reference always asks for, and gets, k = 100 results: 04_evaluate_retrieval.py:134-141)
        ideal = sum(1.0 / math.log2(i + 2) for i in range(min(len(tset), len(p))))
        N.append(dcg / ideal if ideal > 0 else 0.0)
    n = max(len(P), 1)
    return {f"precision_at_{k}": sum(P) / n, f"recall_at_{k}": sum(R) / n, f"ndcg_at_{k}": sum(N) / n}
