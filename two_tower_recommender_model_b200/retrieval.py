"""Retrieval evaluation path: corpus/query embedding and exact top-k.

* ``create_keyed_jagged_tensor`` / ``process_embeddings`` follow
  /root/reference/03_model_training.py:1056-1122 (KJT with ``values=arange(N)`` and
  lengths ``[0]*N+[1]*N`` for the item key or ``[1]*N+[0]*N`` for the user key;
  ``ebc(kjt)[key]`` -> that tower's MLP under ``no_grad``).
* ``BruteForceIndex.similarity_search`` stands in for the remote Databricks Vector
  Search call at /root/reference/04_evaluate_retrieval.py:134-141
  (``index.similarity_search(query_vector=, columns=, num_results=k)``) and returns
  the same response shape; the batched form ``search`` is what a GPU wants.
  Scores are inner products; ordering is descending score, ties -> lower id.
* ``retrieval_metrics`` reproduces the three numbers
  ``mlflow.evaluate(model_type="retriever", evaluator_config={"retriever_k": k})``
  reports (04_evaluate_retrieval.py:202-226).
"""
from typing import Dict, List, Optional, Sequence, Union

import torch

from .functional import score_topk
from .sparse.jagged_tensor import KeyedJaggedTensor


def create_keyed_jagged_tensor(num_embeddings: int, cat_cols: List[str], key: str, device=None,
                               start: int = 0) -> KeyedJaggedTensor:
    """Reference signature plus ``start`` (so a 10M-item corpus can be embedded in
    chunks).  ``cat_cols`` must be ``[query_key, candidate_key]``; ``key`` picks
    which of the two gets one id per sample, the other key gets empty bags."""
    device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if key not in cat_cols:
        raise ValueError(f"Unsupported key: {key}")
    n = num_embeddings
    values = torch.arange(start, start + n, device=device, dtype=torch.int64)
    lengths = torch.zeros(len(cat_cols) * n, dtype=torch.int32, device=device)
    f = cat_cols.index(key)
    lengths[f * n:(f + 1) * n] = 1
    return KeyedJaggedTensor(keys=cat_cols, values=values, lengths=lengths)


def process_embeddings(two_tower_model, kjt: KeyedJaggedTensor, lookup_column: str) -> torch.Tensor:
    """``ebc(kjt)[lookup_column]`` through the tower that owns that feature."""
    with torch.no_grad():
        lookups = two_tower_model.ebc(kjt)
        if lookup_column in two_tower_model._candidate_feature_names:
            return two_tower_model.candidate_proj(lookups[lookup_column])
        if lookup_column in two_tower_model._feature_names_query:
            return two_tower_model.query_proj(lookups[lookup_column])
    raise ValueError(f"Unsupported key: {lookup_column}")


def embed_corpus(two_tower_model, cat_cols: List[str], key: str, num_embeddings: int, device,
                 chunk: int = 1 << 20) -> torch.Tensor:
    """All ``num_embeddings`` rows of ``key``'s table through its tower, chunked."""
    outs = []
    for s in range(0, num_embeddings, chunk):
        n = min(chunk, num_embeddings - s)
        outs.append(process_embeddings(two_tower_model, create_keyed_jagged_tensor(n, cat_cols, key, device, start=s), key))
    return torch.cat(outs, dim=0)


class BruteForceIndex:
    """Exact inner-product index over an item-embedding corpus resident in HBM."""

    def __init__(self, item_embeddings: torch.Tensor, item_ids: Optional[torch.Tensor] = None,
                 primary_key: str = "product_id", precision: str = "fp32") -> None:
        """``precision="bf16"``: the corpus is kept as a bf16 copy in HBM and scored on the tensor cores
        (embedding dim <= 64); ``"fp32"``: exact fp32 CUDA-core scoring."""
        self._items = item_embeddings.contiguous().float()
        self._ids = item_ids
        self._pk = primary_key
        self._precision = precision
        self._items_bf16 = None
        if precision == "bf16":
            from .functional import cast_bf16
            self._items_bf16 = cast_bf16(self._items)

    def search(self, query_embeddings: torch.Tensor, num_results: int = 100, query_chunk: int = 1 << 16):
        """Batched top-k: returns ``(scores [Q,k] f32, ids [Q,k] int64)``."""
        q = query_embeddings.to(self._items.device).float().contiguous()
        k = min(num_results, self._items.shape[0])
        s_out, i_out = [], []
        for s in range(0, q.shape[0], query_chunk):
            sc, ix = score_topk(q[s:s + query_chunk], self._items, k, precision=self._precision, items_bf16=self._items_bf16)
            s_out.append(sc)
            i_out.append(ix)
        scores, idx = torch.cat(s_out), torch.cat(i_out)
        if self._ids is not None:
            idx = self._ids.to(idx.device)[idx]
        return scores, idx

    def similarity_search(self, query_vector: Union[Sequence[float], torch.Tensor], columns: Optional[List[str]] = None,
                          num_results: int = 100, **unused) -> Dict:
        """One query, Vector-Search-shaped response (04_evaluate_retrieval.py:117-123
        reads ``response['manifest']['columns']`` and ``response['result']['data_array']``)."""
        q = torch.as_tensor(query_vector, dtype=torch.float32, device=self._items.device).view(1, -1)
        scores, idx = self.search(q, num_results)
        rows = [[int(i), float(s)] for i, s in zip(idx[0].tolist(), scores[0].tolist())]
        return {"manifest": {"column_count": 2, "columns": [{"name": self._pk}, {"name": "score"}]},
                "result": {"row_count": len(rows), "data_array": rows}}


def retrieval_metrics(pred_ids: torch.Tensor, targets: Sequence[Sequence[int]], k: int) -> Dict[str, float]:
    """precision_at_k / recall_at_k / ndcg_at_k (mean over rows), computed on the
    device: ``pred_ids`` is ``[Q, >=k]`` int64, ``targets[i]`` the relevant ids of row i."""
    dev = pred_ids.device
    Q = pred_ids.shape[0]
    pred = pred_ids[:, :k]
    lens = torch.tensor([len(set(int(x) for x in t)) for t in targets], device=dev)
    maxlen = int(lens.max().item()) if Q > 0 else 0
    tgt = torch.full((Q, max(maxlen, 1)), -(1 << 62), dtype=torch.int64, device=dev)
    for i, t in enumerate(targets):
        u = sorted(set(int(x) for x in t))
        if u:
            tgt[i, :len(u)] = torch.tensor(u, dtype=torch.int64, device=dev)
    hits = (pred.unsqueeze(2) == tgt.unsqueeze(1)).any(dim=2).float()  # [Q, k]
    nh = hits.sum(dim=1)
    kk = pred.shape[1]
    disc = 1.0 / torch.log2(torch.arange(kk, device=dev, dtype=torch.float32) + 2.0)
    dcg = (hits * disc).sum(dim=1)
    cum = torch.cat([torch.zeros(1, device=dev), torch.cumsum(disc, 0)])
    ideal = cum[torch.clamp(lens, max=kk)]
    ndcg = torch.where(ideal > 0, dcg / ideal.clamp(min=1e-30), torch.zeros_like(dcg))
    prec = nh / max(kk, 1)
    rec = torch.where(lens > 0, nh / lens.clamp(min=1).float(), torch.zeros_like(nh))
    n = max(Q, 1)
    return {f"precision_at_{k}": float(prec.sum() / n), f"recall_at_{k}": float(rec.sum() / n),
            f"ndcg_at_{k}": float(ndcg.sum() / n)}
