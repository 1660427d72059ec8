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


def embed_corpus_sharded(two_tower_model, cat_cols: List[str], key: str, num_embeddings: int, device, pg=None,
                         chunk: int = 1 << 18):
    """Corpus embedding with a SHARDED model (03_model_training.py:1095-1122 run against the trained -- here
    table-wise / row-wise sharded -- tables): rank r embeds the contiguous id range
    ``[r * ceil(N / W), (r + 1) * ceil(N / W))`` through the sharded EmbeddingBagCollection (a collective: every
    rank calls it with the same chunk size; ids past a rank's range are clamped and their rows dropped) and its
    tower.  Returns ``(local_embeddings [n_r, d_out], first_id)``; ``BruteForceIndex.from_sharded`` all-gathers."""
    from torch import distributed as dist
    W, r = (dist.get_world_size(pg), dist.get_rank(pg)) if dist.is_available() and dist.is_initialized() else (1, 0)
    per = -(-num_embeddings // W)
    lo, hi = min(r * per, num_embeddings), min((r + 1) * per, num_embeddings)
    chunk = min(chunk, per)
    outs = []
    for s in range(0, per, chunk):                      # the same number of calls on every rank
        kjt = create_keyed_jagged_tensor(chunk, cat_cols, key, device, start=lo + s)
        kjt._values = kjt._values.clamp_(max=num_embeddings - 1)
        e = process_embeddings(two_tower_model, kjt, key)
        outs.append(e[:max(0, min(chunk, hi - (lo + s)))])
    return (torch.cat(outs, dim=0) if outs else torch.empty(0, 0, device=device)), lo


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

    @classmethod
    def from_sharded(cls, local_item_embeddings: torch.Tensor, pg=None, primary_key: str = "product_id",
                     precision: str = "fp32") -> "BruteForceIndex":
        """Multi-GPU retrieval (BASELINE configs[4]): every rank holds the embeddings of a contiguous id
        range (rank order = id order, see ``embed_corpus_sharded``); the corpus is ALL-GATHERED once (10 M x 64 bf16 =
        1.28 GB) and every rank then answers ITS OWN queries against the full corpus with ``search`` -- queries are
        sharded, so no per-query merge across ranks exists.  ``precision="bf16"`` gathers the bf16 copy."""
        from torch import distributed as dist
        x = local_item_embeddings.contiguous().float()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(pg) == 1:
            return cls(x, primary_key=primary_key, precision=precision)
        W = dist.get_world_size(pg)
        dev = x.device
        sizes = torch.zeros(W, dtype=torch.int64, device=dev)
        sizes[dist.get_rank(pg)] = x.shape[0]
        dist.all_reduce(sizes, group=pg)
        sizes = sizes.tolist()
        mx, d = max(sizes), x.shape[1]
        index = cls.__new__(cls)
        index._ids, index._pk, index._precision = None, primary_key, precision
        if precision == "bf16":
            from .functional import cast_bf16
            xb = cast_bf16(x) if x.shape[0] > 0 else torch.empty(0, (d + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)
            pw = (d + 7) // 8 * 8
            pad = torch.zeros(mx, pw, dtype=torch.bfloat16, device=dev)
            if x.shape[0] > 0:
                pad[:x.shape[0], :d] = xb
            allb = torch.empty(W * mx, pw, dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(allb, pad, group=pg)
            full = torch.cat([allb[w * mx:w * mx + sizes[w]] for w in range(W)]) if any(s != mx for s in sizes) else allb
            index._items_bf16 = full[:, :d]
            index._items = None
            index._n, index._dev = full.shape[0], dev
        else:
            pad = torch.zeros(mx, d, dtype=torch.float32, device=dev)
            pad[:x.shape[0]] = x
            allf = torch.empty(W * mx, d, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allf, pad, group=pg)
            full = torch.cat([allf[w * mx:w * mx + sizes[w]] for w in range(W)]) if any(s != mx for s in sizes) else allf
            index._items, index._items_bf16 = full.contiguous(), None
            index._n, index._dev = full.shape[0], dev
        return index

    def search(self, query_embeddings: torch.Tensor, num_results: int = 100, query_chunk: int = 1 << 16):
        """Batched top-k: returns ``(scores [Q,k] f32, ids [Q,k] int64)``."""
        dev = self._items.device if self._items is not None else self._items_bf16.device
        n_items = self._items.shape[0] if self._items is not None else self._items_bf16.shape[0]
        q = query_embeddings.to(dev).float().contiguous()
        k = min(num_results, n_items)
        if q.shape[0] == 0:            # an empty query batch: nothing to launch
            return (torch.empty(0, k, dtype=torch.float32, device=dev), torch.empty(0, k, dtype=torch.int64, device=dev))
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
        dev = self._items.device if self._items is not None else self._items_bf16.device
        q = torch.as_tensor(query_vector, dtype=torch.float32, device=dev).view(1, -1)
        scores, idx = self.search(q, num_results)
        rows = [[int(i), float(s)] for i, s in zip(idx[0].tolist(), scores[0].tolist())]
        return {"manifest": {"column_count": 2, "columns": [{"name": self._pk}, {"name": "score"}]},
                "result": {"row_count": len(rows), "data_array": rows}}


class CorpusShardedIndex:
    """Exact inner-product retrieval when the corpus does NOT fit one GPU ("fall back to corpus-sharding
    + [Q, k]-per-rank all-to-all merge"): every rank keeps the embeddings of ITS contiguous id range (what
    ``embed_corpus_sharded`` returns), the queries of all ranks are all-gathered, each rank scores every query against
    its shard with the single-GPU top-k kernel (``item_index_base`` = first id of the shard, so the kernel already
    returns global ids), the per-shard top-k lists travel to the rank that owns the query in ONE all-to-all, and the W
    lists are merged there (descending score, ties -> lower id: the order of ``BruteForceIndex`` on the whole corpus).
    The query count must be the same on every rank (pad the last block); ``BruteForceIndex.from_sharded`` -- corpus
    all-gathered, queries sharded, no merge -- is the faster layout whenever the corpus fits."""

    def __init__(self, local_item_embeddings: torch.Tensor, first_id: int, pg=None, precision: str = "fp32") -> None:
        self._items = local_item_embeddings.contiguous().float()
        self._first, self._pg, self._precision = int(first_id), pg, precision
        self._items_bf16 = None
        if precision == "bf16" and self._items.shape[0] > 0:
            from .functional import cast_bf16
            self._items_bf16 = cast_bf16(self._items)

    def search(self, query_embeddings: torch.Tensor, num_results: int = 100):
        """``query_embeddings`` [Q, d]: THIS rank's queries.  Returns ``(scores [Q, k] f32, ids [Q, k] int64)`` over the
        whole (sharded) corpus; a collective -- every rank calls it with the same Q and k."""
        from torch import distributed as dist
        pg = self._pg
        W = dist.get_world_size(pg) if dist.is_available() and dist.is_initialized() else 1
        q = query_embeddings.to(self._items.device).float().contiguous()
        Q, k = q.shape[0], int(num_results)
        if Q == 0:                     # the same on every rank: nothing to exchange
            return (torch.empty(0, k, dtype=torch.float32, device=q.device), torch.empty(0, k, dtype=torch.int64, device=q.device))
        if W > 1:
            allq = q.new_empty(W * Q, q.shape[1])
            dist.all_gather_into_tensor(allq, q, group=pg)
        else:
            allq = q
        n_local = self._items.shape[0]
        kl = min(k, n_local)
        part_s = torch.full((W * Q, k), float("-inf"), dtype=torch.float32, device=q.device)
        part_i = torch.full((W * Q, k), (1 << 62), dtype=torch.int64, device=q.device)
        if kl > 0:
            sc, ix = score_topk(allq, self._items, kl, item_index_base=self._first, precision=self._precision,
                                items_bf16=self._items_bf16)
            part_s[:, :kl], part_i[:, :kl] = sc, ix
        if W > 1:
            recv_s, recv_i = torch.empty_like(part_s), torch.empty_like(part_i)
            dist.all_to_all_single(recv_s, part_s, group=pg)      # block r of the send buffer = rank r's queries
            dist.all_to_all_single(recv_i, part_i, group=pg)
            cand_s = recv_s.view(W, Q, k).permute(1, 0, 2).reshape(Q, W * k)
            cand_i = recv_i.view(W, Q, k).permute(1, 0, 2).reshape(Q, W * k)
        else:
            cand_s, cand_i = part_s, part_i
        # merge: by id ascending first, then a STABLE sort by descending score -> equal scores keep the lower id first
        by_id = torch.argsort(cand_i, dim=1, stable=True)
        cand_s, cand_i = cand_s.gather(1, by_id), cand_i.gather(1, by_id)
        order = torch.argsort(cand_s, dim=1, descending=True, stable=True)[:, :k]
        return cand_s.gather(1, order), cand_i.gather(1, order)


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
