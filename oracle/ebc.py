"""Oracle: EmbeddingBagCollection forward, dense gradient, row-wise optimizers.
Test infrastructure only.

Restates ``torchrec.modules.embedding_modules.EmbeddingBagCollection`` as
constructed at /root/reference/03_model_training.py:770-784 and called at
/root/reference/utils/model_training.py:101, and
``torchrec.optim.rowwise_adagrad.RowWiseAdagrad`` fused into backward at
/root/reference/03_model_training.py:791-795.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch

from .kjt import lengths_to_offsets


@dataclass
class TableSpec:
    name: str
    num_embeddings: int
    embedding_dim: int
    feature_names: List[str] = field(default_factory=list)
    pooling: str = "sum"  # "sum" (TorchRec default) | "mean"


def output_layout(tables: Sequence[TableSpec]) -> Tuple[List[str], List[int]]:
    """KeyedTensor column order: tables in config order, features in
    ``feature_names`` order; one ``embedding_dim`` block per feature."""
    keys, dims = [], []
    for t in tables:
        for f in t.feature_names:
            keys.append(f)
            dims.append(t.embedding_dim)
    return keys, dims


def ebc_forward(
    tables: Sequence[TableSpec], weights: Sequence[torch.Tensor],
    kjt_keys: Sequence[str], values: torch.Tensor, lengths: torch.Tensor,
) -> torch.Tensor:
    """Spec form: ``out[b, feat] = sum_{i in bag(feat,b)} W_t[values[i]]``
    (mean: divided by ``len``; empty bag -> zeros).  Returns ``[B, sum D]``."""
    F = len(kjt_keys)
    B = lengths.numel() // F
    offsets = lengths_to_offsets(lengths).to(torch.int64)
    keys, dims = output_layout(tables)
    out = torch.zeros(B, sum(dims), dtype=weights[0].dtype)
    col = 0
    for t, w in zip(tables, weights):
        for feat in t.feature_names:
            f = list(kjt_keys).index(feat)
            s, e = int(offsets[f * B]), int(offsets[(f + 1) * B])
            ids = values[s:e]
            lens = lengths[f * B:(f + 1) * B].to(torch.int64)
            bag = torch.repeat_interleave(torch.arange(B), lens, output_size=ids.numel())
            rows = w[ids]
            pooled = torch.zeros(B, t.embedding_dim, dtype=w.dtype)
            pooled.index_add_(0, bag, rows)
            if t.pooling == "mean":
                pooled = pooled / lens.clamp(min=1).to(w.dtype).unsqueeze(1)
            out[:, col:col + t.embedding_dim] = pooled
            col += t.embedding_dim
    return out


def ebc_forward_torch(
    tables: Sequence[TableSpec], weights: Sequence[torch.Tensor],
    kjt_keys: Sequence[str], values: torch.Tensor, lengths: torch.Tensor,
) -> torch.Tensor:
    """Independent form: literally what unsharded TorchRec executes on CPU --
    one ``torch.nn.functional.embedding_bag(include_last_offset=True)`` per
    feature.  Used to pin :func:`ebc_forward` and as the CPU baseline."""
    F = len(kjt_keys)
    B = lengths.numel() // F
    offsets = lengths_to_offsets(lengths).to(torch.int64)
    outs = []
    for t, w in zip(tables, weights):
        for feat in t.feature_names:
            f = list(kjt_keys).index(feat)
            s = int(offsets[f * B])
            e = int(offsets[(f + 1) * B])
            outs.append(torch.nn.functional.embedding_bag(
                values[s:e], w, offsets[f * B:(f + 1) * B + 1] - s,
                mode=t.pooling, include_last_offset=True))
    return torch.cat(outs, dim=1)


def ebc_dense_grads(
    tables: Sequence[TableSpec], kjt_keys: Sequence[str], values: torch.Tensor,
    lengths: torch.Tensor, grad_out: torch.Tensor,
) -> List[torch.Tensor]:
    """Dense ``[R, D]`` gradient per table: what ``nn.EmbeddingBag(sparse=False)``
    produces on the unsharded path (a row hit k times receives the SUM of the k
    contributions)."""
    F = len(kjt_keys)
    B = lengths.numel() // F
    offsets = lengths_to_offsets(lengths).to(torch.int64)
    grads = []
    col = 0
    for t in tables:
        g = torch.zeros(t.num_embeddings, t.embedding_dim, dtype=grad_out.dtype)
        for feat in t.feature_names:
            f = list(kjt_keys).index(feat)
            s, e = int(offsets[f * B]), int(offsets[(f + 1) * B])
            ids = values[s:e]
            lens = lengths[f * B:(f + 1) * B].to(torch.int64)
            bag = torch.repeat_interleave(torch.arange(B), lens, output_size=ids.numel())
            go = grad_out[:, col:col + t.embedding_dim]
            if t.pooling == "mean":
                go = go / lens.clamp(min=1).to(go.dtype).unsqueeze(1)
            g.index_add_(0, ids, go[bag])
            col += t.embedding_dim
        grads.append(g)
    return grads


def rowwise_adagrad_dense(
    w: torch.Tensor, state_sum: torch.Tensor, grad: torch.Tensor,
    lr: float = 1e-2, eps: float = 1e-10, weight_decay: float = 0.0,
) -> None:
    """``torchrec.optim.rowwise_adagrad`` single-tensor dense step (defaults
    lr=1e-2, lr_decay=0, weight_decay=0, initial_accumulator_value=0,
    eps=1e-10):  ``s += mean_d(g^2)``; ``w -= lr * g / (sqrt(s) + eps)``.
    In place on ``w [R,D]`` and ``state_sum [R]``."""
    if weight_decay != 0.0:
        grad = grad + weight_decay * w
    state_sum.add_(grad.pow(2).mean(dim=1))
    std = state_sum.sqrt().add_(eps)
    w.addcdiv_(grad, std.unsqueeze(1), value=-lr)


def rowwise_adagrad_sparse(
    w: torch.Tensor, state_sum: torch.Tensor, rows: torch.Tensor, row_grads: torch.Tensor,
    lr: float = 1e-2, eps: float = 1e-10,
) -> None:
    """Sparse-exact form (FBGEMM ``EXACT_ROWWISE_ADAGRAD``): ``rows`` are the
    UNIQUE touched rows, ``row_grads[u]`` the summed gradient of ``rows[u]``.
    Equal to the dense form because untouched rows have g = 0."""
    s = state_sum[rows] + row_grads.pow(2).mean(dim=1)
    state_sum[rows] = s
    w[rows] = w[rows] - lr * row_grads / (s.sqrt() + eps).unsqueeze(1)


def rowwise_adam_sparse(
    w: torch.Tensor, m: torch.Tensor, v: torch.Tensor, rows: torch.Tensor,
    row_grads: torch.Tensor, step: int, lr: float = 1e-2, beta1: float = 0.9,
    beta2: float = 0.999, eps: float = 1e-8,
) -> None:
    """Extension (BASELINE config 4): FBGEMM ``PARTIAL_ROWWISE_ADAM``.
    ``m [R,D]`` per element, ``v [R]`` per row, ``step`` = 1-based global
    iteration.  Only touched rows advance."""
    g = row_grads
    v_new = beta2 * v[rows] + (1.0 - beta2) * g.pow(2).mean(dim=1)
    v[rows] = v_new
    m_new = beta1 * m[rows] + (1.0 - beta1) * g
    m[rows] = m_new
    v_hat = v_new / (1.0 - beta2 ** step)
    m_hat = m_new / (1.0 - beta1 ** step)
    w[rows] = w[rows] - lr * m_hat / (v_hat.sqrt() + eps).unsqueeze(1)


def unique_row_grads(dense_grad: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Helper: the unique touched rows (ascending) and their summed gradients."""
    rows = torch.unique(ids, sorted=True)
    return rows, dense_grad[rows]
