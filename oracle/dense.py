"""Oracle: tower MLP, losses, dense Adam.  Test infrastructure only."""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def mlp_forward(x: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]]) -> torch.Tensor:
    """``torchrec.modules.mlp.MLP`` (utils/model_training.py:95-96): a stack of
    ``Perceptron`` = ``relu(Linear(x))`` with the activation after EVERY layer,
    the last one included.  ``layers[i] = (weight [out,in], bias [out])``."""
    for w, b in layers:
        x = torch.relu(F.linear(x, w, b))
    return x


def dot_bce_loss(q: torch.Tensor, c: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/model_training.py:136-140: ``logits = (q*c).sum(1).squeeze()``;
    ``BCEWithLogitsLoss()(logits, labels.float())`` (mean).  Written out:
    ``mean(max(x,0) - x*y + log1p(exp(-|x|)))``."""
    logits = (q * c).sum(dim=1).squeeze()
    y = labels.to(logits.dtype)
    loss = (logits.clamp(min=0) - logits * y + torch.log1p(torch.exp(-logits.abs()))).mean()
    return loss, logits


def in_batch_softmax_loss(q: torch.Tensor, c: torch.Tensor, temperature: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Extension named by BASELINE.json (not in the reference): sampled-softmax
    with in-batch negatives, ``CE(q @ c.T / temperature, arange(B))`` (mean).
    Returns ``(loss, diag_logits)``."""
    s = (q @ c.t()) / temperature
    lse = torch.logsumexp(s, dim=1)
    diag = s.diagonal()
    return (lse - diag).mean(), diag


def adam_step(
    p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
    lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
) -> None:
    """``torch.optim.Adam`` defaults (03_model_training.py:826-829), single
    tensor, in place.  ``step`` is 1-based."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


class _ChunkedInBatchSoftmax(torch.autograd.Function):
    """Same loss and gradients as :func:`in_batch_softmax_loss`, computed over row chunks so that the
    ``[B, B]`` logits never exist at once (at B = 65 536 they are 17 GB in fp32, and autograd would
    keep several copies).  Used by the CPU baseline at BASELINE configs[1]'s full batch; checked
    equal to the plain form in tests/test_oracle_golden.py."""

    @staticmethod
    def forward(ctx, q, c, temperature: float, chunk: int):
        B = q.shape[0]
        inv_t = 1.0 / temperature
        dq = torch.empty_like(q)
        dc = torch.zeros_like(c)
        loss = torch.zeros((), dtype=torch.float64)
        diag = torch.empty(B, dtype=q.dtype)
        for s in range(0, B, chunk):
            e = min(B, s + chunk)
            logits = (q[s:e] @ c.t()) * inv_t                         # [chunk, B]
            lse = torch.logsumexp(logits, dim=1)
            d = logits[torch.arange(e - s), torch.arange(s, e)]
            diag[s:e] = d
            loss += (lse - d).double().sum()
            p = torch.exp(logits - lse.unsqueeze(1))                  # softmax rows
            p[torch.arange(e - s), torch.arange(s, e)] -= 1.0         # dL/dlogits * B
            p *= inv_t / B
            dq[s:e] = p @ c
            dc += p.t() @ q[s:e]
        ctx.save_for_backward(dq, dc)
        ctx.mark_non_differentiable(diag)
        return (loss / B).to(q.dtype), diag

    @staticmethod
    def backward(ctx, g_loss, _g_diag):
        dq, dc = ctx.saved_tensors
        return dq * g_loss, dc * g_loss, None, None


def in_batch_softmax_loss_chunked(q: torch.Tensor, c: torch.Tensor, temperature: float = 1.0,
                                  chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    return _ChunkedInBatchSoftmax.apply(q, c, temperature, chunk)
