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
