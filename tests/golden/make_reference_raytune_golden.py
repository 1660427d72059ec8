"""Generates tests/golden/reference_raytune.npz by executing the reference's Ray-Tune ``TwoTower`` class.

    python tests/golden/make_reference_raytune_golden.py     (build container only: reads /root/reference)

The class definition at /root/reference/ray_tune_optuna_tuning_alex_test.py:181-306 (several features per tower, one layer
stack per tower, dense features concatenated to the tower inputs) is extracted by ``ast`` -- the rest of that file is a
Databricks / Ray driver -- and run on CPU over the stock-torch stand-ins of make_reference_golden.py (nn.EmbeddingBag,
relu(nn.Linear)); loss = the base task's ``BCEWithLogitsLoss`` on ``(q * c).sum(1)`` (utils/model_training.py:136-140).
Stored: the batch (KJT values / lengths, dense features, labels), seeded weights under THIS package's key names
(the reference calls its towers user_proj / item_proj), q, c, logits, loss and the gradient of every parameter.
Shape = tests/test_gpu_train.py::test_ray_tune_variant_towers."""
import ast
import os
import sys
from typing import List, Optional, Tuple  # noqa: F401  (names the extracted class's annotations use)

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as S  # noqa: E402  -- the stand-ins

REF = "/root/reference/ray_tune_optuna_tuning_alex_test.py"
FEATS_U, FEATS_I = ["u_a", "u_b"], ["i_a"]
DIMS = {"u_a": 36, "u_b": 4, "i_a": 36}
ROWS = {"u_a": 100, "u_b": 7, "i_a": 90}
LAYERS, DENSE_INDEX, DENSE_DIM, B = [[64, 16], [32, 16]], 3, 5, 97


def main():
    with open(REF) as f:
        tree = ast.parse(f.read(), REF)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TwoTower"]
    assert len(cls) == 1
    ns = {"torch": torch, "nn": nn, "List": List, "Optional": Optional, "Tuple": Tuple, "MLP": S.MLP, "Batch": S.Batch,
          "EmbeddingBagCollection": S.EmbeddingBagCollection}
    exec(compile(ast.Module(body=cls, type_ignores=[]), REF, "exec"), ns)
    keys = list(DIMS)
    ebc = S.EmbeddingBagCollection([S.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=DIMS[k], num_embeddings=ROWS[k], feature_names=[k])
                                    for k in keys])
    user_in = sum(DIMS[k] for k in FEATS_U) + DENSE_INDEX
    item_in = sum(DIMS[k] for k in FEATS_I) + (DENSE_DIM - DENSE_INDEX)
    model = ns["TwoTower"](ebc, LAYERS, [user_in, item_in], FEATS_U, FEATS_I, dense_index=DENSE_INDEX, device=torch.device("cpu"))
    g = torch.Generator().manual_seed(20261019)
    with torch.no_grad():
        for name, p in model.named_parameters():
            bound = (1.0 / p.shape[0] ** 0.5) if "embedding_bags" in name else ((2.0 / p.shape[1] ** 0.5) if p.dim() == 2 else 0.1)
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    lens, vals = [], []
    for k in keys:                                                         # up to 3 ids per bag, ~20 % empty bags
        ln = torch.randint(0, 4, (B,), generator=g)
        ln[torch.rand(B, generator=g) < 0.2] = 0
        vals.append(torch.randint(0, ROWS[k], (int(ln.sum()),), generator=g))
        lens.append(ln)
    v, l = torch.cat(vals).to(torch.int64), torch.cat(lens).to(torch.int32)
    dense = torch.randn(B, DENSE_DIM, generator=g)
    labels = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
    batch = S.Batch(dense, S.KeyedJaggedTensor(keys, v, l), labels)
    q, c = model(batch)                                                    # the reference's forward
    logits = (q * c).sum(dim=1).squeeze()
    loss = nn.BCEWithLogitsLoss()(logits, labels.float())
    loss.backward()
    rename = {"user_proj": "query_proj", "item_proj": "candidate_proj"}
    out = {"values": v.numpy(), "lengths": l.numpy(), "dense": dense.numpy(), "labels": labels.numpy(),
           "q": q.detach().numpy(), "c": c.detach().numpy(), "logits": logits.detach().numpy(), "loss": loss.detach().numpy()}
    for name, p in model.named_parameters():
        head, rest = name.split(".", 1)
        ours = f"{rename.get(head, head)}.{rest}"
        out["weight." + ours] = p.detach().numpy()
        out["grad." + ours] = p.grad.numpy()
    path = os.path.join(os.environ.get("TT_GOLDEN_OUT", HERE), "reference_raytune.npz")      # TT_GOLDEN_OUT: write elsewhere (tests/test_golden_provenance.py)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, loss {float(loss):.6f}, |logits| max {float(logits.abs().max()):.3f}")


if __name__ == "__main__":
    main()
