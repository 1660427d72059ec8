"""Generates tests/golden/golden.json.

The reference ships no golden vectors and TorchRec/FBGEMM are not installable here
(SURVEY.md section 8c), so the fixtures come from two sources, both recorded per case:

  "hand"  -- known answers written out by hand below (tiny cases);
  "torch" -- outputs of STOCK torch ops that the unsharded TorchRec CPU path is
             made of (nn.functional.embedding_bag, F.linear+relu,
             F.binary_cross_entropy_with_logits, F.cross_entropy, torch.optim.Adam,
             torch.sort) -- NOT outputs of oracle/.

Run:  python tests/golden/make_golden.py   (deterministic; rewrites golden.json)
"""
import json
import os

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


def L(t):
    return t.tolist() if isinstance(t, torch.Tensor) else t


def main():
    torch.manual_seed(1234)
    G = {}

    # ---- hand: transform_to_torchrec_batch -------------------------------
    G["transform_kat"] = {
        "source": "hand",
        "batch": {"user_id": [1, 0, 3], "product_id": [10, 20, 0], "label": [1, 0, 1]},
        "cat_cols": ["user_id", "product_id"], "emb_counts": [4, 16],
        "values": [1, 3, 10, 4], "lengths": [1, 0, 1, 1, 1, 0], "labels": [1, 0, 1],
        "offsets": [0, 1, 1, 2, 3, 4, 4], "length_per_key": [2, 2],
    }
    # ---- hand: permute_2D ------------------------------------------------------------------
    G["permute_kat"] = {
        "source": "hand", "T": 3, "B": 2, "lengths": [2, 0, 1, 1, 0, 3], "values": [5, 6, 7, 8, 9, 10, 11],
        "permute": [2, 0, 2],
        "out_lengths": [0, 3, 2, 0, 0, 3], "out_values": [9, 10, 11, 5, 6, 9, 10, 11],
    }
    # ---- hand: block_bucketize (R=10, W=2 -> block 5; R=7, W=2 -> block 4) -------------------
    # F=2, B=2; bags: f0b0=[9,1,5] f0b1=[] f1b0=[6] f1b1=[3,4]
    G["bucketize_kat"] = {
        "source": "hand", "F": 2, "B": 2, "W": 2, "rows": [10, 7],
        "lengths": [3, 0, 1, 2], "values": [9, 1, 5, 6, 3, 4],
        # output order (w, f, b): w0f0b0=[1] w0f0b1=[] w0f1b0=[] w0f1b1=[3] | w1f0b0=[9-5,5-5] w1f0b1=[] w1f1b0=[6-4] w1f1b1=[4-4]
        "new_lengths": [1, 0, 0, 1, 2, 0, 1, 1], "new_values": [1, 3, 4, 0, 2, 0],
        "unbucketize": [2, 0, 3, 4, 1, 5],
    }
    # ---- hand: row-wise Adagrad, one row hit twice -------------------------------------------
    # W = [[1,2],[3,4]], ids [0,0] with grads [1,1] and [1,3] -> G0 = [2,4]; s0 = mean(4,16)=10
    # w0 -= 0.1 * G0 / (sqrt(10)+1e-10)
    s = 10.0 ** 0.5
    G["rowwise_adagrad_kat"] = {
        "source": "hand", "weights": [[1.0, 2.0], [3.0, 4.0]], "ids": [0, 0], "grads": [[1.0, 1.0], [1.0, 3.0]],
        "lr": 0.1, "eps": 1e-10, "sum_after": [10.0, 0.0],
        "weights_after": [[1.0 - 0.1 * 2.0 / s, 2.0 - 0.1 * 4.0 / s], [3.0, 4.0]],
    }
    # ---- hand: partial row-wise Adam, step 1 ---------------------------------------------------
    # g=[2,4]: v = 0.001*10 = 0.01; v_hat = 0.01/0.001 = 10; m = 0.1*g; m_hat = g; w -= lr*g/(sqrt(10)+eps)
    G["rowwise_adam_kat"] = {
        "source": "hand", "weights": [[1.0, 2.0], [3.0, 4.0]], "ids": [0, 0], "grads": [[1.0, 1.0], [1.0, 3.0]],
        "lr": 0.1, "eps": 1e-8, "beta1": 0.9, "beta2": 0.999, "step": 1,
        "v_after": [0.01, 0.0], "m_after": [[0.2, 0.4], [0.0, 0.0]],
        "weights_after": [[1.0 - 0.1 * 2.0 / (s + 1e-8), 2.0 - 0.1 * 4.0 / (s + 1e-8)], [3.0, 4.0]],
    }
    # ---- hand: top-k with ties (lower index wins) ----------------------------------------------
    G["topk_ties_kat"] = {
        "source": "hand", "queries": [[1.0, 0.0]], "items": [[0.5, 9.0], [1.0, 1.0], [0.5, -3.0], [1.0, 0.0], [-1.0, 0.0]],
        "k": 4, "indices": [[1, 3, 0, 2]], "scores": [[1.0, 1.0, 0.5, 0.5]],
    }
    # ---- torch: embedding_bag sum / mean with empty bags and duplicates -------------------------
    W0 = torch.randn(6, 4); W1 = torch.randn(5, 4)
    lengths = torch.tensor([2, 0, 3, 1, 1, 0], dtype=torch.int32)          # F=2, B=3
    values = torch.tensor([1, 1, 5, 0, 2, 4, 4])
    off = torch.zeros(7, dtype=torch.int64); off[1:] = torch.cumsum(lengths.long(), 0)
    p0 = F.embedding_bag(values[off[0]:off[3]], W0, off[0:4] - off[0], mode="sum", include_last_offset=True)
    p1 = F.embedding_bag(values[off[3]:off[6]], W1, off[3:7] - off[3], mode="mean", include_last_offset=True)
    G["ebc_forward_torch"] = {
        "source": "torch", "tables": [{"name": "t_a", "rows": 6, "dim": 4, "features": ["a"], "pooling": "sum"},
                                      {"name": "t_b", "rows": 5, "dim": 4, "features": ["b"], "pooling": "mean"}],
        "weights": [L(W0), L(W1)], "keys": ["a", "b"], "values": L(values), "lengths": L(lengths),
        "pooled": L(torch.cat([p0, p1], dim=1)),
    }
    # dense grads through autograd of the same op
    W0g = W0.clone().requires_grad_(True); W1g = W1.clone().requires_grad_(True)
    q0 = F.embedding_bag(values[off[0]:off[3]], W0g, off[0:4] - off[0], mode="sum", include_last_offset=True)
    q1 = F.embedding_bag(values[off[3]:off[6]], W1g, off[3:7] - off[3], mode="mean", include_last_offset=True)
    go = torch.randn(3, 8)
    torch.cat([q0, q1], dim=1).backward(go)
    G["ebc_backward_torch"] = {"source": "torch", "grad_out": L(go), "grads": [L(W0g.grad), L(W1g.grad)]}
    # ---- torch: MLP (relu after every layer) -----------------------------------------------------
    x = torch.randn(5, 4); w1 = torch.randn(6, 4); b1 = torch.randn(6); w2 = torch.randn(3, 6); b2 = torch.randn(3)
    G["mlp_torch"] = {"source": "torch", "x": L(x), "layers": [[L(w1), L(b1)], [L(w2), L(b2)]],
                      "y": L(torch.relu(F.linear(torch.relu(F.linear(x, w1, b1)), w2, b2)))}
    # ---- torch: losses ---------------------------------------------------------------------------
    q = torch.randn(6, 3); c = torch.randn(6, 3); y = torch.tensor([1, 0, 0, 1, 1, 0], dtype=torch.int32)
    logits = (q * c).sum(dim=1)
    G["bce_torch"] = {"source": "torch", "q": L(q), "c": L(c), "labels": L(y), "logits": L(logits),
                      "loss": float(F.binary_cross_entropy_with_logits(logits, y.float()))}
    G["softmax_torch"] = {"source": "torch", "q": L(q), "c": L(c), "temperature": 0.5,
                          "loss": float(F.cross_entropy((q @ c.t()) / 0.5, torch.arange(6)))}
    # ---- torch: Adam, 3 steps ----------------------------------------------------------------------
    p = torch.randn(7, requires_grad=True); p0_ = p.detach().clone()
    opt = torch.optim.Adam([p], lr=0.05)
    gs = []
    for _ in range(3):
        g = torch.randn(7); gs.append(L(g)); p.grad = g.clone(); opt.step()
    G["adam_torch"] = {"source": "torch", "p0": L(p0_), "grads": gs, "lr": 0.05, "p3": L(p.detach())}

    with open(os.path.join(os.environ.get("TT_GOLDEN_OUT", HERE), "golden.json"), "w") as f:      # TT_GOLDEN_OUT: tests/test_golden_provenance.py
        json.dump(G, f, indent=1)
    print("wrote", len(G), "cases")


if __name__ == "__main__":
    main()
