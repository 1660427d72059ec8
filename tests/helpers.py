"""Shared builders for the parity tests: the same seeded inputs go to the oracle
(CPU) and to the CUDA path."""
from typing import List, Sequence

import torch

from oracle.ebc import TableSpec


FBGEMM_BUCKETIZE_VECTOR = dict(
    # fbgemm_gpu's own unit-test vector for block_bucketize_sparse_features (fbgemm_gpu/test/sparse_ops_test.py,
    # test_block_bucketize_sparse_features; the dependency pinned at requirements.txt:2 is not installable here, so the vector
    # is quoted from its test suite and RE-DERIVED BY HAND in the comments of test_bucketize_fbgemm_unit_test_vector):
    # T = 4 features, B = 2, my_size = 2, block_sizes = [5, 15, 10, 20]
    lengths=[0, 2, 1, 3, 2, 3, 3, 1],
    indices=[3, 4, 15, 11, 28, 29, 1, 10, 11, 12, 13, 11, 22, 20, 20],
    block_sizes=[5, 15, 10, 20], my_size=2, B=2,
    new_lengths=[0, 2, 0, 1, 1, 0, 1, 0, 0, 0, 1, 2, 1, 3, 2, 1],
    new_indices=[3, 4, 11, 1, 11, 0, 13, 14, 0, 1, 2, 3, 2, 0, 0],
    unbucketize_permute=[0, 1, 5, 2, 6, 7, 3, 8, 9, 10, 11, 4, 12, 13, 14])


def random_kjt(keys: Sequence[str], rows: Sequence[int], batch: int, max_len: int, seed: int,
               empty_frac: float = 0.2, dup_pool: int = 0):
    """Returns (values int64, lengths int32).  ``dup_pool`` > 0 draws ids from a small
    pool so many bags hit the same rows (dedup / summed-gradient case)."""
    g = torch.Generator().manual_seed(seed)
    lens, vals = [], []
    for R in rows:
        ln = torch.randint(0, max_len + 1, (batch,), generator=g)
        ln[torch.rand(batch, generator=g) < empty_frac] = 0
        n = int(ln.sum())
        hi = min(R, dup_pool) if dup_pool else R
        vals.append(torch.randint(0, hi, (n,), generator=g))
        lens.append(ln)
    return torch.cat(vals).to(torch.int64), torch.cat(lens).to(torch.int32)


def make_tables(dims: Sequence[int], rows: Sequence[int], pooling: Sequence[str]) -> List[TableSpec]:
    return [TableSpec(f"t_f{i}", rows[i], dims[i], [f"f{i}"], pooling[i]) for i in range(len(dims))]


def load_reference_golden():
    """tests/golden/reference_train.npz: outputs of the reference's own bodies (tests/golden/make_reference_golden.py)."""
    import os

    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_train.npz"))
    e0, e1, dim, l0, l1, B, steps = (int(x) for x in z["meta"])

    def T_(k):
        return torch.from_numpy(np.asarray(z[k]))

    return {"z": z, "emb": [e0, e1], "dim": dim, "layers": [l0, l1], "B": B, "steps": steps, "lr": float(z["lr"]), "T": T_,
            "init": {k[5:]: T_(k) for k in z.files if k.startswith("init.")},
            "final": {k[6:]: T_(k) for k in z.files if k.startswith("final.")}}


def load_raytune_golden():
    """tests/golden/reference_raytune.npz: the reference's Ray-Tune TwoTower class run on stock torch
    (tests/golden/make_reference_raytune_golden.py)."""
    import os

    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_raytune.npz"))
    G = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if "." not in k}
    G["weights"] = {k[7:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith("weight.")}
    G["grads"] = {k[5:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith("grad.")}
    G.update(feats_u=["u_a", "u_b"], feats_i=["i_a"], dims={"u_a": 36, "u_b": 4, "i_a": 36}, rows={"u_a": 100, "u_b": 7, "i_a": 90},
             layers=[[64, 16], [32, 16]], dense_index=3, dense_dim=5)
    return G
