"""Oracle: integer / jagged bookkeeping (bit-exact ops).  Test infrastructure only.

KeyedJaggedTensor layout restated from TorchRec (key-major): ``lengths`` is
``[F*B]`` with ``lengths[f*B + b]``, ``values`` concatenated in the same order,
``offsets = cat([0], cumsum(lengths))``.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch


def transform_to_torchrec_batch(
    batch: Dict[str, Sequence[int]],
    cat_cols: List[str],
    num_embeddings_per_feature: List[int],
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Follows /root/reference/utils/model_training.py:43-69 literally.

    A falsy id (0) gives an EMPTY bag; every other id is ``id % num_embeddings``
    with bag length 1.  Returns ``(values int64, lengths int32, labels int32)``.
    Pure-Python double loop on purpose: it is the reference's own algorithm.
    """
    kjt_values: List[int] = []
    kjt_lengths: List[int] = []
    for col_idx, col_name in enumerate(cat_cols):
        for value in batch[col_name]:
            value = int(value)
            if value:
                kjt_values.append(value % num_embeddings_per_feature[col_idx])
                kjt_lengths.append(1)
            else:
                kjt_lengths.append(0)
    values = torch.tensor(kjt_values, dtype=torch.int64)
    lengths = torch.tensor(kjt_lengths, dtype=torch.int32)
    labels = torch.tensor([int(x) for x in batch["label"]], dtype=torch.int32)
    return values, lengths, labels


def transform_row_shard(
    batch: Dict[str, Sequence[int]],
    cat_cols: List[str],
    num_embeddings_per_feature: List[int],
    world_size: int,
    rank: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """What one rank of a ROW-WISE sharded table sees of a batch: the reference transform
    (utils/model_training.py:43-61) followed by the bucket ``rank`` of TorchRec's row-wise input dist
    (``block = ceil(R / W)``; ids with ``id // block == rank`` stay, as ``id - rank * block``; every other bag is
    empty on this rank).  Returns ``(values int64, lengths int32)`` in the same key-major order.  This is the
    semantics of ``tt_kjt_from_columns_range``; it equals slicing ``block_bucketize_sparse_features`` at ``rank``."""
    values: List[int] = []
    lengths: List[int] = []
    for col_idx, col_name in enumerate(cat_cols):
        rows = int(num_embeddings_per_feature[col_idx])
        block = -(-rows // world_size)
        for value in batch[col_name]:
            value = int(value)
            keep = False
            if value:
                r = value % rows
                keep = r // block == rank
            if keep:
                values.append(r - rank * block)
                lengths.append(1)
            else:
                lengths.append(0)
    return torch.tensor(values, dtype=torch.int64), torch.tensor(lengths, dtype=torch.int32)


def lengths_to_offsets(lengths: torch.Tensor) -> torch.Tensor:
    """``fbgemm::asynchronous_complete_cumsum`` as used by
    ``KeyedJaggedTensor.from_lengths_sync`` (utils/model_training.py:57)."""
    out = torch.zeros(lengths.numel() + 1, dtype=lengths.dtype)
    out[1:] = torch.cumsum(lengths, 0)
    return out


def permute_2d_sparse_data(
    permute: Sequence[int], lengths: torch.Tensor, values: torch.Tensor,
    weights: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """``fbgemm::permute_2D_sparse_data`` (used by ``KJT.permute`` and the
    KJT all-to-all): ``lengths`` is ``[T, B]``; output segment ``i`` is input
    segment ``permute[i]`` (repeats allowed), jagged values moved with it."""
    T, B = lengths.shape
    offsets = lengths_to_offsets(lengths.reshape(-1)).to(torch.int64)
    out_len = torch.empty((len(permute), B), dtype=lengths.dtype)
    out_vals: List[torch.Tensor] = []
    out_w: List[torch.Tensor] = []
    for i, src in enumerate(permute):
        out_len[i] = lengths[src]
        s, e = int(offsets[src * B]), int(offsets[(src + 1) * B])
        out_vals.append(values[s:e])
        if weights is not None:
            out_w.append(weights[s:e])
    pv = torch.cat(out_vals) if out_vals else values[:0]
    pw = (torch.cat(out_w) if out_w else weights[:0]) if weights is not None else None
    return out_len, pv, pw


def _bucket_of(v: int, block: int, world_size: int) -> Tuple[int, int]:
    """(bucket, local index) of one id.  fbgemm reads ids as unsigned (``uindex_t``) and sends an id past
    ``block * W`` to bucket ``id % W`` with local index ``id / W`` -- so a bucket is always in ``[0, W)``."""
    u = v % (1 << 64)
    if block > 0 and u < block * world_size:
        return u // block, u % block
    q = u // world_size
    return u % world_size, q - (1 << 64) if q >= (1 << 63) else q


def block_bucketize_sparse_features(
    lengths: torch.Tensor, values: torch.Tensor, num_rows_per_feature: Sequence[int],
    world_size: int, batch_size: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``fbgemm::block_bucketize_sparse_features`` with ``bucketize_pos=False,
    sequence=True`` as TorchRec's row-wise input_dist calls it.

    ``lengths`` is ``[F*B]`` key-major.  ``block = ceil(R_f / W)``;
    ``bucket = id // block``; ``local = id - bucket*block``.  Output is
    bucket-major: ``new_lengths[(w*F + f)*B + b]``; within one (w, f, b) bag the
    original order of ids is preserved (stable).  Also returns
    ``unbucketize_permute``: for every input position its position in the
    output values.
    """
    F = len(num_rows_per_feature)
    B = batch_size
    assert lengths.numel() == F * B
    offsets = lengths_to_offsets(lengths).to(torch.int64)
    new_lengths = torch.zeros(world_size * F * B, dtype=lengths.dtype)
    buckets_of: List[int] = []
    for f in range(F):
        block = -(-int(num_rows_per_feature[f]) // world_size)
        for b in range(B):
            for p in range(int(offsets[f * B + b]), int(offsets[f * B + b + 1])):
                w, _ = _bucket_of(int(values[p]), block, world_size)
                buckets_of.append(w)
                new_lengths[(w * F + f) * B + b] += 1
    new_offsets = lengths_to_offsets(new_lengths).to(torch.int64)
    cursor = new_offsets[:-1].clone()
    new_values = torch.empty_like(values)
    unbucketize = torch.empty(values.numel(), dtype=torch.int64)
    p = 0
    for f in range(F):
        block = -(-int(num_rows_per_feature[f]) // world_size)
        for b in range(B):
            for p in range(int(offsets[f * B + b]), int(offsets[f * B + b + 1])):
                w = buckets_of[p]
                slot = (w * F + f) * B + b
                dst = int(cursor[slot])
                cursor[slot] += 1
                new_values[dst] = _bucket_of(int(values[p]), block, world_size)[1]
                unbucketize[p] = dst
    return new_lengths, new_values, unbucketize


def block_bucketize_vectorized(
    lengths: torch.Tensor, values: torch.Tensor, num_rows_per_feature: Sequence[int],
    world_size: int, batch_size: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same contract as :func:`block_bucketize_sparse_features`, vectorised so
    tests can use it at BASELINE sizes.  Checked equal to the loop version in
    tests/test_oracle_kjt.py."""
    F = len(num_rows_per_feature)
    B = batch_size
    n = values.numel()
    bag = torch.repeat_interleave(torch.arange(F * B), lengths.to(torch.int64), output_size=n)
    f = bag // B
    b = bag - f * B
    rows = torch.tensor(list(num_rows_per_feature), dtype=torch.int64)
    block = (rows + world_size - 1) // world_size
    blk = block[f]
    w = values // blk
    local = values - w * blk
    for p in torch.nonzero((values < 0) | (values >= blk * world_size)).flatten().tolist():   # fbgemm's fallback
        w[p], local[p] = _bucket_of(int(values[p]), int(blk[p]), world_size)
    slot = (w * F + f) * B + b
    new_lengths = torch.bincount(slot, minlength=world_size * F * B).to(lengths.dtype)
    order = torch.argsort(slot, stable=True)
    new_values = local[order]
    unbucketize = torch.empty(n, dtype=torch.int64)
    unbucketize[order] = torch.arange(n)
    return new_lengths, new_values, unbucketize


def gathered_range_shard(values_per_rank: Sequence[torch.Tensor], lengths_per_rank: Sequence[torch.Tensor],
                         row_lo: Sequence[int], row_hi: Sequence[int], batch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """What one rank holds after TorchRec's sparse input dist (``torchrec.distributed.dist_data.KJTAllToAll`` after
    ``block_bucketize_sparse_features`` for row-wise tables; plain feature routing for table-wise ones), written as a
    filter over every rank's key-major KJT: the key-major KJT over the GLOBAL batch (bag ``f*(W*B) + r*B + b``) with,
    in source order, the ids of bag ``(r, f, b)`` that fall into ``[row_lo[f], row_hi[f])``, rebased to ``row_lo[f]``.
    For ``row_lo/hi`` = a block range this is bucket ``w`` of :func:`block_bucketize_sparse_features` on the
    concatenated batch (checked in tests/test_oracle_golden.py).  Restates include/tt_b200.h::tt_kjt_gathered_range."""
    W, F, B = len(values_per_rank), len(row_lo), batch_size
    out_vals: List[int] = []
    out_len = torch.zeros(F * W * B, dtype=torch.int32)
    offs = [lengths_to_offsets(l).tolist() for l in lengths_per_rank]
    for f in range(F):
        for r in range(W):
            for b in range(B):
                for p in range(offs[r][f * B + b], offs[r][f * B + b + 1]):
                    v = int(values_per_rank[r][p])
                    if row_lo[f] <= v < row_hi[f]:
                        out_vals.append(v - row_lo[f])
                        out_len[f * W * B + r * B + b] += 1
    return torch.tensor(out_vals, dtype=torch.int64), out_len


def dedup_rows(linear_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Unique (table,row) keys, ascending, with inverse map and counts
    (``torch.unique(sorted=True, return_inverse=True, return_counts=True)``)."""
    return torch.unique(linear_ids, sorted=True, return_inverse=True, return_counts=True)
