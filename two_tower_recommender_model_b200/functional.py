"""Autograd bridges from torch tensors to the C ABI (include/tt_b200.h).

Each ``torch.autograd.Function`` here does nothing but shape checks, output
allocation and one or two library calls on torch's current stream.  All of them
refuse CPU tensors: there is no CPU path.
"""
from ctypes import byref
from typing import Optional, Tuple

import os

import torch

from . import _native as N


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    N.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 CUDA tensor whose rows are contiguous (may be a column window)."""
    N.require_cuda(t, name)
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError(f"{name} must be a 2-D float32 tensor")
    if t.shape[1] > 0 and t.stride(1) != 1:
        t = t.contiguous()
    return t


# --------------------------------------------------------------------------- EBC
class EbcLookup(torch.autograd.Function):
    """Pooled lookup; backward either applies the fused row-wise optimizer in
    place (weights get no ``.grad``, as with TorchRec's fused TBE) or, when no
    in-backward optimizer was registered, returns dense ``[R, D]`` gradients."""

    @staticmethod
    def forward(ctx, ebc, kjt_keys, values, offsets, batch, *weights):
        N.require_cuda(values, "KeyedJaggedTensor.values")
        dev = values.device
        plan, total_dim = ebc._build_plan(kjt_keys, batch, with_state=False)
        pooled = torch.empty(batch, total_dim, dtype=torch.float32, device=dev)
        N.call("tt_ebc_forward", byref(plan), N.ptr(values), N.ptr(offsets), N.ptr(pooled), N.stream_ptr(dev))
        ctx.ebc = ebc
        ctx.kjt_keys = kjt_keys
        ctx.batch = batch
        ctx.n_weights = len(weights)
        ctx.save_for_backward(values, offsets)
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        values, offsets = ctx.saved_tensors
        ebc = ctx.ebc
        dev = values.device
        grad_pooled = _f32c(grad_pooled, "grad_pooled")
        spec = ebc._sparse_optimizer_spec(advance_step=True)
        dense_grads = None
        if spec is None:
            dense_grads = ebc._alloc_dense_grads()
            spec = N.SparseOptimizer(kind=N.OPT_DENSE_GRAD)
        plan, _ = ebc._build_plan(ctx.kjt_keys, ctx.batch, with_state=True, dense_grads=dense_grads)
        n = values.numel()
        ws = N.workspace(N.load().tt_ebc_backward_workspace_bytes(n), dev)
        N.call("tt_ebc_backward_fused", byref(plan), byref(spec), N.ptr(values), n, N.ptr(offsets),
               N.ptr(grad_pooled), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        grads = tuple(dense_grads) if dense_grads is not None else (None,) * ctx.n_weights
        return (None, None, None, None, None) + grads


# --------------------------------------------------------------------------- tower layers
class LinearAct(torch.autograd.Function):
    """``y = relu?(x @ w.T + b)`` in fp32 on CUDA cores (exact-fp32 path)."""

    @staticmethod
    def forward(ctx, x, w, b, relu: bool):
        x = _rows(x, "x")
        w = _f32c(w, "weight")
        if b is not None:
            b = _f32c(b, "bias")
        M, K = x.shape
        Nn = w.shape[0]
        if w.shape[1] != K:
            raise ValueError(f"shape mismatch: x {tuple(x.shape)} vs weight {tuple(w.shape)}")
        y = torch.empty(M, Nn, dtype=torch.float32, device=x.device)
        N.call("tt_linear_forward_f32", N.ptr(x), x.stride(0) if M > 0 else K, N.ptr(w), N.ptr(b), N.ptr(y),
               M, Nn, K, 1 if relu else 0, N.stream_ptr(x.device))
        ctx.relu = relu
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _f32c(dy, "dy")
        M, K = x.shape
        Nn = w.shape[0]
        dev = x.device
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty(M, K, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty_like(w)
        db = torch.empty(Nn, dtype=torch.float32, device=dev) if ctx.has_bias else None
        ws = N.workspace(N.load().tt_linear_backward_workspace_bytes(M, Nn, K), dev)
        N.call("tt_linear_backward_f32", N.ptr(x), x.stride(0) if M > 0 else K, N.ptr(w), N.ptr(y), N.ptr(dy),
               N.ptr(dx), K, N.ptr(dw), N.ptr(db), M, Nn, K, 1 if ctx.relu else 0, N.ptr(ws), ws.numel(),
               N.stream_ptr(dev))
        return dx, dw, db, None


def linear_act(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
    return LinearAct.apply(x, w, b, relu)


# --------------------------------------------------------------------------- losses
class DotBceLoss(torch.autograd.Function):
    """utils/model_training.py:136-140 in one launch: logits, mean BCE-with-logits
    loss and (saved for backward) d loss / d q, d loss / d c."""

    @staticmethod
    def forward(ctx, q, c, labels):
        q = _f32c(q, "query_embedding")
        c = _f32c(c, "candidate_embedding")
        N.require_cuda(labels, "labels")
        if labels.dtype != torch.int32:
            labels = labels.to(torch.int32)
        labels = labels.contiguous()
        B, d = q.shape
        dev = q.device
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dq = torch.empty_like(q) if need_grad else None
        dc = torch.empty_like(c) if need_grad else None
        ws = N.workspace(N.load().tt_dot_bce_workspace_bytes(B), dev)
        N.call("tt_dot_bce", N.ptr(q), N.ptr(c), N.ptr(labels), B, d, N.ptr(logits), N.ptr(loss), N.ptr(dq),
               N.ptr(dc), 1.0, N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        if need_grad:
            ctx.save_for_backward(dq, dc)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, g_loss, _g_logits):
        dq, dc = ctx.saved_tensors
        return dq * g_loss, dc * g_loss, None


def dot_bce_loss(q: torch.Tensor, c: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return DotBceLoss.apply(q, c, labels)


class InBatchSoftmaxLoss(torch.autograd.Function):
    """Sampled softmax with in-batch negatives: ``CE(q c^T / T, arange(B))``; the
    ``[B, B]`` logits are never materialised (fp32 CUDA-core path)."""

    @staticmethod
    def forward(ctx, q, c, temperature: float):
        q = _f32c(q, "query_embedding")
        c = _f32c(c, "candidate_embedding")
        B, d = q.shape
        dev = q.device
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        diag = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = N.workspace(N.load().tt_inbatch_softmax_workspace_bytes(B), dev)
        N.call("tt_inbatch_softmax_forward_f32", N.ptr(q), N.ptr(c), B, d, 1.0 / temperature, N.ptr(lse),
               N.ptr(diag), N.ptr(loss), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        ctx.inv_t = 1.0 / temperature
        ctx.save_for_backward(q, c, lse)
        ctx.mark_non_differentiable(diag)
        return loss, diag

    @staticmethod
    def backward(ctx, g_loss, _g_diag):
        q, c, lse = ctx.saved_tensors
        B, d = q.shape
        dq = torch.empty_like(q)
        dc = torch.empty_like(c)
        N.call("tt_inbatch_softmax_backward_f32", N.ptr(q), N.ptr(c), N.ptr(lse), B, d, ctx.inv_t, 1.0,
               N.ptr(dq), N.ptr(dc), N.stream_ptr(q.device))
        return dq * g_loss, dc * g_loss, None


_deterministic_softmax_backward = os.environ.get("TT_SOFTMAX_BWD") == "split"


def set_deterministic_softmax_backward(on: bool) -> None:
    """``True``: the bf16 in-batch softmax backward always runs the two-pass kernels, which are bit-reproducible
    from run to run; ``False`` (default): d <= 64 uses the one-pass kernel (about 1.7x faster), whose partial
    sums meet in L2 in CTA arrival order, so the low bits of the gradients vary between runs."""
    global _deterministic_softmax_backward
    N.call("tt_set_softmax_backward_mode", 1 if on else 0)
    _deterministic_softmax_backward = bool(on)


_softmax_wide = os.environ.get("TT_SOFTMAX_WIDE") != "0"


def set_softmax_wide(on: bool) -> None:
    """``True`` (default): in-batch softmax with 64 < d <= 256 runs the TS-form kernels (the CTA's rows resident in
    TMEM, no transposed operand copies); ``False``: the SS-form kernels of round 1 (kept for A/B comparisons)."""
    global _softmax_wide
    N.call("tt_set_softmax_wide_mode", 1 if on else 0)
    _softmax_wide = bool(on)


class InBatchSoftmaxLossTC(torch.autograd.Function):
    """Same loss on the tensor cores (tcgen05): bf16 operands, fp32 accumulation in TMEM,
    softmax out of TMEM; the backward recomputes S tile by tile and feeds P back to the tensor
    core from TMEM.  Nothing of size [B, B] is ever written to HBM."""

    @staticmethod
    def forward(ctx, q, c, temperature: float, q_bf16=None, c_bf16=None):
        q = _f32c(q, "query_embedding")
        c = _f32c(c, "candidate_embedding")
        B, d = q.shape
        dev = q.device
        # d <= 64: the fused one-pass backward reads q / c row-major only (MN-major tcgen05 operands), and so do the
        # TS-form kernels for wider embeddings; the SS-form two-pass kernels want the transposed copies as well
        if (d > 64 and not _softmax_wide) or (d <= 64 and (d % 4 != 0 or _deterministic_softmax_backward)):
            qb, qbt = cast_bf16(q, both=True)
            cb, cbt = cast_bf16(c, both=True)
        else:
            # the fused towers already produced the bf16 copies (same rounding): no cast kernels
            qb = q_bf16 if (q_bf16 is not None and d <= 64) else cast_bf16(q)
            cb = c_bf16 if (c_bf16 is not None and d <= 64) else cast_bf16(c)
            qbt = cbt = None
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        diag = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = N.workspace(N.load().tt_inbatch_softmax_bf16_workspace_bytes(B), dev)
        N.call("tt_inbatch_softmax_forward_bf16", N.ptr(qb), qb.stride(0), N.ptr(cb), cb.stride(0), B, d,
               1.0 / temperature, N.ptr(lse), N.ptr(diag), N.ptr(loss), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        ctx.inv_t = 1.0 / temperature
        ctx.save_for_backward(q, c, qb, cb, qbt, cbt, lse)
        ctx.mark_non_differentiable(diag)
        return loss, diag

    @staticmethod
    def backward(ctx, g_loss, _g_diag):
        q, c, qb, cb, qbt, cbt, lse = ctx.saved_tensors
        B, d = q.shape
        dq = torch.empty_like(q)
        dc = torch.empty_like(c)
        # one-pass path (d <= 64): the incoming dLoss is multiplied in by the finalize kernel (device scalar, no sync)
        fused = qbt is None and d <= 64 and g_loss.numel() == 1 and g_loss.dtype == torch.float32 and g_loss.is_cuda
        gs = g_loss.contiguous() if fused else None
        N.call("tt_inbatch_softmax_backward_bf16", N.ptr(qb), qb.stride(0), N.ptr(cb), cb.stride(0),
               N.ptr(qbt), qbt.stride(0) if qbt is not None else 0, N.ptr(cbt), cbt.stride(0) if cbt is not None else 0,
               N.ptr(q), q.stride(0), N.ptr(c), c.stride(0),
               N.ptr(lse), B, d, ctx.inv_t, 1.0, 0, N.ptr(dq), d, N.ptr(dc), d, N.ptr(gs), N.stream_ptr(q.device))
        if fused:
            return dq, dc, None, None, None
        return dq * g_loss, dc * g_loss, None, None, None


class InBatchSoftmaxGlobalTC(torch.autograd.Function):
    """In-batch softmax with GLOBAL negatives over a process group: rank r scores its queries against the
    candidates of EVERY rank, ``loss_r = mean_i CE(q_i . [c_0; ...; c_{W-1}]^T / T, r*B + i)`` -- the same function
    of the global batch as the single-GPU loss (all-gather of the candidate embeddings, reduce-scatter
    of their gradient).  The [B, W*B] logits are processed as W square [B, B] blocks with the tcgen05 kernels
    (``tt_inbatch_softmax_forward_bf16`` gives each block's row log-sum-exp; the row's global one is the
    log-sum-exp of those; ``tt_inbatch_softmax_backward_bf16`` takes it back in), so per rank the flops are
    ``6 B (W B) d`` = 1/W of the single-GPU loss at the same global batch.  The positive-pair term is subtracted
    only in the block of the rank's own candidates (the other blocks pass zeros as the "other side")."""

    @staticmethod
    def forward(ctx, q, c, temperature: float, pg, q_bf16=None, c_bf16=None):
        from torch import distributed as dist
        q = _f32c(q, "query_embedding")
        c = _f32c(c, "candidate_embedding")
        B, d = q.shape
        dev = q.device
        W, r = dist.get_world_size(pg), dist.get_rank(pg)
        if d > 64 or d % 4 != 0:
            raise NotImplementedError("global in-batch negatives run on the one-pass tcgen05 kernels: d <= 64, d % 4 == 0")
        qb = q_bf16 if q_bf16 is not None else cast_bf16(q)
        cb = c_bf16 if c_bf16 is not None else cast_bf16(c)
        cb = cb if (cb.stride(0) == cb.shape[1] and cb.is_contiguous()) else cb.contiguous()
        cb_all = torch.empty(W, B, cb.shape[1], dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(cb_all.view(W * B, cb.shape[1]), cb, group=pg)
        lses = torch.empty(W, B, dtype=torch.float32, device=dev)
        diags = torch.empty(W, B, dtype=torch.float32, device=dev)
        scratch = torch.empty((), dtype=torch.float32, device=dev)
        ws = N.workspace(N.load().tt_inbatch_softmax_bf16_workspace_bytes(B), dev)
        for w in range(W):
            N.call("tt_inbatch_softmax_forward_bf16", N.ptr(qb), qb.stride(0), N.ptr(cb_all[w]), cb_all.stride(1), B, d,
                   1.0 / temperature, N.ptr(lses[w]), N.ptr(diags[w]), N.ptr(scratch), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        lse = torch.logsumexp(lses, dim=0)
        diag = diags[r].clone()
        loss = (lse - diag).mean()
        ctx.inv_t, ctx.pg, ctx.W, ctx.r = 1.0 / temperature, pg, W, r
        ctx.save_for_backward(q, c, qb, cb_all, lse)
        ctx.mark_non_differentiable(diag)
        return loss, diag

    @staticmethod
    def backward(ctx, g_loss, _g_diag):
        from torch import distributed as dist
        q, c, qb, cb_all, lse = ctx.saved_tensors
        B, d = q.shape
        dev = q.device
        W, r = ctx.W, ctx.r
        zeros = torch.zeros(B, d, dtype=torch.float32, device=dev)
        dq_w = torch.empty(W, B, d, dtype=torch.float32, device=dev)
        dc_all = torch.empty(W, B, d, dtype=torch.float32, device=dev)
        gs = g_loss.contiguous() if (g_loss.numel() == 1 and g_loss.dtype == torch.float32 and g_loss.is_cuda) else None
        for w in range(W):
            own = w == r
            N.call("tt_inbatch_softmax_backward_bf16", N.ptr(qb), qb.stride(0), N.ptr(cb_all[w]), cb_all.stride(1),
                   None, 0, None, 0, N.ptr(q if own else zeros), d, N.ptr(c if own else zeros), d,
                   N.ptr(lse), B, d, ctx.inv_t, 1.0, 0, N.ptr(dq_w[w]), d, N.ptr(dc_all[w]), d, N.ptr(gs), N.stream_ptr(dev))
        dq = dq_w.sum(dim=0)
        dc = torch.empty(B, d, dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(dc, dc_all.view(W * B, d), op=dist.ReduceOp.SUM, group=ctx.pg)
        if gs is None:
            dq, dc = dq * g_loss, dc * g_loss
        return dq, dc, None, None, None, None


def in_batch_softmax_loss(q: torch.Tensor, c: torch.Tensor, temperature: float = 1.0, precision: str = "fp32",
                          negatives: str = "local", pg=None):
    """``precision="fp32"``: CUDA-core path, exact fp32.  ``"bf16"``: tcgen05 tensor-core path.
    ``negatives="global"`` (bf16 only, inside a process group): every rank's candidates are negatives for every
    rank's queries (all-gather + reduce-scatter); ``"local"``: per-rank negatives, no collective."""
    if precision == "bf16":
        def bf16_of(t):   # bf16 copy attached by FusedTowersTC ([B, 64], zero padded), valid for the tensor it came with
            b = getattr(t, "_tt_bf16", None)
            ok = (b is not None and b.dtype == torch.bfloat16 and b.shape[0] == t.shape[0] and b.shape[1] >= t.shape[1]
                  and b.device == t.device and getattr(t, "_tt_bf16_version", None) == t._version)
            return b if ok else None
        if negatives == "global":
            from torch import distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(pg) > 1:
                return InBatchSoftmaxGlobalTC.apply(q, c, temperature, pg, bf16_of(q), bf16_of(c))
        return InBatchSoftmaxLossTC.apply(q, c, temperature, bf16_of(q), bf16_of(c))
    if negatives == "global":
        from torch import distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(pg) > 1:
            raise NotImplementedError("global in-batch negatives need precision='bf16' (tcgen05 path)")
    return InBatchSoftmaxLoss.apply(q, c, temperature)


# --------------------------------------------------------------------------- integer ops (no autograd)
def block_bucketize(lengths: torch.Tensor, offsets: torch.Tensor, values: torch.Tensor, num_rows: torch.Tensor,
                    num_features: int, batch: int, world: int):
    """fbgemm::block_bucketize_sparse_features; returns
    ``(new_lengths, new_offsets, new_values, unbucketize_permute)``."""
    N.require_cuda(values, "values")
    dev = values.device
    n = values.numel()
    n_out = world * num_features * batch
    new_len = torch.empty(n_out, dtype=torch.int32, device=dev)
    new_off = torch.empty(n_out + 1, dtype=torch.int32, device=dev)
    new_val = torch.empty(n, dtype=torch.int64, device=dev)
    unb = torch.empty(n, dtype=torch.int64, device=dev)
    rows = num_rows.to(device=dev, dtype=torch.int64).contiguous()
    ws = N.workspace(N.load().tt_kjt_bucketize_workspace_bytes(num_features, batch, world, n), dev)
    N.call("tt_kjt_block_bucketize", N.ptr(lengths.contiguous()), N.ptr(offsets.contiguous()),
           N.ptr(values.contiguous()), n, N.ptr(rows), num_features, batch, world, N.ptr(new_len), N.ptr(new_off),
           N.ptr(new_val), N.ptr(unb), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    return new_len, new_off, new_val, unb


def kjt_gathered_range(values: torch.Tensor, capacity: int, offsets: torch.Tensor, row_lo: torch.Tensor, row_hi: torch.Tensor,
                       world: int, num_features: int, batch: int):
    """Row-range shard of ``world`` gathered KJTs (``tt_kjt_gathered_range``): ``values`` [world * capacity] int64 and
    ``offsets`` [world * (F*B + 1)] int32 as all-gathered; returns ``(values [world*capacity], lengths [F*world*B],
    offsets [F*world*B + 1])`` of the key-major KJT over the global batch that holds the ids in ``[row_lo[f], row_hi[f])``,
    rebased to ``row_lo``.  No host sync: the live count stays in ``offsets[-1]``."""
    N.require_cuda(values, "values")
    dev = values.device
    n = world * num_features * batch
    out_v = torch.empty(world * capacity, dtype=torch.int64, device=dev)
    out_l = torch.empty(n, dtype=torch.int32, device=dev)
    out_o = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ws = N.workspace(N.load().tt_kjt_gathered_range_workspace_bytes(world, num_features, batch), dev)
    N.call("tt_kjt_gathered_range", N.ptr(values), capacity, N.ptr(offsets), N.ptr(row_lo), N.ptr(row_hi), world, num_features, batch,
           N.ptr(out_v), N.ptr(out_l), N.ptr(out_o), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    return out_v, out_l, out_o


def dedup_rows(ebc, kjt):
    """Unique linearised (table,row) keys of a batch, ascending, with counts and the inverse map --
    ``torch.unique(sorted=True, return_inverse=True, return_counts=True)`` over ``row_base[table] + id`` -- from
    the key construction and radix sort of the fused backward (``tt_ebc_dedup``).  Returns
    ``(unique_keys int64 [U], inverse int64 [nnz], counts int64 [U])``; ids outside their table map to -1."""
    values = kjt.values()
    N.require_cuda(values, "KeyedJaggedTensor.values")
    dev = values.device
    values = values.to(torch.int64).contiguous()
    offsets = kjt.offsets().to(torch.int32).contiguous()
    n = values.numel()
    plan, _ = ebc._build_plan(tuple(kjt.keys()), kjt.stride(), with_state=False)
    uk = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    inv = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = N.workspace(N.load().tt_ebc_dedup_workspace_bytes(n), dev)
    N.call("tt_ebc_dedup", byref(plan), N.ptr(values), n, N.ptr(offsets), N.ptr(uk), N.ptr(cnt), N.ptr(inv), N.ptr(nu),
           N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    u = int(nu.item())
    return uk[:u], inv[:n].to(torch.int64), cnt[:u].to(torch.int64)


def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, key_bits: int = 32):
    """Stable radix sort of (uint32 key, uint32 payload) pairs held in int32 tensors."""
    N.require_cuda(keys, "keys")
    dev = keys.device
    n = keys.numel()
    ko = torch.empty_like(keys)
    vo = torch.empty_like(vals)
    ws = N.workspace(N.load().tt_sort_pairs_workspace_bytes(n), dev)
    N.call("tt_sort_pairs_u32", N.ptr(keys), N.ptr(vals), N.ptr(ko), N.ptr(vo), n, key_bits, N.ptr(ws), ws.numel(),
           N.stream_ptr(dev))
    return ko, vo


def score_topk(queries: torch.Tensor, items: torch.Tensor, k: int, item_index_base: int = 0, precision: str = "fp32",
               items_bf16: Optional[torch.Tensor] = None):
    """Exact dot-product top-k (descending score, ties -> lower index).  ``precision="bf16"`` scores
    on the tensor cores (tcgen05; d <= 64): operands are rounded to bf16, accumulation is fp32, and
    the top-k is fused into the TMEM epilogue.  ``items_bf16`` lets a caller reuse a resident bf16
    corpus instead of casting it per call."""
    queries = _f32c(queries, "queries")
    Q, d = queries.shape
    dev = queries.device
    Nn = items.shape[0] if items is not None else items_bf16.shape[0]
    scores = torch.empty(Q, k, dtype=torch.float32, device=dev)
    idx = torch.empty(Q, k, dtype=torch.int64, device=dev)
    if precision == "bf16":
        qb = cast_bf16(queries)
        ib = items_bf16 if items_bf16 is not None else cast_bf16(_f32c(items, "items"))
        ws = N.workspace(N.load().tt_topk_bf16_workspace_bytes(Q, Nn, k), dev)
        N.call("tt_score_topk_bf16", N.ptr(qb), qb.stride(0), N.ptr(ib), ib.stride(0), Q, Nn, d, k, item_index_base,
               N.ptr(scores), N.ptr(idx), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        return scores, idx
    items = _f32c(items, "items")
    ws = N.workspace(N.load().tt_topk_workspace_bytes(Q, Nn, k), dev)
    N.call("tt_score_topk_f32", N.ptr(queries), N.ptr(items), Q, Nn, d, k, item_index_base, N.ptr(scores),
           N.ptr(idx), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    return scores, idx


# --------------------------------------------------------------------------- tensor-core (bf16) primitives
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def cast_bf16(x: torch.Tensor, transposed: bool = False, both: bool = False, gate: Optional[torch.Tensor] = None):
    """fp32 [rows, cols] (row-contiguous, may be a column window) -> bf16 copy; ``transposed`` returns
    the [cols, rows] copy instead, ``both`` returns (row-major, transposed).  ``gate`` (fp32, same
    shape) zeroes the elements whose gate is <= 0 (ReLU backward).  Row pitches are padded to a
    multiple of 8 elements (TMA needs 16-byte pitches)."""
    x = _rows(x, "x")
    rows, cols = x.shape
    dev = x.device
    out = out_t = None
    if both or not transposed:
        out = torch.empty(rows, _pad8(cols), dtype=torch.bfloat16, device=dev)[:, :cols]
    if both or transposed:
        out_t = torch.empty(cols, _pad8(rows), dtype=torch.bfloat16, device=dev)[:, :rows]
    if gate is not None:
        gate = _rows(gate, "gate")
    N.call("tt_cast_f32_to_bf16", N.ptr(x), x.stride(0) if rows > 0 else cols, N.ptr(gate),
           gate.stride(0) if gate is not None else 0, rows, cols,
           N.ptr(out), out.stride(0) if out is not None else 0, N.ptr(out_t), out_t.stride(0) if out_t is not None else 0,
           N.stream_ptr(dev))
    if both:
        return out, out_t
    return out_t if transposed else out


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
              mask: Optional[torch.Tensor] = None, mask_bf16: Optional[torch.Tensor] = None, out_f32: bool = True,
              out_bf16: bool = False, out_bf16_t: bool = False):
    """``epilogue(a @ b.T)`` on tcgen05: a [M,K] bf16, b [N,K] bf16 (row pitch multiple of 8).
    Returns a dict with the requested outputs ("f32", "bf16", "bf16_t")."""
    N.require_cuda(a, "a")
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    Nn = b.shape[0]
    dev = a.device
    res = {}
    f32 = torch.empty(M, Nn, dtype=torch.float32, device=dev) if out_f32 else None
    b16 = torch.empty(M, _pad8(Nn), dtype=torch.bfloat16, device=dev)[:, :Nn] if out_bf16 else None
    b16t = torch.empty(Nn, _pad8(M), dtype=torch.bfloat16, device=dev)[:, :M] if out_bf16_t else None
    N.call("tt_gemm_bf16", N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), M, Nn, K, N.ptr(bias), 1 if relu else 0,
           N.ptr(mask), mask.stride(0) if mask is not None else 0,
           N.ptr(mask_bf16), mask_bf16.stride(0) if mask_bf16 is not None else 0, N.ptr(f32), Nn,
           N.ptr(b16), b16.stride(0) if b16 is not None else 0, N.ptr(b16t), b16t.stride(0) if b16t is not None else 0,
           N.stream_ptr(dev))
    if f32 is not None:
        res["f32"] = f32
    if b16 is not None:
        res["bf16"] = b16
    if b16t is not None:
        res["bf16_t"] = b16t
    return res


def gemm_bf16_splitk(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32 ``a @ b.T`` for small [M,N] and a long K (weight gradients): K split over the SMs."""
    M, K = a.shape
    Nn = b.shape[0]
    dev = a.device
    out = torch.empty(M, Nn, dtype=torch.float32, device=dev)
    ws = N.workspace(N.load().tt_gemm_bf16_splitk_workspace_bytes(M, Nn, K), dev)
    N.call("tt_gemm_bf16_splitk", N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), M, Nn, K, N.ptr(out), N.ptr(ws), ws.numel(),
           N.stream_ptr(dev))
    return out


def gemm_bf16_splitk_mn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32 ``a.T @ b`` for a [K, M], b [K, N] row-major bf16 and a long K (``dW = dZ^T A_prev`` from the row-major
    activations: the tensor core reads both operands MN-major, no transposed copies exist)."""
    K, M = a.shape
    Nn = b.shape[1]
    dev = a.device
    out = torch.empty(M, Nn, dtype=torch.float32, device=dev)
    ws = N.workspace(N.load().tt_gemm_bf16_splitk_workspace_bytes(M, Nn, K), dev)
    N.call("tt_gemm_bf16_splitk_mn", N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), M, Nn, K, N.ptr(out), N.ptr(ws), ws.numel(),
           N.stream_ptr(dev))
    return out


def colsum_bf16(x: torch.Tensor) -> torch.Tensor:
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    ws = N.workspace(N.load().tt_colsum_bf16_workspace_bytes(rows, cols), x.device)
    N.call("tt_colsum_bf16", N.ptr(x), x.stride(0), rows, cols, N.ptr(out), N.ptr(ws), ws.numel(), N.stream_ptr(x.device))
    return out


def fused_towers_supported(in_dims, hidden, out_dim, n_layers) -> bool:
    """Shapes ``tt_towers_forward_fused`` takes: two-layer towers, in <= 64, hidden <= 128, out <= 64, multiples of 8."""
    return (n_layers == 2 and len(in_dims) <= N.TT_MAX_TOWERS and len(set(in_dims)) == 1 and 8 <= in_dims[0] <= 64
            and 8 <= hidden <= 128 and 8 <= out_dim <= 64 and (in_dims[0] | hidden | out_dim) % 8 == 0)


class FusedTowersTC(torch.autograd.Function):
    """Both two-layer ReLU towers in ONE launch per direction (``tt_towers_forward_fused`` /
    ``tt_towers_backward_fused``): ``pooled`` is the [B, sum D] matrix of pooled embeddings, tower t reads the
    column window ``[cols[t], cols[t] + in_dim)`` of it and gets parameters ``params[4t : 4t+4]`` =
    (W1, b1, W2, b2).  Returns one fp32 [B, out] embedding per tower.  bf16 operands, fp32 accumulation,
    fp32 master weights and gradients, same numerics as ``MlpTC``."""

    @staticmethod
    def forward(ctx, pooled, cols, in_dim, grad_dst, *params):
        """``grad_dst``: optional fp32 ``[B, width]`` buffer that receives d(pooled) (the sharded module passes its
        NVLink exchange buffer, so the gradient needs no staging copy); only used when the towers' windows tile
        the whole matrix."""
        ctx.grad_dst = grad_dst
        pooled = _rows(pooled, "pooled embeddings")
        T = len(cols)
        B = pooled.shape[0]
        dev = pooled.device
        hidden, out_dim = params[0].shape[0], params[2].shape[0]
        ys = [torch.empty(B, out_dim, dtype=torch.float32, device=dev) for _ in range(T)]
        xb = torch.empty(T, B, 64, dtype=torch.bfloat16, device=dev)
        hb = torch.empty(T, B, 128, dtype=torch.bfloat16, device=dev)
        yb = torch.empty(T, B, 64, dtype=torch.bfloat16, device=dev)
        wbs = []
        arr = (N.TowerForward * T)()
        for t in range(T):
            w1, b1, w2, b2 = params[4 * t: 4 * t + 4]
            w1b, w2b = cast_bf16(_f32c(w1, "weight")), cast_bf16(_f32c(w2, "weight"))
            wbs += [w1b, w2b]
            a = arr[t]
            a.x, a.ldx = pooled.data_ptr() + 4 * cols[t], pooled.stride(0)
            a.w1_bf16, a.ldw1, a.b1 = w1b.data_ptr(), w1b.stride(0), (0 if b1 is None else _f32c(b1, "bias").data_ptr())
            a.w2_bf16, a.ldw2, a.b2 = w2b.data_ptr(), w2b.stride(0), (0 if b2 is None else _f32c(b2, "bias").data_ptr())
            a.xb, a.hb, a.yb = xb[t].data_ptr(), hb[t].data_ptr(), yb[t].data_ptr()
            a.y, a.ldy = ys[t].data_ptr(), out_dim
        N.call("tt_towers_forward_fused", arr, T, B, in_dim, hidden, out_dim, N.stream_ptr(dev))
        ctx.cols, ctx.in_dim, ctx.shape, ctx.width = cols, in_dim, (B, hidden, out_dim), pooled.shape[1]
        ctx.saved = (xb, hb, yb, wbs)
        ctx.has_bias = [p is not None for p in params]
        ctx.save_for_backward(*[p for p in params if p is not None])
        ctx.mark_non_differentiable(yb)
        ctx.set_materialize_grads(False)      # no 17 MB zero "gradient" for the bf16 copy; unused towers arrive as None
        return (*ys, yb)

    @staticmethod
    def backward(ctx, *dys):
        dys = dys[:-1]                      # the last output is the (non-differentiable) bf16 copy
        xb, hb, yb, wbs = ctx.saved
        B, hidden, out_dim = ctx.shape
        T, in_dim = len(ctx.cols), ctx.in_dim
        dev = xb.device
        need_dx = ctx.needs_input_grad[0]
        covered = sorted(ctx.cols) == list(range(0, ctx.width, in_dim))     # the windows tile the pooled matrix
        d_pooled = None
        if need_dx:
            gd = ctx.grad_dst
            if (covered and gd is not None and gd.shape == (B, ctx.width) and gd.dtype == torch.float32 and gd.is_contiguous()
                    and gd.device == dev):
                d_pooled = gd
            else:
                d_pooled = (torch.empty if covered else torch.zeros)(B, ctx.width, dtype=torch.float32, device=dev)
        grads = []
        arr = (N.TowerBackward * T)()
        keep = []
        for t in range(T):
            dy = _f32c(dys[t], "grad of tower output") if dys[t] is not None else torch.zeros(B, out_dim, dtype=torch.float32, device=dev)
            keep.append(dy)
            gw1 = torch.empty(hidden, in_dim, dtype=torch.float32, device=dev)
            gw2 = torch.empty(out_dim, hidden, dtype=torch.float32, device=dev)
            gb1 = torch.empty(hidden, dtype=torch.float32, device=dev) if ctx.has_bias[4 * t + 1] else None
            gb2 = torch.empty(out_dim, dtype=torch.float32, device=dev) if ctx.has_bias[4 * t + 3] else None
            grads += [gw1, gb1, gw2, gb2]
            a = arr[t]
            a.dy, a.lddy = dy.data_ptr(), dy.stride(0)
            a.w1_bf16, a.ldw1 = wbs[2 * t].data_ptr(), wbs[2 * t].stride(0)
            a.w2_bf16, a.ldw2 = wbs[2 * t + 1].data_ptr(), wbs[2 * t + 1].stride(0)
            a.xb, a.hb, a.yb = xb[t].data_ptr(), hb[t].data_ptr(), yb[t].data_ptr()
            a.dx, a.lddx = (d_pooled.data_ptr() + 4 * ctx.cols[t], d_pooled.stride(0)) if need_dx else (0, 0)
            a.dw1, a.dw2 = gw1.data_ptr(), gw2.data_ptr()
            a.db1, a.db2 = (0 if gb1 is None else gb1.data_ptr()), (0 if gb2 is None else gb2.data_ptr())
        ws = N.workspace(N.load().tt_towers_backward_workspace_bytes(B), dev)
        N.call("tt_towers_backward_fused", arr, T, B, in_dim, hidden, out_dim, N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        ctx.saved = None
        return (d_pooled, None, None, None, *grads)


class MlpTC(torch.autograd.Function):
    """A whole ReLU tower on the tensor cores: ``x -> relu(x W1^T + b1) -> ... -> relu(. WL^T + bL)``.
    Operands are bf16 (activations are produced in bf16 by the GEMM epilogues, row-major only: the
    weight-gradient GEMMs read them through MN-major descriptors), accumulation is fp32 in TMEM, master weights,
    the returned output and all gradients are fp32.  Backward: dZ_l = dA_l * (A_l > 0) is fused into
    the epilogue of the previous data-gradient GEMM; dW_l = dZ_l^T A_{l-1} is a split-K GEMM."""

    @staticmethod
    def forward(ctx, x, *params):
        L = len(params) // 2
        x = _rows(x, "x")
        acts = [cast_bf16(x)]          # row-major bf16 activations only: the weight-gradient GEMM reads them MN-major
        out = None
        for l in range(L):
            w, b = params[2 * l], params[2 * l + 1]
            wb = cast_bf16(_f32c(w, "weight"))
            last = l == L - 1
            r = gemm_bf16(acts[-1], wb, bias=None if b is None else _f32c(b, "bias"), relu=True,
                          out_f32=last, out_bf16=True)
            acts.append(r["bf16"])
            if last:
                out = r["f32"]
        ctx.L = L
        ctx.acts = acts
        ctx.save_for_backward(out, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, *params = ctx.saved_tensors
        L, acts = ctx.L, ctx.acts
        dz = cast_bf16(_f32c(dout, "dout"), gate=out)
        grads = [None] * (2 * L)
        dx = None
        for l in range(L - 1, -1, -1):
            w, b = params[2 * l], params[2 * l + 1]
            grads[2 * l] = gemm_bf16_splitk_mn(dz, acts[l])              # dZ_l^T A_{l-1}: [N_l, K_l]
            if b is not None:
                grads[2 * l + 1] = colsum_bf16(dz)
            if l > 0:
                wt = cast_bf16(w, transposed=True)                       # [K_l, N_l]
                dz = gemm_bf16(dz, wt, mask_bf16=acts[l], out_f32=False, out_bf16=True)["bf16"]
            elif ctx.needs_input_grad[0]:
                wt = cast_bf16(w, transposed=True)
                dx = gemm_bf16(dz, wt, out_f32=True)["f32"]
        ctx.acts = None
        return (dx, *grads)
