"""Model-parallel EmbeddingBagCollection: table-wise and row-wise sharding over the GPUs of one
box (what ``DistributedModelParallel`` builds from the plan, /root/reference/03_model_training.py:798-815).

Per step and per sharding group (TW tables, RW tables) the exchange is TorchRec's:

  input_dist   ids travel to the rank that stores their rows
               TW: feature f -> owner(table of f);   RW: id -> id // block (``block_bucketize``)
               = all-to-all of per-destination counts (one tiny tensor, the only host sync),
                 all-to-all of lengths, all-to-all of values, then ``permute_2D_sparse_data``
                 from (source rank, feature) order to key-major over the GLOBAL batch
  lookup       the local (unsharded) EmbeddingBagCollection over the global batch W*B
  output_dist  TW: pooled rows go back to the sample's rank      -> all-to-all
               RW: partial sums of every rank are added          -> reduce-scatter (mean divides after)
  backward     the transposes: all-to-all / all-gather of grad, then the local fused
               backward + row-wise optimizer (no dense gradient anywhere)

Every peer is one NVSwitch hop away at full bandwidth, so the collectives are flat NCCL
all-to-all / reduce-scatter on the kernels' own output buffers; there is no hierarchical
(table-row-wise) stage.  Dense towers stay replicated; their gradients are summed by ONE
all-reduce over the flat gradient buffer (``DenseGradSync``).

The device work (local lookup, bucketize) is injected (``local_ebc_factory``, ``bucketize_fn``)
so that the routing / split / permutation bookkeeping can be exercised on CPU with gloo in
tests; the product defaults are the CUDA kernels and fail loudly without them.
"""
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
from torch import distributed as dist
from torch import nn

from ..modules.embedding_configs import EmbeddingBagConfig, PoolingType
from ..modules.embedding_modules import _TAG_ATTRS, EmbeddingBagCollection
from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .planner import ParameterSharding, ShardingPlan


# TorchRec divides the gradient that flows back through the pooled-embedding all-to-all / reduce-scatter by the
# world size (torchrec.distributed.comm_ops, GRADIENT_DIVISION = True by default): every rank's loss is the mean over
# ITS batch, the data-parallel towers average their gradients over ranks, and the tables must see the gradient of
# the same global objective (1/W) * sum_r loss_r.  Here the division is folded into the fused embedding backward
# (tt_sparse_optimizer.grad_scale = 1/W on every local shard), so it costs nothing.
_GRADIENT_DIVISION = True


def set_gradient_division(val: bool) -> None:
    """``torchrec.distributed.comm_ops.set_gradient_division``; read when a module is sharded."""
    global _GRADIENT_DIVISION
    _GRADIENT_DIVISION = bool(val)


def get_gradient_division() -> bool:
    return _GRADIENT_DIVISION


# --------------------------------------------------------------------------- differentiable collectives
class _AllToAllRows(torch.autograd.Function):
    """``all_to_all_single`` over flattened row blocks; backward is the reverse exchange."""

    @staticmethod
    def forward(ctx, x, in_splits, out_splits, pg):
        ctx.pg, ctx.in_splits, ctx.out_splits = pg, in_splits, out_splits
        out = x.new_empty(sum(out_splits))
        dist.all_to_all_single(out, x.contiguous().view(-1), output_split_sizes=out_splits, input_split_sizes=in_splits, group=pg)
        return out

    @staticmethod
    def backward(ctx, g):
        out = g.new_empty(sum(ctx.in_splits))
        dist.all_to_all_single(out, g.contiguous().view(-1), output_split_sizes=ctx.in_splits, input_split_sizes=ctx.out_splits, group=ctx.pg)
        return out, None, None, None


class _ReduceScatterRows(torch.autograd.Function):
    """[W*B, D] partial sums -> [B, D] (sum over ranks); backward all-gathers the gradient."""

    @staticmethod
    def forward(ctx, x, pg):
        ctx.pg = pg
        W = dist.get_world_size(pg)
        out = x.new_empty(x.shape[0] // W, x.shape[1])
        dist.reduce_scatter_tensor(out, x.contiguous(), group=pg)
        return out

    @staticmethod
    def backward(ctx, g):
        W = dist.get_world_size(ctx.pg)
        out = g.new_empty(g.shape[0] * W, g.shape[1])
        dist.all_gather_into_tensor(out, g.contiguous(), group=ctx.pg)
        return out, None


class PeerExchange:
    """Symmetric (NVLink peer-mapped) buffers for the fused exchange: per rank TWO ``[B, D]`` fp32 matrices for pooled
    rows and two for their gradients (double buffered, alternating every step), each mapped into every process of
    the group.  The owner of a table writes its pooled rows straight into the buffer of the rank the sample belongs
    to (``tt_ebc_forward_peer``) and reads the gradient rows straight from it (``tt_ebc_backward_fused_peer``): the
    lookup kernel IS the all-to-all.  Allocation is collective.

    Why two buffers: with one, a step needs four cross-rank barriers (nobody still uses the old pooled rows | rows
    landed | gradients staged | nobody still reads the gradients) and two staging copies.  With two, "rows landed"
    of step i+1 already implies every rank is done with step i-1's buffer, and the same for the gradients, so a
    step has TWO barriers, the module's output IS the exchange buffer (no clone) and the towers' backward writes
    its input gradient straight into the gradient buffer (no copy)."""

    def __init__(self, rows: int, cols: int, device: torch.device, pg: Any) -> None:
        import torch.distributed._symmetric_memory as symm
        from .. import _native as N
        self.rows, self.cols, self.pg = rows, cols, pg
        self.world = dist.get_world_size(pg)
        if self.world > N.TT_MAX_PEERS:
            raise ValueError(f"peer exchange supports up to {N.TT_MAX_PEERS} ranks")
        group = pg if pg is not None else dist.group.WORLD
        self._pooled2 = symm.empty(2, rows, cols, dtype=torch.float32, device=device)
        self._grad2 = symm.empty(2, rows, cols, dtype=torch.float32, device=device)
        self._h_pooled = symm.rendezvous(self._pooled2, group)
        self._h_grad = symm.rendezvous(self._grad2, group)
        nbytes = rows * cols * 4
        self._pooled_peers = [self._peers(self._h_pooled, b * nbytes) for b in range(2)]
        self._pooled_peers_scatter = [self._peers(self._h_pooled, b * nbytes, N.TT_PEER_SCATTER_ADD) for b in range(2)]
        self._grad_peers = [self._peers(self._h_grad, b * nbytes) for b in range(2)]
        self.cur = 0                    # buffer of the step in flight (flipped at the start of every forward)
        # row-wise shards ADD into pre-zeroed buffers: both start zeroed; afterwards the backward of a step clears the
        # buffer the step used (ordered against the peers' next adds by the gradient barrier).
        # dirty[b]: buffer b holds sums that no backward has cleared (a forward without a backward: eval, or training
        # resumed after eval) -> the next forward that uses it clears it itself and takes one more barrier.  Tracked
        # PER BUFFER: after train -> eval -> train the step that follows the first training step lands on the buffer the
        # last-but-one eval forward left behind.  Every rank runs the same forward / backward sequence, so the flags (and
        # with them the number of barriers) agree across ranks.
        self._pooled2.zero_()
        self.dirty = [False, False]
        # forward_open: the last thing this exchange did was a forward (no backward since).  Eager steps alternate buffers,
        # so the next eager forward never touches the rows a peer may still be reading; a captured step replays on ONE
        # fixed buffer and must not start while that can be the case (see clean()).
        self.forward_open = False
        dist.barrier(group=pg)

    def _peers(self, handle, byte_offset: int = 0, flags: int = 0):
        from .. import _native as N
        pb = N.PeerBuffers()
        pb.world, pb.rows_per_peer, pb.flags = self.world, self.rows, flags
        for r, p in enumerate(handle.buffer_ptrs):
            pb.ptr[r] = p + byte_offset
        return pb

    def flip(self) -> int:
        self.cur ^= 1
        return self.cur

    def pooled(self, b: int) -> torch.Tensor:
        return self._pooled2[b]

    def grad(self, b: int) -> torch.Tensor:
        return self._grad2[b]

    def pooled_peers(self, b: int, scatter_add: bool):
        return (self._pooled_peers_scatter if scatter_add else self._pooled_peers)[b]

    def grad_peers(self, b: int):
        return self._grad_peers[b]

    def clean(self) -> None:
        """What ``CudaGraphTrainStep`` does before a replay (a captured step runs on ONE fixed buffer and cannot look at
        flags): clears every scatter-add buffer a backward-less forward left dirty, and -- if the last operation was an
        eager forward, whose rows a peer may still be reading from the very buffer the replay writes -- lines the ranks
        up.  Collective: one barrier when there was anything to do, none in the steady state of a training loop."""
        if not any(self.dirty) and not self.forward_open:
            return
        for b in (0, 1):
            if self.dirty[b]:
                self.pooled(b).zero_()
                self.dirty[b] = False
        self.forward_open = False
        self.barrier_pooled()

    def barrier_pooled(self) -> None:
        self._h_pooled.barrier(channel=0)

    def barrier_grad(self) -> None:
        self._h_grad.barrier(channel=0)


class _PeerTwLookup(torch.autograd.Function):
    """Lookup over the global batch whose stores are the output exchange and whose backward reads the gradient from
    the peers (see ``PeerExchange``).  ``ebc`` is None on a rank that owns no table of the group: it still takes
    part in the barriers and stages its gradient.  ``pre_backward`` (optional callable) runs right before the
    embedding backward is launched -- the data-parallel towers start their gradient all-reduce there, on a side
    stream, so that it overlaps the embedding update."""

    @staticmethod
    def forward(ctx, ex, ebc, layout, kjt_keys, values, offsets, scatter_add, pre_backward, *anchors):
        from ctypes import byref
        from .. import _native as N
        b = ex.flip()
        pooled = ex.pooled(b)
        if scatter_add and ex.dirty[b]:
            # no backward has cleared this buffer since a forward last added into it (eval forwards, or the training
            # steps right after them) -- clear it now and make sure every rank has done so before anybody adds
            pooled.zero_()
            ex.barrier_pooled()
        if ebc is not None:
            dev = values.device
            plan, _ = ebc._build_plan(kjt_keys, ex.world * ex.rows, with_state=False, out_layout=layout)
            N.call("tt_ebc_forward_peer", byref(plan), N.ptr(values), N.ptr(offsets), byref(ex.pooled_peers(b, scatter_add)),
                   N.stream_ptr(dev))
            ctx.save_for_backward(values, offsets)
        ex.barrier_pooled()           # every owner's rows have landed here
        if scatter_add:
            ex.dirty[b] = True
        ex.forward_open = True
        ctx.ex, ctx.ebc, ctx.layout, ctx.kjt_keys, ctx.n_anchors, ctx.buf = ex, ebc, layout, kjt_keys, len(anchors), b
        ctx.scatter_add = scatter_add
        ctx.pre_backward = pre_backward
        out = pooled.view(ex.rows, ex.cols)
        return out

    @staticmethod
    def backward(ctx, grad):
        from ctypes import byref
        from .. import _native as N
        ex, ebc, b = ctx.ex, ctx.ebc, ctx.buf
        ex.forward_open = False        # every reader of the pooled rows is upstream of this node; the gradient barrier follows
        gbuf = ex.grad(b)
        if grad.data_ptr() != gbuf.data_ptr():
            gbuf.copy_(grad)          # the producer did not write in place (see FusedTowersTC grad_dst)
        if ctx.scatter_add:
            # Row-wise shards ADD their rows, so the buffer must be zero before its next use.  This node runs after every
            # consumer of the pooled rows (it is the last one of the towers' backward), and the peers' next adds come after
            # the gradient barrier below, which this rank reaches after the clear.  Clearing HERE -- not "the other buffer in
            # the next forward" -- is what keeps a captured step correct: a CUDA graph replays with ONE fixed buffer index.
            ex.pooled(b).zero_()
            ex.dirty[b] = False
        if ctx.pre_backward is not None:
            ctx.pre_backward()
        ex.barrier_grad()             # every rank's gradient rows are staged
        grads = (None,) * ctx.n_anchors
        if ebc is not None:
            values, offsets = ctx.saved_tensors
            dev = values.device
            spec = ebc._sparse_optimizer_spec(advance_step=True)
            dense_grads = None
            if spec is None:
                dense_grads = ebc._alloc_dense_grads()
                spec = N.SparseOptimizer(kind=N.OPT_DENSE_GRAD)
                spec.grad_scale = float(getattr(ebc, "_grad_scale", 1.0))
            plan, _ = ebc._build_plan(ctx.kjt_keys, ex.world * ex.rows, with_state=True, dense_grads=dense_grads, out_layout=ctx.layout)
            n = values.numel()
            ws = N.workspace(N.load().tt_ebc_backward_workspace_bytes(n), dev)
            N.call("tt_ebc_backward_fused_peer", byref(plan), byref(spec), N.ptr(values), n, N.ptr(offsets),
                   byref(ex.grad_peers(b)), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
            if dense_grads is not None:
                grads = tuple(dense_grads)
        return (None,) * 8 + grads


def _native_bucketize(lengths, offsets, values, num_rows, F, B, W):
    from ..functional import block_bucketize
    return block_bucketize(lengths, offsets, values, num_rows, F, B, W)


def _native_gather_range(values, capacity, offsets, lo, hi, W, F, B):
    from ..functional import kjt_gathered_range
    return kjt_gathered_range(values, capacity, offsets, lo, hi, W, F, B)


def _default_local_ebc(tables: List[EmbeddingBagConfig], device: torch.device) -> nn.Module:
    return EmbeddingBagCollection(tables=tables, device=device)


class _Group:
    """One sharding group (all TW tables, or all RW tables) of a sharded collection."""

    def __init__(self, kind: str) -> None:
        self.kind = kind
        self.features: List[str] = []          # features of the group, in send order
        self.feat_dim: Dict[str, int] = {}
        self.feat_rows: Dict[str, int] = {}
        self.dest_features: List[List[str]] = []  # per destination rank
        self.local_ebc: Optional[nn.Module] = None
        self.local_features: List[str] = []    # features this rank looks up, local EBC order
        self.mean_features: List[str] = []     # RW features with mean pooling (divide after reduce-scatter)


class ShardedEmbeddingBagCollection(nn.Module):
    def __init__(self, ebc: EmbeddingBagCollection, plan: Dict[str, ParameterSharding], device: torch.device, pg: Any = None,
                 local_ebc_factory: Callable[[List[EmbeddingBagConfig], torch.device], nn.Module] = _default_local_ebc,
                 bucketize_fn: Callable = _native_bucketize, peer_exchange: bool = False,
                 gather_range_fn: Callable = _native_gather_range) -> None:
        """``peer_exchange=True`` replaces the table-wise output all-to-all (and its backward) by
        stores / loads over NVLink peer memory issued by the lookup kernels themselves."""
        super().__init__()
        self._peer_exchange = bool(peer_exchange)
        self._peer: Dict[str, PeerExchange] = {}
        self._whole: Optional[torch.Tensor] = None
        self._pg = pg
        self._rank = dist.get_rank(pg)
        self._world = dist.get_world_size(pg)
        self._device = torch.device(device)
        self._configs = ebc.embedding_bag_configs()
        self._plan = plan
        self._bucketize = bucketize_fn
        self._gather_range = gather_range_fn
        self._out_features = ebc.feature_names()
        self._out_dims = [c.embedding_dim for c in self._configs for _ in c.feature_names]
        W, r = self._world, self._rank
        tags = {c.name: {a: getattr(ebc.embedding_bags[c.name].weight, a) for a in _TAG_ATTRS
                         if hasattr(ebc.embedding_bags[c.name].weight, a)} for c in self._configs}

        self._tw, self._rw = _Group("table_wise"), _Group("row_wise")
        self._tw.dest_features = [[] for _ in range(W)]
        tw_local_cfgs: List[EmbeddingBagConfig] = []
        rw_local_cfgs: List[EmbeddingBagConfig] = []
        dp_local_cfgs: List[EmbeddingBagConfig] = []
        self._dp_features: List[str] = []                        # features of data_parallel tables, local EBC order
        self._dp_tags: Dict[str, Dict[str, Any]] = {}            # data_parallel table -> its in-backward optimizer tags
        self._dp_state: Dict[str, Dict[str, torch.Tensor]] = {}  # ... -> row-wise optimizer state of the replica
        self._dp_step = 0
        self._rw_block: Dict[str, int] = {}
        self._rw_feat_block: Dict[str, int] = {}
        self._shard_info: Dict[str, Tuple[str, int, int]] = {}   # table -> (kind, row offset, local rows)
        self._col_info: Dict[str, Tuple[int, int]] = {}          # column-wise table -> (first column, columns) held here
        self._real_feat: Dict[str, str] = {}                     # internal name of a column shard's feature -> KJT key
        self._cw_pieces: Dict[str, List[str]] = {}               # feature of a column-wise table -> its pieces, shard order
        for c in self._configs:
            ps = plan[c.name]
            if ps.sharding_type == "table_wise":
                owner = ps.ranks[0]
                for f in c.feature_names:
                    self._tw.dest_features[owner].append(f)
                    self._tw.feat_dim[f], self._tw.feat_rows[f] = c.embedding_dim, c.num_embeddings
                if owner == r:
                    tw_local_cfgs.append(EmbeddingBagConfig(name=c.name, embedding_dim=c.embedding_dim, num_embeddings=c.num_embeddings,
                                                            feature_names=list(c.feature_names), pooling=c.pooling,
                                                            weight_init_min=c.get_weight_init_min(), weight_init_max=c.get_weight_init_max()))
                    self._shard_info[c.name] = ("table_wise", 0, c.num_embeddings)
                else:
                    self._shard_info[c.name] = ("table_wise", 0, 0)
            elif ps.sharding_type == "column_wise":
                # Shard j = columns [j*dw, (j+1)*dw) of the table on ranks[j].  Each shard is handled as a table of the
                # table-wise group whose features carry an internal name ("<feature>@cw<j>"): ids travel to every shard
                # owner, every owner looks up the global batch in its column slice, the slices return with the table-wise
                # output exchange and are laid side by side.  Each shard keeps its OWN row-wise optimizer state (the
                # mean of g^2 runs over the shard's columns), as TorchRec's column-wise shards -- separate fused tables -- do.
                ranks = list(ps.ranks) if ps.ranks else list(range(W))
                if len(set(ranks)) != len(ranks) or c.embedding_dim % len(ranks):
                    raise ValueError(f"column_wise {c.name}: needs distinct ranks whose number divides embedding_dim "
                                     f"(ranks {ranks}, dim {c.embedding_dim})")
                if self._peer_exchange:
                    raise NotImplementedError("column_wise tables use the NCCL exchange; build the module with peer_exchange=False")
                dw = c.embedding_dim // len(ranks)
                self._shard_info[c.name] = ("column_wise", 0, 0)
                for j, owner in enumerate(ranks):
                    vfs = []
                    for f in c.feature_names:
                        vf = f"{f}@cw{j}"
                        vfs.append(vf)
                        self._real_feat[vf] = f
                        self._cw_pieces.setdefault(f, []).append(vf)
                        self._tw.dest_features[owner].append(vf)
                        self._tw.feat_dim[vf], self._tw.feat_rows[vf] = dw, c.num_embeddings
                    if owner == r:
                        tw_local_cfgs.append(EmbeddingBagConfig(name=c.name, embedding_dim=dw, num_embeddings=c.num_embeddings,
                                                                feature_names=vfs, pooling=c.pooling,
                                                                weight_init_min=c.get_weight_init_min(), weight_init_max=c.get_weight_init_max()))
                        self._shard_info[c.name] = ("column_wise", 0, c.num_embeddings)
                        self._col_info[c.name] = (j * dw, dw)
            elif ps.sharding_type == "row_wise":
                block = ps.block_size or -(-c.num_embeddings // W)
                self._rw_block[c.name] = block
                local_rows = max(0, min(block, c.num_embeddings - r * block))
                for f in c.feature_names:
                    self._rw.features.append(f)
                    self._rw_feat_block[f] = block
                    self._rw.feat_dim[f], self._rw.feat_rows[f] = c.embedding_dim, c.num_embeddings
                    if c.pooling == PoolingType.MEAN:
                        self._rw.mean_features.append(f)
                # local shard: SUM pooling (the mean divisor is applied after the reduce-scatter);
                # at least one row so that the table exists on every rank
                rw_local_cfgs.append(EmbeddingBagConfig(name=c.name, embedding_dim=c.embedding_dim, num_embeddings=max(local_rows, 1),
                                                        feature_names=list(c.feature_names), pooling=PoolingType.SUM,
                                                        weight_init_min=c.get_weight_init_min(), weight_init_max=c.get_weight_init_max()))
                self._shard_info[c.name] = ("row_wise", r * block, local_rows)
            elif ps.sharding_type == "data_parallel":
                # a full replica on every rank: looked up locally on the rank's own batch (no exchange); its dense
                # gradient is all-reduced and the table's optimizer applied to the replica in sync_data_parallel()
                dp_local_cfgs.append(EmbeddingBagConfig(name=c.name, embedding_dim=c.embedding_dim, num_embeddings=c.num_embeddings,
                                                        feature_names=list(c.feature_names), pooling=c.pooling,
                                                        weight_init_min=c.get_weight_init_min(), weight_init_max=c.get_weight_init_max()))
                self._dp_features.extend(c.feature_names)
                self._shard_info[c.name] = ("data_parallel", 0, c.num_embeddings)
            else:
                raise NotImplementedError(f"sharding type {ps.sharding_type}")
        self._tw.features = [f for d in self._tw.dest_features for f in d]
        self._tw.local_features = list(self._tw.dest_features[r])
        self._rw.dest_features = [list(self._rw.features) for _ in range(W)]
        self._rw.local_features = list(self._rw.features)
        # The local collections are deliberately NOT registered as sub-modules: this module's
        # state_dict / load_state_dict speak TorchRec's key names (one ShardedTensor per table).
        tw_ebc = local_ebc_factory(tw_local_cfgs, self._device) if tw_local_cfgs else None
        rw_ebc = local_ebc_factory(rw_local_cfgs, self._device) if rw_local_cfgs else None
        dp_ebc = local_ebc_factory(dp_local_cfgs, self._device) if dp_local_cfgs else None
        object.__setattr__(self, "tw_ebc", tw_ebc)
        object.__setattr__(self, "rw_ebc", rw_ebc)
        object.__setattr__(self, "dp_ebc", dp_ebc)
        self._tw.local_ebc, self._rw.local_ebc = tw_ebc, rw_ebc
        if dp_ebc is not None:
            # The replicas carry NO in-backward tags: their lookup's backward returns the dense [R, D] gradient
            # (OPT_DENSE_GRAD), which sync_data_parallel() all-reduces before it applies the table's optimizer itself.
            # Identical start on every rank, as DDP does for the towers.
            self._dp_tags = {c.name: tags.get(c.name, {}) for c in dp_local_cfgs}
            src = dist.get_global_rank(pg, 0) if pg is not None else 0
            for c in dp_local_cfgs:
                dist.broadcast(dp_ebc.embedding_bags[c.name].weight.data, src=src, group=pg)
        self._gradient_division = get_gradient_division()
        for local in (tw_ebc, rw_ebc):
            if local is not None:
                local._grad_scale = 1.0 / W if self._gradient_division else 1.0
        for local in (tw_ebc, rw_ebc):
            if local is None or not hasattr(local, "embedding_bags"):
                continue
            for name, bag in local.embedding_bags.items():
                for a, v in tags.get(name, {}).items():
                    setattr(bag.weight, a, v)
        self._prefetched: Dict[int, Any] = {}
        self._local_rows_dev: Dict[Any, torch.Tensor] = {}
        self._gather_cache: Dict[Any, Any] = {}       # (lo, hi) device arrays of the sync-free KJT gather, per key set
        self._gathered: Any = None                    # (id(kjt), values, offsets) all-gathered once per batch, both groups use it

    # ---- surface shared with EmbeddingBagCollection
    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        return self._configs

    def feature_names(self) -> List[str]:
        return list(self._out_features)

    def shard_info(self) -> Dict[str, Tuple[str, int, int]]:
        return dict(self._shard_info)

    def set_pre_backward(self, fn) -> None:
        """``fn()`` is called right before the embedding backward is launched (tower gradients are complete by
        then): DistributedModelParallel starts the dense all-reduce there so that it overlaps the update."""
        self._pre_backward = fn

    def local_parameters(self):
        for local in (self.tw_ebc, self.rw_ebc, self.dp_ebc):
            if local is not None:
                yield from local.parameters()

    # ---- data-parallel tables -------------------------------------------------------------------
    def _dp_sub_kjt(self, kjt: KeyedJaggedTensor) -> KeyedJaggedTensor:
        """The rank's own batch restricted to the features of the data_parallel tables (no exchange)."""
        keys, feats = list(kjt.keys()), self._dp_features
        if keys == feats:
            return kjt
        order = [keys.index(f) for f in feats]
        dense = getattr(kjt, "_id_columns", None)
        if dense is not None and dense[0].is_cuda:      # one id per bag: rebuild from the id columns, no host sync
            dev = dense[0].device
            sel = self._cached(("dp_order", tuple(order)), lambda: torch.tensor(order, dtype=torch.int64, device=dev))
            table_rows = {f: c.num_embeddings for c in self._configs for f in c.feature_names}
            rows = self._cached(("dp_rows",), lambda: torch.tensor([table_rows[f] for f in feats], dtype=torch.int64, device=dev))
            return KeyedJaggedTensor.from_id_columns(list(feats), dense[0].index_select(0, sel), rows)
        return kjt.permute(order)

    def sync_data_parallel(self) -> None:
        """After the backward of a training step (``DistributedModelParallel.sync_dense_grads`` calls it, next to the
        towers' all-reduce): the dense ``[R, D]`` gradient of every data_parallel table is averaged over the ranks -- the
        gradient of the same global objective the sharded tables see through the gradient division -- and the optimizer
        the table was tagged with (``apply_optimizer_in_backward``) is applied to the replica, identically on every
        rank.  Row-wise Adagrad / SGD in their dense forms (a row without gradient does not move); row-wise Adam advances
        the rows whose averaged gradient is non-zero.  Untagged tables keep the averaged ``.grad`` for the caller's
        optimizer.  Collective: every rank calls it once per training step."""
        if self.dp_ebc is None:
            return
        from ..optim.rowwise_adagrad import RowWiseAdagrad
        from ..optim.rowwise_adam import RowWiseAdam
        W = self._world
        advanced = False
        for name, tags in self._dp_tags.items():
            w = self.dp_ebc.embedding_bags[name].weight
            g = w.grad if w.grad is not None else torch.zeros_like(w)       # a rank without a gradient still joins
            if self._gradient_division and g.is_cuda:
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self._pg)
            else:
                dist.all_reduce(g, group=self._pg)
                if self._gradient_division:
                    g.div_(W)
            classes = tags.get("_optimizer_classes")
            if not classes:
                w.grad = g
                continue
            cls, kw = classes[0], tags.get("_optimizer_kwargs", [{}])[0]
            st = self._dp_state.setdefault(name, {})
            with torch.no_grad():
                if issubclass(cls, RowWiseAdagrad):
                    if "sum" not in st:
                        st["sum"] = torch.full((w.shape[0],), float(kw.get("initial_accumulator_value", 0.0)), dtype=torch.float32, device=w.device)
                    st["sum"] += g.pow(2).mean(dim=1)
                    w -= float(kw.get("lr", RowWiseAdagrad.DEFAULT_LR)) * g / (st["sum"].sqrt() + float(kw.get("eps", RowWiseAdagrad.DEFAULT_EPS))).unsqueeze(1)
                elif issubclass(cls, RowWiseAdam):
                    if "exp_avg" not in st:
                        st["exp_avg"] = torch.zeros_like(w)
                        st["exp_avg_sq"] = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
                    if not advanced:
                        self._dp_step += 1
                        advanced = True
                    b1, b2 = kw.get("betas", (0.9, 0.999))
                    t = self._dp_step
                    hit = (g != 0).any(dim=1)
                    m = torch.where(hit.unsqueeze(1), b1 * st["exp_avg"] + (1 - b1) * g, st["exp_avg"])
                    v = torch.where(hit, b2 * st["exp_avg_sq"] + (1 - b2) * g.pow(2).mean(dim=1), st["exp_avg_sq"])
                    st["exp_avg"], st["exp_avg_sq"] = m, v
                    upd = (m / (1 - b1 ** t)) / ((v / (1 - b2 ** t)).sqrt() + float(kw.get("eps", RowWiseAdam.DEFAULT_EPS))).unsqueeze(1)
                    w -= float(kw.get("lr", RowWiseAdam.DEFAULT_LR)) * torch.where(hit.unsqueeze(1), upd, torch.zeros_like(upd))
                elif issubclass(cls, torch.optim.SGD):
                    w -= float(kw.get("lr", 1e-3)) * g
                else:
                    raise NotImplementedError(f"optimizer {cls.__name__} on a data_parallel table; supported: RowWiseAdagrad, "
                                              "RowWiseAdam, torch.optim.SGD (plain)")
            w.grad = None

    # ---- input dist ----------------------------------------------------------------------------
    def _dist_group(self, grp: _Group, kjt: KeyedJaggedTensor):
        """Returns the key-major KJT of this rank's features over the global batch (or None)."""
        W, B, pg = self._world, kjt.stride(), self._pg
        keys = kjt.keys()
        if not grp.features:
            return None
        dense = getattr(kjt, "_id_columns", None)
        if dense is not None and grp.kind == "table_wise" and dense[0].is_cuda:
            return self._dist_dense_ids(grp, keys, dense[0], B)
        if dense is not None and grp.kind == "row_wise" and self._peer_exchange and dense[0].is_cuda:
            return self._dist_dense_ids_rw(grp, keys, dense[0], B)
        if self._peer_exchange and getattr(kjt, "_values_padded", False) and kjt.values().is_cuda:
            return self._dist_kjt_gather(grp, kjt, B)
        sub = kjt.permute([keys.index(self._real_feat.get(f, f)) for f in grp.features])
        lengths, values = sub.lengths(), sub.values()
        if grp.kind == "row_wise":
            F = len(grp.features)
            # the kernel derives block = ceil(rows / W): hand it block * W so that a plan's own block_size is honoured
            rows = torch.tensor([self._rw_feat_block[f] * W for f in grp.features], dtype=torch.int64)
            new_len, _new_off, new_val, _unb = self._bucketize(lengths, sub.offsets(), values, rows, F, B, W)
            lengths, values = new_len, new_val
            seg_per_dest = [F] * W
        else:
            seg_per_dest = [len(d) for d in grp.dest_features]
        # per-destination value counts (device) -> exchange -> host (the one sync of the input dist)
        seg_ends = [0]
        for n in seg_per_dest:
            seg_ends.append(seg_ends[-1] + n * B)
        lens64 = lengths.to(torch.int64)
        csum = torch.zeros(lens64.numel() + 1, dtype=torch.int64, device=lengths.device)
        torch.cumsum(lens64, 0, out=csum[1:])
        ends = csum[torch.tensor(seg_ends, device=lengths.device)]
        send_counts = (ends[1:] - ends[:-1]).contiguous()
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=pg)
        both = torch.stack([send_counts, recv_counts]).cpu()
        send_v, recv_v = both[0].tolist(), both[1].tolist()
        n_local = seg_per_dest[self._rank]
        send_l = [n * B for n in seg_per_dest]
        recv_l = [n_local * B] * W
        lengths_recv = lengths.new_empty(sum(recv_l))
        dist.all_to_all_single(lengths_recv, lengths.contiguous(), output_split_sizes=recv_l, input_split_sizes=send_l, group=pg)
        values_recv = values.new_empty(sum(recv_v))
        dist.all_to_all_single(values_recv, values[:sum(send_v)].contiguous(), output_split_sizes=recv_v, input_split_sizes=send_v, group=pg)
        if n_local == 0:
            return None
        # (source rank, feature) segments -> key-major over the global batch
        seg_keys = [f"{s}:{i}" for s in range(W) for i in range(n_local)]
        recv = KeyedJaggedTensor(keys=seg_keys, values=values_recv, lengths=lengths_recv, stride=B)
        recv._length_per_key = None
        perm = [s * n_local + i for i in range(n_local) for s in range(W)]
        if W > 1 or n_local > 1:
            recv = self._permute_no_sync(recv, perm, sum(recv_v))
        return KeyedJaggedTensor(keys=list(grp.local_features), values=recv.values(), lengths=recv.lengths(),
                                 offsets=recv._offsets, stride=W * B)

    def _dist_dense_ids(self, grp: _Group, keys: List[str], ids: torch.Tensor, B: int):
        """Table-wise input dist for batches that arrived as dense id columns (``KeyedJaggedTensor.from_id_columns``,
        the reference's one-id-per-bag format): the [F, B] id columns themselves are exchanged -- every split is
        known on the host, so there is ONE all-to-all, no count / length exchange and no host sync -- and the
        key-major KJT over the global batch is built by ``tt_kjt_from_columns`` where the ids land."""
        W, pg = self._world, self._pg
        n_local = len(grp.dest_features[self._rank])
        order = [keys.index(self._real_feat.get(f, f)) for f in grp.features]
        if order == list(range(ids.shape[0])):
            send = ids
        else:
            sel = self._local_rows_dev.get(("order", grp.kind, tuple(order)))    # cached: no H2D copy per step
            if sel is None:
                sel = torch.tensor(order, dtype=torch.int64, device=ids.device)
                self._local_rows_dev[("order", grp.kind, tuple(order))] = sel
            send = ids.index_select(0, sel)
        recv = ids.new_empty(W * n_local * B)
        dist.all_to_all_single(recv, send.contiguous().view(-1), output_split_sizes=[n_local * B] * W,
                               input_split_sizes=[len(d) * B for d in grp.dest_features], group=pg)
        if n_local == 0:
            return None
        cols = recv.view(W, n_local, B).permute(1, 0, 2).reshape(n_local, W * B)
        rows = self._local_rows_dev.get(grp.kind)
        if rows is None:
            rows = torch.tensor([grp.feat_rows[f] for f in grp.local_features], dtype=torch.int64, device=ids.device)
            self._local_rows_dev[grp.kind] = rows
        return KeyedJaggedTensor.from_id_columns(list(grp.local_features), cols, rows)

    def _cached(self, key, make):
        t = self._local_rows_dev.get(key)
        if t is None:
            t = make()
            self._local_rows_dev[key] = t
        return t

    def _dist_dense_ids_rw(self, grp: _Group, keys: List[str], ids: torch.Tensor, B: int):
        """Row-wise input dist for dense id columns with the peer-memory exchange: the [F, B] id columns are
        ALL-GATHERED (8 bytes per id, fixed size, no host sync) and every rank keeps the ids of its own row range
        (``tt_kjt_from_columns_range``); all other bags of the global batch are empty on this rank."""
        W, pg, dev = self._world, self._pg, ids.device
        F = len(grp.features)
        order = [keys.index(f) for f in grp.features]
        send = ids if order == list(range(ids.shape[0])) else ids.index_select(
            0, self._cached(("order", grp.kind, tuple(order)), lambda: torch.tensor(order, dtype=torch.int64, device=dev)))
        gathered = ids.new_empty(W * F * B)
        dist.all_gather_into_tensor(gathered, send.contiguous().view(-1), group=pg)
        cols = gathered.view(W, F, B).permute(1, 0, 2).reshape(F, W * B)
        rows = self._cached(("rows", grp.kind), lambda: torch.tensor([grp.feat_rows[f] for f in grp.features], dtype=torch.int64, device=dev))
        table_of = {f: c.name for c in self._configs for f in c.feature_names}
        lo = self._cached(("lo", grp.kind), lambda: torch.tensor([self._shard_info[table_of[f]][1] for f in grp.features], dtype=torch.int64, device=dev))
        hi = self._cached(("hi", grp.kind), lambda: torch.tensor([self._shard_info[table_of[f]][1] + self._shard_info[table_of[f]][2]
                                                                 for f in grp.features], dtype=torch.int64, device=dev))
        kjt = KeyedJaggedTensor.from_id_columns(list(grp.local_features), cols, rows, row_range=(lo, hi))
        kjt._peer_scatter = True
        return kjt

    def _dist_kjt_gather(self, grp: _Group, kjt: KeyedJaggedTensor, B: int):
        """Sync-free input dist for multi-hot KJTs whose ``values`` sit in a FIXED-CAPACITY buffer (the same capacity on
        every rank: ``CudaGraphTrainStep.step_kjt``, or any KJT flagged ``_values_padded``): the whole KJT (padded values +
        offsets) is ALL-GATHERED -- two fixed-size collectives, no count exchange, no host sync, so the step can be
        captured -- and ``tt_kjt_gathered_range`` keeps what this rank stores: its row range of every row-wise table,
        all rows of the table-wise tables it owns, nothing of the rest (those bags are empty here).  TorchRec reaches
        the same KJT through block_bucketize + three all-to-alls + permute (KJTAllToAll), with a host sync for the
        split sizes; the ids a rank receives that it does not keep cost 8 bytes each on NVLink."""
        W, pg, dev = self._world, self._pg, kjt.values().device
        keys = list(kjt.keys())
        F = len(keys)
        cap = kjt.values().numel()
        key = (grp.kind, tuple(keys), B, cap)
        st = self._gather_cache.get(key)
        if st is None:
            table_of = {f: c.name for c in self._configs for f in c.feature_names}
            lo, hi = [], []
            for f in keys:
                if f not in grp.feat_dim:
                    lo.append(0); hi.append(0)                       # not a feature of this sharding group
                elif grp.kind == "row_wise":
                    _k, off, rows = self._shard_info[table_of[f]]
                    lo.append(off); hi.append(off + rows)
                else:
                    owned = f in grp.dest_features[self._rank]
                    lo.append(0); hi.append(grp.feat_rows[f] if owned else 0)
            st = (torch.tensor(lo, dtype=torch.int64, device=dev), torch.tensor(hi, dtype=torch.int64, device=dev))
            self._gather_cache[key] = st
        lo, hi = st
        if self._gathered is not None and self._gathered[0] is kjt:
            _, g_vals, g_offs = self._gathered
        else:
            vals = kjt.values().contiguous()
            offs = kjt.offsets().to(torch.int32).contiguous()
            g_vals = vals.new_empty(W * cap)
            g_offs = offs.new_empty(W * (F * B + 1))
            dist.all_gather_into_tensor(g_vals, vals, group=pg)
            dist.all_gather_into_tensor(g_offs, offs, group=pg)
            self._gathered = (kjt, g_vals, g_offs)
        if grp.kind == "table_wise" and not grp.dest_features[self._rank]:
            return None
        out_v, out_l, out_o = self._gather_range(g_vals, cap, g_offs, lo, hi, W, F, B)
        out = KeyedJaggedTensor(keys=keys, values=out_v, lengths=out_l, offsets=out_o, stride=W * B)
        out._values_padded = True
        if grp.kind == "row_wise":
            out._peer_scatter = True
            out._multi_hot = True
        return out

    @staticmethod
    def _permute_no_sync(kjt: KeyedJaggedTensor, perm: List[int], total: int) -> KeyedJaggedTensor:
        # KJT.permute needs length_per_key only to size the output; a true permutation keeps the total.
        kjt._length_per_key = [0] * len(kjt.keys())
        kjt._length_per_key[0] = total
        if not kjt.values().is_cuda:
            kjt._length_per_key = None  # CPU path slices by real per-key lengths
        return kjt.permute(perm)

    def input_dist(self, kjt: KeyedJaggedTensor):
        try:
            return {"B": kjt.stride(), "lengths": kjt.lengths(), "keys": kjt.keys(),
                    "tw": self._dist_group(self._tw, kjt), "rw": self._dist_group(self._rw, kjt)}
        finally:
            self._gathered = None

    def prefetch_input_dist(self, batch, ready_event=None) -> None:
        """Called by TrainPipelineSparseDist one batch ahead: issues the KJT exchange early."""
        kjt = batch.sparse_features
        if ready_event is not None and kjt.values().is_cuda:
            torch.cuda.current_stream(kjt.values().device).wait_event(ready_event)
        self._prefetched[id(kjt)] = self.input_dist(kjt)

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        ctx = self._prefetched.pop(id(features), None) or self.input_dist(features)
        W, B, pg = self._world, ctx["B"], self._pg
        cols: Dict[str, torch.Tensor] = {}
        # table-wise: lookup over the global batch, rows go home by all-to-all
        d_by_rank = [sum(self._tw.feat_dim[f] for f in d) for d in self._tw.dest_features]
        if self._tw.features and self._peer_exchange:
            cols.update(self._group_forward_peer(self._tw, ctx["tw"], self.tw_ebc, B, False))
        elif self._tw.features:
            d_loc = d_by_rank[self._rank]
            if ctx["tw"] is not None:
                pooled = self.tw_ebc(ctx["tw"]).values()            # [W*B, d_loc]
                flat = pooled.reshape(-1)
            else:
                # no table of this group lives here; still take part in the exchange, and in its
                # backward (requires_grad keeps the reverse all-to-all in this rank's autograd graph)
                flat = torch.zeros(0, dtype=torch.float32, device=self._device, requires_grad=torch.is_grad_enabled())
            recv = _AllToAllRows.apply(flat, [B * d_loc] * W, [B * d for d in d_by_rank], pg)
            off = 0
            for r in range(W):
                if d_by_rank[r] == 0:
                    continue
                blk = recv[off:off + B * d_by_rank[r]].view(B, d_by_rank[r])
                off += B * d_by_rank[r]
                c0 = 0
                for f in self._tw.dest_features[r]:
                    cols[f] = blk[:, c0:c0 + self._tw.feat_dim[f]]
                    c0 += self._tw.feat_dim[f]
        # row-wise: partial pools over the global batch, summed by reduce-scatter
        if self._rw.features and getattr(ctx["rw"], "_peer_scatter", False):
            # dense id columns: one id per bag at most, mean pooling == sum pooling, nothing to divide
            rw_cols = self._group_forward_peer(self._rw, ctx["rw"], self.rw_ebc, B, True)
            if getattr(ctx["rw"], "_multi_hot", False) and self._rw.mean_features:
                # the shards ADD their partial sums into this rank's buffer; the mean divides by the bag's FULL length
                keys = ctx["keys"]
                for f in self._rw.mean_features:
                    k = keys.index(f)
                    ln = ctx["lengths"][k * B:(k + 1) * B].to(torch.float32).clamp(min=1.0)
                    rw_cols[f] = rw_cols[f] / ln.unsqueeze(1)
                self._whole = None            # the output is no longer the exchange buffer itself: concat below
            cols.update(rw_cols)
        elif self._rw.features:
            part = self.rw_ebc(ctx["rw"]).values()                  # [W*B, sum D_rw]
            pooled = _ReduceScatterRows.apply(part, pg)             # [B, sum D_rw]
            c0 = 0
            keys = ctx["keys"]
            for f in self._rw.features:
                d = self._rw.feat_dim[f]
                col = pooled[:, c0:c0 + d]
                if f in self._rw.mean_features:
                    k = keys.index(f)
                    ln = ctx["lengths"][k * B:(k + 1) * B].to(col.dtype).clamp(min=1.0)
                    col = col / ln.unsqueeze(1)
                cols[f] = col
                c0 += d
        # data-parallel: the replica is looked up on the rank's own batch, nothing travels
        if self._dp_features:
            kt = self.dp_ebc(self._dp_sub_kjt(features))
            for f in self._dp_features:
                cols[f] = kt[f]
            self._whole = None
        whole, self._whole = self._whole, None
        if whole is None:
            # a column-wise table's feature is the row of its shards' slices, in shard order
            parts = [cols[p] for f in self._out_features for p in self._cw_pieces.get(f, (f,))]
        values = whole if whole is not None else torch.cat(parts, dim=1)
        return KeyedTensor(keys=self._out_features, length_per_key=self._out_dims, values=values)

    def _group_forward_peer(self, grp: _Group, kjt, ebc, B: int, scatter_add: bool) -> Dict[str, torch.Tensor]:
        """Lookup of one sharding group whose output exchange (and its backward) is done by the lookup kernels
        themselves over NVLink peer memory: table-wise owners STORE rows into the sample's rank, row-wise shards
        ADD theirs (``scatter_add``)."""
        feats = [f for f in self._out_features if f in grp.feat_dim]
        col, layout_cols = 0, {}
        for f in feats:
            layout_cols[f] = col
            col += grp.feat_dim[f]
        ex = self._peer.get(grp.kind)
        if ex is None or ex.rows != B:
            ex = PeerExchange(B, col, self._device, self._pg)   # collective: B is the same on every rank
            self._peer[grp.kind] = ex
        if kjt is None:
            ebc, keys = None, ()
            values = offsets = None
            anchors = (torch.zeros(0, dtype=torch.float32, device=self._device, requires_grad=torch.is_grad_enabled()),)
        else:
            keys = tuple(kjt.keys())
            values, offsets = kjt.values().contiguous(), kjt.offsets().to(torch.int32).contiguous()
            if ebc._in_backward_kind() is not None and torch.is_grad_enabled():
                anchors = (torch.zeros(0, dtype=torch.float32, device=self._device, requires_grad=True),)
            else:
                anchors = tuple(ebc.embedding_bags[c.name].weight for c in ebc.embedding_bag_configs())
        out = _PeerTwLookup.apply(ex, ebc, (col, layout_cols), keys, values, offsets, scatter_add,
                                  getattr(self, "_pre_backward", None), *anchors)
        if feats == self._out_features:
            self._whole = out          # this group alone is the module's output, already in output order: no concat copy
            out._tt_grad_dst = ex.grad(ex.cur)   # a producer of d(out) may write it straight into the exchange buffer
        return {f: out[:, layout_cols[f]:layout_cols[f] + grp.feat_dim[f]] for f in feats}

    # ---- checkpoint surface: ShardedTensor entries named like the unsharded module --------------
    def _local_weight(self, name: str) -> Optional[torch.Tensor]:
        kind, _off, rows = self._shard_info[name]
        if rows == 0:
            return None
        local = self._local_ebc_of(name)
        return local.embedding_bags[name].weight.detach()[:rows]

    def include_optimizer_state(self, on: bool = True) -> "ShardedEmbeddingBagCollection":
        """``True``: ``state_dict()`` also carries the fused row-wise optimizer state of every table
        (``embedding_bags.<t>.{sum | exp_avg | exp_avg_sq}``, row-sharded like the weights) and the Adam step, and
        ``load_state_dict`` restores them; default ``False`` = the reference's weights-only format."""
        self._state_dict_with_optimizer = bool(on)
        return self

    def _local_ebc_of(self, name: str):
        return {"row_wise": self.rw_ebc, "data_parallel": self.dp_ebc}.get(self._shard_info[name][0], self.tw_ebc)

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        from .. import _native as N
        from .sharded_tensor import make_col_sharded, make_row_sharded
        destination = {} if destination is None else destination
        for c in self._configs:
            kind, off, _rows = self._shard_info[c.name]
            if kind == "column_wise":
                destination[f"{prefix}embedding_bags.{c.name}.weight"] = make_col_sharded(
                    self._local_weight(c.name), self._col_info.get(c.name, (0, 0))[0], (c.num_embeddings, c.embedding_dim), self._pg)
                continue
            if kind == "data_parallel":
                # replicated: a plain tensor, as TorchRec's data-parallel tables appear in a state dict
                # (utils/model_training.py:178-180 keeps rank 0's copy)
                destination[f"{prefix}embedding_bags.{c.name}.weight"] = self._local_weight(c.name)
                continue
            destination[f"{prefix}embedding_bags.{c.name}.weight"] = make_row_sharded(
                self._local_weight(c.name), off, (c.num_embeddings, c.embedding_dim), self._pg)
        if getattr(self, "_state_dict_with_optimizer", False):
            if any(k == "column_wise" for k, _o, _r in self._shard_info.values()):
                raise NotImplementedError("include_optimizer_state: every column shard keeps its own row-wise state, which has no "
                                          "place under the unsharded key names; checkpoint column-wise tables weights-only")
            step = self._dp_step
            for c in self._configs:
                _kind, off, rows = self._shard_info[c.name]
                if _kind == "data_parallel":
                    for k, v in self._dp_state.get(c.name, {}).items():
                        destination[f"{prefix}embedding_bags.{c.name}.{k}"] = v.detach()
                    continue
                local = self._local_ebc_of(c.name)
                kind = next((l._in_backward_kind() for l in (self.tw_ebc, self.rw_ebc) if l is not None), None)
                names = {N.OPT_ROWWISE_ADAGRAD: ("sum",), N.OPT_ROWWISE_ADAM: ("exp_avg", "exp_avg_sq")}.get(kind, ())
                for k in names:
                    st = None
                    if rows > 0:
                        cfg = next(x for x in local.embedding_bag_configs() if x.name == c.name)
                        st = local._state_for(cfg, local.embedding_bags[c.name].weight, kind)[k][:rows]
                    shape = (c.num_embeddings, c.embedding_dim) if k == "exp_avg" else (c.num_embeddings,)
                    destination[f"{prefix}embedding_bags.{c.name}.{k}"] = make_row_sharded(st, off, shape, self._pg)
            for l in (self.tw_ebc, self.rw_ebc):
                if l is not None and hasattr(l, "fused_step"):
                    step = max(step, l.fused_step())
            destination[f"{prefix}fused_optimizer_step"] = torch.tensor(float(step), dtype=torch.float32)
        return destination

    def _cols_from(self, src, c0: int, cols: int, name: str) -> torch.Tensor:
        """This rank's columns [c0, c0 + cols) of a column-wise table out of a checkpoint entry (full tensor, or the
        ShardedTensor this module itself writes)."""
        from .sharded_tensor import ShardedTensor
        if ShardedTensor is not None and isinstance(src, ShardedTensor):
            for sh in src.local_shards():
                if sh.metadata.shard_offsets[1] == c0 and sh.metadata.shard_sizes[1] == cols:
                    return sh.tensor
            raise RuntimeError(f"{name}: the ShardedTensor in the checkpoint is split differently from this module's plan "
                               f"(need columns [{c0}, {c0 + cols}) on this rank); gather it to a full tensor first and load that")
        return src[:, c0:c0 + cols]

    def _rows_from(self, src, off: int, rows: int, name: str) -> torch.Tensor:
        """This rank's rows [off, off + rows) out of a checkpoint entry: a full tensor, or the ShardedTensor this
        module itself writes (same row split -> the local shard is the answer, no communication)."""
        from .sharded_tensor import ShardedTensor
        if ShardedTensor is not None and isinstance(src, ShardedTensor):
            for sh in src.local_shards():
                if sh.metadata.shard_offsets[0] == off and sh.metadata.shard_sizes[0] == rows:
                    return sh.tensor
            raise RuntimeError(f"{name}: the ShardedTensor in the checkpoint is split differently from this module's plan "
                               f"(need rows [{off}, {off + rows}) on this rank); gather it to a full tensor first "
                               "(utils/model_training.py:161-182 does) and load that")
        return src[off:off + rows]

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """Accepts, under the TorchRec key names, FULL (unsharded) tensors -- every rank keeps its rows -- or the
        ShardedTensors of this module's own ``state_dict()`` (resume of a sharded job without a gather)."""
        for c in self._configs:
            key = f"{prefix}embedding_bags.{c.name}.weight"
            if key not in state_dict:
                if strict:
                    missing_keys.append(key)
                continue
            w = self._local_weight(c.name)
            kind, off, rows = self._shard_info[c.name]
            if kind == "column_wise":
                if w is not None:
                    with torch.no_grad():
                        w.copy_(self._cols_from(state_dict[key], *self._col_info[c.name], key).to(w.device))
                continue
            if w is not None:
                with torch.no_grad():
                    w.copy_(self._rows_from(state_dict[key], off, rows, key).to(w.device))
            for k in ("sum", "exp_avg", "exp_avg_sq"):
                skey = f"{prefix}embedding_bags.{c.name}.{k}"
                if skey in state_dict and kind == "data_parallel":
                    self._dp_state.setdefault(c.name, {})[k] = state_dict[skey].detach().to(device=w.device, dtype=torch.float32).clone()
                    continue
                if skey in state_dict and rows > 0:
                    local = self._local_ebc_of(c.name)
                    part = self._rows_from(state_dict[skey], off, rows, skey).detach().to(device=w.device, dtype=torch.float32)
                    if part.shape[0] != rows or (k == "exp_avg" and tuple(part.shape[1:]) != (c.embedding_dim,)):
                        raise RuntimeError(f"size mismatch for {skey}: this rank needs rows [{off}, {off + rows}) of a "
                                           f"{c.num_embeddings}-row table, the checkpoint entry gives {tuple(part.shape)}")
                    full_rows = local.embedding_bags[c.name].weight.shape[0]
                    if part.shape[0] != full_rows:      # a row-wise shard padded to >= 1 row
                        pad = torch.zeros((full_rows - part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
                        part = torch.cat([part, pad])
                    local._fused_state.setdefault(c.name, {})[k] = part.clone()
        skey = f"{prefix}fused_optimizer_step"
        if skey in state_dict:
            self._dp_step = int(round(float(state_dict[skey])))
            for l in (self.tw_ebc, self.rw_ebc):
                if l is not None and hasattr(l, "_fused_step"):
                    l._fused_step = int(round(float(state_dict[skey])))
                    l._fused_step_dev = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing: List[str] = []
        self._load_from_state_dict(state_dict, "", {}, strict, missing, [], [])
        if strict and missing:
            raise RuntimeError(f"missing keys {missing}")
        return torch.nn.modules.module._IncompatibleKeys(missing, [])


def shard_embedding_modules(module: nn.Module, plan: ShardingPlan, device: torch.device, pg: Any = None, **kw) -> nn.Module:
    """Replaces every EmbeddingBagCollection under ``module`` by its sharded form, in place."""
    targets = [(path, m) for path, m in module.named_modules() if isinstance(m, EmbeddingBagCollection)]
    for path, ebc in targets:
        table_plan = plan.get_plan_for_module(path)
        if table_plan is None:
            raise RuntimeError(f"no sharding plan for module '{path}'")
        sharded = ShardedEmbeddingBagCollection(ebc, table_plan, device, pg, **kw)
        if path == "":
            return sharded
        parent = module
        parts = path.split(".")
        for p in parts[:-1]:
            parent = getattr(parent, p)
        setattr(parent, parts[-1], sharded)
    return module


class DenseGradSync:
    """Data-parallel towers: gradients of every non-embedding parameter are averaged across ranks with
    one all-reduce.  When the dense optimizer is ``FlatAdam`` the gradients already live in one flat
    buffer; otherwise they are flattened into a staging buffer and copied back."""

    def __init__(self, module: nn.Module, pg: Any = None) -> None:
        self._pg = pg
        self._params = [p for n, p in module.named_parameters() if "embedding_bags" not in n]
        self._side: Optional[torch.cuda.Stream] = None
        self._pending = False
        # identical initial dense weights on every rank
        for p in self._params:
            dist.broadcast(p.data, src=0, group=pg)

    def _flat(self):
        grads = [p.grad for p in self._params if p.grad is not None]
        if not grads:
            return None, grads
        base = grads[0].untyped_storage()
        if all(g.untyped_storage().data_ptr() == base.data_ptr() for g in grads):   # FlatAdam: one contiguous buffer already
            return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(base, 0, (base.nbytes() // 4,)), grads
        return None, grads

    def _avg_op(self, t: torch.Tensor):
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide
        return dist.ReduceOp.AVG if t.is_cuda else dist.ReduceOp.SUM

    def start_async(self) -> None:
        """Issues the all-reduce of the (flat) tower gradients on a side stream; ``all_reduce()`` then only joins.
        Called from the sharded module's pre-backward hook, i.e. after the towers' backward and before the
        embedding backward, which it overlaps.  Falls back to doing nothing when the gradients are not one flat
        CUDA buffer (``all_reduce()`` then does the whole job)."""
        if self._pending:
            return                   # a module with several sharding groups calls the hook once per group
        flat, _ = self._flat()
        if flat is None or not flat.is_cuda:
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=flat.device)
        cur = torch.cuda.current_stream(flat.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self._pg)
        self._pending = True

    def all_reduce(self) -> None:
        W = dist.get_world_size(self._pg)
        if self._pending:
            self._pending = False
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)
            return
        flat, grads = self._flat()
        if not grads:
            return
        if flat is not None:
            op = self._avg_op(flat)
            dist.all_reduce(flat, op=op, group=self._pg)
            if op != dist.ReduceOp.AVG:
                flat.div_(W)
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        op = self._avg_op(flat)
        dist.all_reduce(flat, op=op, group=self._pg)
        if op != dist.ReduceOp.AVG:
            flat.div_(W)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
