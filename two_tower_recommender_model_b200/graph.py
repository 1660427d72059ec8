"""Whole-step CUDA graph for the train loop (one GPU, or one rank of a DistributedModelParallel job: the
NCCL / peer-memory exchanges of the sharded embedding path are captured with the rest -- table-wise sharding
with id-column batches has no host sync in its input dist, so every rank replays the same fixed sequence).

The step (device batch construction from raw id columns, EBC lookup, towers, loss, backward with the
fused row-wise update, dense optimizer) is a fixed sequence of ~50 kernel launches whose arguments do
not change from step to step once the inputs live in static buffers.  ``CudaGraphTrainStep`` runs the
first calls eagerly (warm-up with REAL batches, so nothing is trained on dummy data), then captures
one step and replays it: per step the host issues two async copies and one graph launch.

Requirements: static shapes -- batches given as raw id columns ``[F, B]`` + labels ``[B]``, or (``kjt_capacity=``,
one GPU) as multi-hot KJT pieces: ``values`` of any length up to the capacity and ``lengths [F*B]``; the values live in a
fixed-capacity buffer, the offsets are scanned on the device inside the graph, and the kernels take the live count from
``offsets[-1]`` (the lookup never looks past it, the fused backward parks the unused tail on its sentinel key), so a
batch's number of ids may change from step to step without a re-capture; sparse optimizer
RowWiseAdagrad, RowWiseAdam or SGD (row-wise Adam keeps its step counter on the device: ``tt_sparse_optimizer.step_dev``
is incremented by the fused backward itself, so the replayed launch arguments never go stale); dense optimizer
``FlatAdam`` (device-side step counter) or SGD -- ``torch.optim.Adam`` computes its bias correction on the host
and is refused.
"""
from typing import List, Optional, Sequence

import torch

from .datasets.utils import Batch
from .sparse.jagged_tensor import KeyedJaggedTensor


class CudaGraphTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, keys: Sequence[str],
                 num_embeddings: Sequence[int], batch_size: int, device: torch.device, warmup_steps: int = 3,
                 kjt_capacity: Optional[int] = None) -> None:
        self._model, self._opt = model, optimizer
        self._check_capturable(optimizer)
        self._keys = list(keys)
        self._dev = torch.device(device)
        F = len(self._keys)
        self._ids = torch.zeros(F, batch_size, dtype=torch.int64, device=self._dev)
        self._labels = torch.zeros(batch_size, dtype=torch.int32, device=self._dev)
        self._rows = torch.tensor(list(num_embeddings), dtype=torch.int64, device=self._dev)
        self._dense = torch.zeros(1, device=self._dev)
        # multi-hot mode: the KJT's values (fixed capacity) and lengths are the static inputs instead of id columns
        self._values = self._lengths = None
        if kjt_capacity is not None:
            self._values = torch.zeros(int(kjt_capacity), dtype=torch.int64, device=self._dev)
            self._lengths = torch.zeros(F * batch_size, dtype=torch.int32, device=self._dev)
        self._warmup = warmup_steps
        self._calls = 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._out = None
        self._stream = torch.cuda.Stream(device=self._dev)
        # sharded modules whose row-wise peer exchange adds into pre-cleared buffers (see PeerExchange.clean)
        self._sharded = [m for m in model.modules() if isinstance(getattr(m, "_peer", None), dict)]
        for m in self._sharded:
            if getattr(m, "dp_ebc", None) is not None:
                # their update runs in Python after an all-reduce (row-wise Adam's bias correction is a host-side step count)
                raise NotImplementedError("CudaGraphTrainStep: data_parallel embedding tables are updated outside the kernels "
                                          "(sharding.py: sync_data_parallel); run the step eagerly (TrainPipelineSparseDist) or "
                                          "shard these tables table_wise / row_wise")

    def _clean_exchanges(self) -> None:
        """A captured step assumes the scatter-add buffer it was captured with is clear; eager forwards without a
        backward in between replays (an evaluation pass) leave theirs dirty -- clear those first.  Every rank runs
        the same sequence of calls, so either all ranks take the barrier inside or none does."""
        for m in self._sharded:
            for ex in m._peer.values():
                ex.clean()

    @staticmethod
    def _check_capturable(optimizer) -> None:
        """An optimizer whose step() bakes host-computed, step-dependent scalars into its launches would replay
        them frozen: refuse it instead of training silently wrong."""
        inner = getattr(optimizer, "_optimizer", optimizer)
        if isinstance(inner, (torch.optim.Adam, torch.optim.AdamW)) and not any(g.get("capturable") for g in inner.param_groups):
            raise ValueError("CudaGraphTrainStep: torch.optim.Adam keeps its step count on the host; use tt.FlatAdam "
                             "(device-side step counter), torch.optim.SGD, or Adam(capturable=True)")

    def _step(self):
        if self._values is not None:
            kjt = KeyedJaggedTensor(keys=self._keys, values=self._values, lengths=self._lengths)
            kjt._values_padded = True
        else:
            kjt = KeyedJaggedTensor.from_id_columns(self._keys, self._ids, self._rows)
        batch = Batch(dense_features=self._dense, sparse_features=kjt, labels=self._labels)
        self._opt.zero_grad()
        loss, out = self._model(batch)
        loss.backward()
        sync = getattr(self._model, "sync_dense_grads", None)   # DistributedModelParallel: all-reduce of the tower gradients
        if sync is not None:
            sync()
        self._opt.step()
        return out

    def __call__(self, ids: torch.Tensor, labels: torch.Tensor):
        """``ids`` [F, B] int64 and ``labels`` [B] int32 (pinned host or device).  Returns the model's
        second output ``(loss, logits, labels)``; the tensors are static buffers, overwritten by the next call."""
        self._ids.copy_(ids, non_blocking=True)
        self._labels.copy_(labels, non_blocking=True)
        return self._run()

    def step_kjt(self, values: torch.Tensor, lengths: torch.Tensor, labels: torch.Tensor):
        """Multi-hot batch (``kjt_capacity`` mode): ``values`` int64 [n <= capacity] in key-major KJT order, ``lengths``
        int32 [F*B], ``labels`` [B] (pinned host or device).  ``sum(lengths)`` must equal ``n``; what lies past ``n`` in
        the static buffer is never read as an id."""
        if self._values is None:
            raise ValueError("CudaGraphTrainStep.step_kjt needs kjt_capacity= at construction")
        n = values.numel()
        if n > self._values.numel():
            raise ValueError(f"batch holds {n} ids, capacity is {self._values.numel()}")
        self._values[:n].copy_(values, non_blocking=True)
        self._lengths.copy_(lengths, non_blocking=True)
        self._labels.copy_(labels, non_blocking=True)
        return self._run()

    def _run(self):
        self._calls += 1
        cur = torch.cuda.current_stream(self._dev)
        if self._calls <= self._warmup:
            # eager warm-up on the SAME side stream the capture will use (autograd's stream
            # bookkeeping for the persistent .grad buffers must not point at another stream)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                out = self._step()
            cur.wait_stream(self._stream)
            return out
        self._clean_exchanges()
        if self._graph is None:
            torch.cuda.synchronize(self._dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._out = self._step()
            self._graph = g
        self._graph.replay()
        return self._out

    @property
    def captured(self) -> bool:
        return self._graph is not None
