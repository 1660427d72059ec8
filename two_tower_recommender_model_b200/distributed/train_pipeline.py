"""``TrainPipelineSparseDist(model, optimizer, device)`` with ``.progress(iter)`` as
driven by /root/reference/utils/model_training.py:232,305,335 (private
attributes ``_model`` / ``_optimizer`` / ``_device`` are read at 209-210,281,303,364).

Stages, as in TorchRec: while batch i computes on the current stream, batch i+1
is already on the device (copied by a side stream from pinned host memory) and,
for sharded models, its KJT all-to-all (``input_dist``) has been issued.
``progress`` returns the model's second output ``(loss, logits, labels)``, skips
backward/step when ``model.training`` is false, and raises ``StopIteration`` once
the iterator and the queue are both drained.
"""
from collections import deque
from typing import Any, Deque, Iterator, Optional, Tuple

import torch


class TrainPipelineSparseDist:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, device: torch.device,
                 execute_all_batches: bool = True, depth: int = 2) -> None:
        self._model = model
        self._optimizer = optimizer
        self._device = torch.device(device)
        self._execute_all_batches = execute_all_batches
        self._depth = depth
        self._batches: Deque[Tuple[Any, Optional[torch.cuda.Event]]] = deque()
        self._iter_done = False
        self._last_iter: Optional[Iterator] = None
        self._memcpy_stream = torch.cuda.Stream(device=self._device) if self._device.type == "cuda" else None

    # -- stage 1: host -> device on the memcpy stream
    def _enqueue(self, it: Iterator) -> bool:
        if self._iter_done:
            return False
        try:
            batch = next(it)
        except StopIteration:
            self._iter_done = True
            return False
        ev = None
        if self._memcpy_stream is not None:
            with torch.cuda.stream(self._memcpy_stream):
                batch = batch.to(self._device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._memcpy_stream)
        else:
            batch = batch.to(self._device)
        # -- stage 2: sparse input dist for sharded models (issued ahead of use)
        start = getattr(self._model, "start_sparse_data_dist", None)
        if start is not None:
            batch = start(batch, ev)
        self._batches.append((batch, ev))
        return True

    def _fill(self, it: Iterator) -> None:
        if it is not self._last_iter:
            # a new iterator (new epoch / eval pass): the reference builds one per call site
            self._last_iter = it
            self._iter_done = False
        while len(self._batches) < self._depth and self._enqueue(it):
            pass

    def progress(self, dataloader_iter: Iterator) -> Any:
        self._fill(dataloader_iter)
        if not self._batches:
            raise StopIteration
        training = self._model.training
        if training:
            self._optimizer.zero_grad()
        batch, ev = self._batches.popleft()
        cur = torch.cuda.current_stream(self._device) if self._memcpy_stream is not None else None
        if ev is not None:
            cur.wait_event(ev)
            batch.record_stream(cur)
        # keep the copy of the next batch in flight while this one computes
        self._fill(dataloader_iter)
        if training:
            loss, output = self._model(batch)
            loss.backward()
            sync = getattr(self._model, "sync_dense_grads", None)
            if sync is not None:
                sync()  # data-parallel towers: one all-reduce over the flat gradient buffer
            self._optimizer.step()
        else:
            with torch.no_grad():
                loss, output = self._model(batch)
        return output
