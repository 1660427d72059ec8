"""JaggedTensor / KeyedJaggedTensor / KeyedTensor with the call surface the
reference uses from ``torchrec.sparse.jagged_tensor``:

* ``KeyedJaggedTensor.from_lengths_sync(keys, values, lengths)``
  (/root/reference/utils/model_training.py:57)
* ``KeyedJaggedTensor(keys=, values=, lengths=)``
  (/root/reference/03_model_training.py:1081-1085)
* ``.keys() .values() .lengths() .length_per_key() .to_dict() .to(device)``
  (03_model_training.py:1088-1091,1156; workshop/02-mosaic-model-training.py:970)
* ``kjt[key].values()`` (ray_tune_optuna_tuning_alex_test.py:372)

Layout (TorchRec): key-major -- ``lengths`` is ``[F*B]`` with
``lengths[f*B + b]``, ``values`` concatenated in the same order, ``offsets`` the
complete cumsum.  Containers accept CPU tensors (that is where the reference
builds them) and CUDA tensors; on CUDA the bookkeeping (cumsum, permute) runs in
libtt_b200.so.
"""
from typing import Dict, List, Optional, Tuple

import torch

from .. import _native as N


def _complete_cumsum(lengths: torch.Tensor) -> torch.Tensor:
    n = lengths.numel()
    if lengths.is_cuda:
        lengths = lengths.contiguous()
        if lengths.dtype != torch.int32:
            lengths = lengths.to(torch.int32)
        out = torch.empty(n + 1, dtype=torch.int32, device=lengths.device)
        ws = N.workspace(N.load().tt_kjt_offsets_workspace_bytes(n), lengths.device)
        N.call("tt_kjt_lengths_to_offsets", N.ptr(lengths), N.ptr(out), n, N.ptr(ws), ws.numel(),
               N.stream_ptr(lengths.device))
        return out
    out = torch.zeros(n + 1, dtype=lengths.dtype)
    torch.cumsum(lengths, 0, out=out[1:])
    return out


class JaggedTensor:
    def __init__(self, values: torch.Tensor, weights: Optional[torch.Tensor] = None,
                 lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None) -> None:
        assert lengths is not None or offsets is not None, "lengths or offsets required"
        self._values = values
        self._weights = weights
        self._lengths = lengths
        self._offsets = offsets

    def values(self) -> torch.Tensor:
        return self._values

    def weights_or_none(self) -> Optional[torch.Tensor]:
        return self._weights

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            self._lengths = self._offsets[1:] - self._offsets[:-1]
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            self._offsets = _complete_cumsum(self._lengths)
        return self._offsets

    def to(self, device, non_blocking: bool = False) -> "JaggedTensor":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return JaggedTensor(mv(self._values), mv(self._weights), mv(self._lengths), mv(self._offsets))

    def to_dense(self) -> List[torch.Tensor]:
        off = self.offsets().tolist()
        return [self._values[off[i]:off[i + 1]] for i in range(len(off) - 1)]

    def __repr__(self) -> str:
        return f"JaggedTensor(values={self._values}, lengths={self.lengths()})"


class KeyedJaggedTensor:
    def __init__(self, keys: List[str], values: torch.Tensor, weights: Optional[torch.Tensor] = None,
                 lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                 stride: Optional[int] = None, length_per_key: Optional[List[int]] = None,
                 offset_per_key: Optional[List[int]] = None) -> None:
        assert lengths is not None or offsets is not None, "lengths or offsets required"
        self._keys = list(keys)
        self._values = values
        self._weights = weights
        self._lengths = lengths
        self._offsets = offsets
        n = lengths.numel() if lengths is not None else offsets.numel() - 1
        if stride is None:
            stride = n // len(self._keys) if self._keys else 0
        assert stride * len(self._keys) == n, "lengths must hold len(keys) * stride entries"
        self._stride = stride
        self._length_per_key = length_per_key
        self._offset_per_key = offset_per_key

    # ---- constructors
    @staticmethod
    def from_lengths_sync(keys: List[str], values: torch.Tensor, lengths: torch.Tensor,
                          weights: Optional[torch.Tensor] = None) -> "KeyedJaggedTensor":
        kjt = KeyedJaggedTensor(keys=keys, values=values, weights=weights, lengths=lengths)
        kjt.offsets()
        kjt.length_per_key()  # "sync": materialises the per-key totals on the host
        return kjt

    @staticmethod
    def from_offsets_sync(keys: List[str], values: torch.Tensor, offsets: torch.Tensor,
                          weights: Optional[torch.Tensor] = None) -> "KeyedJaggedTensor":
        kjt = KeyedJaggedTensor(keys=keys, values=values, weights=weights, offsets=offsets)
        kjt.length_per_key()
        return kjt

    @staticmethod
    def from_id_columns(keys: List[str], ids: torch.Tensor, num_embeddings: torch.Tensor,
                        row_range: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> "KeyedJaggedTensor":
        """Device-side ``transform_to_torchrec_batch`` (utils/model_training.py:43-61):
        ``ids`` is ``[F, B]`` int64 on CUDA; id 0 -> empty bag, else
        ``id % num_embeddings[f]`` with length 1.  No host sync: ``values`` keeps
        capacity ``F*B`` and the live count stays on the device (``offsets[-1]``)."""
        N.require_cuda(ids, "ids")
        F, B = ids.shape
        ids = ids.contiguous()
        dev = ids.device
        values = torch.empty(F * B, dtype=torch.int64, device=dev)
        lengths = torch.empty(F * B, dtype=torch.int32, device=dev)
        offsets = torch.empty(F * B + 1, dtype=torch.int32, device=dev)
        ne = num_embeddings.to(device=dev, dtype=torch.int64).contiguous()
        ws = N.workspace(N.load().tt_kjt_from_columns_workspace_bytes(F, B), dev)
        if row_range is not None:
            # one row-wise shard: keep the ids of rows [lo[f], hi[f]) (values become shard-local), other bags are empty
            lo, hi = row_range
            N.call("tt_kjt_from_columns_range", N.ptr(ids), N.ptr(ne), N.ptr(lo), N.ptr(hi), F, B, N.ptr(values), N.ptr(lengths),
                   N.ptr(offsets), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
            kjt = KeyedJaggedTensor(keys=keys, values=values, lengths=lengths, offsets=offsets, stride=B)
            kjt._values_padded = True
            return kjt
        N.call("tt_kjt_from_columns", N.ptr(ids), N.ptr(ne), F, B, N.ptr(values), N.ptr(lengths), N.ptr(offsets),
               N.ptr(ws), ws.numel(), N.stream_ptr(dev))
        kjt = KeyedJaggedTensor(keys=keys, values=values, lengths=lengths, offsets=offsets, stride=B)
        kjt._values_padded = True
        # kept for model-parallel input dists: single-id features can travel as dense id columns
        # (fixed sizes: no count exchange, no host sync) and be turned into a KJT where they land
        kjt._id_columns = (ids, ne)
        return kjt

    # ---- accessors
    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def weights_or_none(self) -> Optional[torch.Tensor]:
        return self._weights

    def stride(self) -> int:
        return self._stride

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            self._lengths = self._offsets[1:] - self._offsets[:-1]
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            self._offsets = _complete_cumsum(self._lengths)
        return self._offsets

    def length_per_key(self) -> List[int]:
        if self._length_per_key is None:
            if not self._keys:
                self._length_per_key = []
            else:
                self._length_per_key = self.lengths().view(len(self._keys), self._stride).sum(dim=1).tolist()
        return self._length_per_key

    def offset_per_key(self) -> List[int]:
        if self._offset_per_key is None:
            acc, out = 0, [0]
            for n in self.length_per_key():
                acc += n
                out.append(acc)
            self._offset_per_key = out
        return self._offset_per_key

    @property
    def device(self) -> torch.device:
        return self._values.device

    # ---- per-key views
    def __getitem__(self, key: str) -> JaggedTensor:
        f = self._keys.index(key)
        opk = self.offset_per_key()
        B = self._stride
        lengths = self.lengths()[f * B:(f + 1) * B]
        offs = self.offsets()[f * B:(f + 1) * B + 1] - opk[f]
        w = None if self._weights is None else self._weights[opk[f]:opk[f + 1]]
        return JaggedTensor(values=self._values[opk[f]:opk[f + 1]], weights=w, lengths=lengths, offsets=offs)

    def to_dict(self) -> Dict[str, JaggedTensor]:
        return {k: self[k] for k in self._keys}

    # ---- movement
    def to(self, device, non_blocking: bool = False) -> "KeyedJaggedTensor":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        out = KeyedJaggedTensor(self._keys, mv(self._values), mv(self._weights), mv(self._lengths),
                                mv(self._offsets), self._stride, self._length_per_key, self._offset_per_key)
        return out

    def pin_memory(self) -> "KeyedJaggedTensor":
        pm = lambda t: None if t is None else t.pin_memory()
        return KeyedJaggedTensor(self._keys, pm(self._values), pm(self._weights), pm(self._lengths),
                                 pm(self._offsets), self._stride, self._length_per_key, self._offset_per_key)

    def record_stream(self, stream) -> None:
        for t in (self._values, self._weights, self._lengths, self._offsets):
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    # ---- reordering (fbgemm::permute_2D_sparse_data)
    def permute(self, indices: List[int]) -> "KeyedJaggedTensor":
        keys = [self._keys[i] for i in indices]
        B = self._stride
        lpk = self.length_per_key()
        total = sum(lpk[i] for i in indices)
        if self._values.is_cuda:
            if self._weights is not None:
                raise NotImplementedError("weighted KJT permute is not implemented on CUDA")
            dev = self._values.device
            T_out = len(indices)
            perm = torch.tensor(indices, dtype=torch.int32, device=dev)
            lengths = self.lengths().contiguous()
            if lengths.dtype != torch.int32:
                lengths = lengths.to(torch.int32)
            out_len = torch.empty(T_out * B, dtype=torch.int32, device=dev)
            out_off = torch.empty(T_out * B + 1, dtype=torch.int32, device=dev)
            out_val = torch.empty(total, dtype=torch.int64, device=dev)
            ws = N.workspace(N.load().tt_kjt_permute_workspace_bytes(T_out, B), dev)
            N.call("tt_kjt_permute_2d", N.ptr(perm), T_out, len(self._keys), B, N.ptr(lengths),
                   N.ptr(self.offsets().contiguous()), N.ptr(self._values.contiguous()), N.ptr(out_len),
                   N.ptr(out_off), N.ptr(out_val), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
            return KeyedJaggedTensor(keys, out_val, None, out_len, out_off, B, [lpk[i] for i in indices])
        opk = self.offset_per_key()
        lengths = self.lengths().view(len(self._keys), B)
        out_len = lengths[indices].reshape(-1) if indices else lengths[:0].reshape(-1)
        segs = [self._values[opk[i]:opk[i + 1]] for i in indices]
        out_val = torch.cat(segs) if segs else self._values[:0]
        w = None
        if self._weights is not None:
            ws_ = [self._weights[opk[i]:opk[i + 1]] for i in indices]
            w = torch.cat(ws_) if ws_ else self._weights[:0]
        return KeyedJaggedTensor(keys, out_val, w, out_len.contiguous(), None, B, [lpk[i] for i in indices])

    def split(self, segments: List[int]) -> List["KeyedJaggedTensor"]:
        out, start = [], 0
        opk = self.offset_per_key()
        B = self._stride
        for seg in segments:
            end = start + seg
            keys = self._keys[start:end]
            lengths = self.lengths()[start * B:end * B]
            vals = self._values[opk[start]:opk[end]]
            w = None if self._weights is None else self._weights[opk[start]:opk[end]]
            out.append(KeyedJaggedTensor(keys, vals, w, lengths, None, B, self.length_per_key()[start:end]))
            start = end
        return out

    def __repr__(self) -> str:
        return (f"KeyedJaggedTensor(keys={self._keys}, stride={self._stride}, "
                f"values={tuple(self._values.shape)}, device={self._values.device})")


class KeyedTensor:
    """Pooled output of an EmbeddingBagCollection: ``values`` is ``[B, sum(dims)]``,
    ``kt[key]`` the ``[B, dim]`` column block of that feature
    (utils/model_training.py:106,115)."""

    def __init__(self, keys: List[str], length_per_key: List[int], values: torch.Tensor, key_dim: int = 1) -> None:
        self._keys = list(keys)
        self._length_per_key = list(length_per_key)
        self._values = values
        self._key_dim = key_dim
        acc, off = 0, [0]
        for n in self._length_per_key:
            acc += n
            off.append(acc)
        self._offset_per_key = off

    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def length_per_key(self) -> List[int]:
        return self._length_per_key

    def offset_per_key(self) -> List[int]:
        return self._offset_per_key

    def key_dim(self) -> int:
        return self._key_dim

    def __getitem__(self, key: str) -> torch.Tensor:
        i = self._keys.index(key)
        return self._values.narrow(self._key_dim, self._offset_per_key[i], self._length_per_key[i])

    def to_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self[k] for k in self._keys}

    def columns(self, keys: List[str]) -> Tuple[int, int]:
        """(first column, width) when ``keys`` are adjacent in this tensor, else (-1, -1)."""
        idx = [self._keys.index(k) for k in keys]
        if idx != list(range(idx[0], idx[0] + len(idx))):
            return -1, -1
        return self._offset_per_key[idx[0]], self._offset_per_key[idx[-1] + 1] - self._offset_per_key[idx[0]]

    def __repr__(self) -> str:
        return f"KeyedTensor(keys={self._keys}, values={tuple(self._values.shape)})"
