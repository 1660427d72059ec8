"""Buffer bookkeeping of the row-wise (scatter-add) peer exchange, on CPU.

The real ``_PeerTwLookup`` autograd node (two_tower_recommender_model_b200/distributed/sharding.py) runs here against
a fake exchange: plain CPU tensors stand for the symmetric buffers, the two kernels are replaced by "every shard
adds 1 to every element" / a no-op, and the barriers only count.  What is checked is the protocol the kernels rely
on -- whatever the order of training steps, evaluation forwards and captured-graph replays, the buffer a forward adds
into is clear when the adds arrive, so its output is the sum of THIS forward's contributions alone (TorchRec's
reduce-scatter semantics, torchrec PooledEmbeddingsReduceScatter; the reference alternates train / evaluate per epoch,
/root/reference/utils/model_training.py:255-317 and :320-372).
"""
import itertools

import pytest
import torch

from two_tower_recommender_model_b200 import _native as N
from two_tower_recommender_model_b200.distributed import sharding as S


class FakeExchange:
    """The attributes and methods of PeerExchange that _PeerTwLookup touches."""
    rows, cols, world = 4, 3, 2

    def __init__(self):
        self._pooled2 = torch.zeros(2, self.rows, self.cols)
        self._grad2 = torch.zeros(2, self.rows, self.cols)
        self.cur, self.dirty, self.forward_open = 0, [False, False], False
        self.barriers = 0

    flip = S.PeerExchange.flip
    pooled = S.PeerExchange.pooled
    grad = S.PeerExchange.grad
    clean = S.PeerExchange.clean

    def pooled_peers(self, b, scatter_add):
        return N.PeerBuffers()

    def grad_peers(self, b):
        return N.PeerBuffers()

    def barrier_pooled(self):
        self.barriers += 1

    def barrier_grad(self):
        self.barriers += 1


class FakeLocalEbc:
    def _build_plan(self, *a, **k):
        return N.EbcPlan(), 0

    def _sparse_optimizer_spec(self, advance_step=True):
        return N.SparseOptimizer(kind=N.OPT_ROWWISE_ADAGRAD)


@pytest.fixture
def fake_kernels(monkeypatch):
    state = {"ex": None, "fwd": 0, "bwd": 0}

    def call(name, *args):
        ex = state["ex"]
        if name == "tt_ebc_forward_peer":
            ex._pooled2[ex.cur] += 1.0          # this rank's shard and the peer's shard each add their partial sums
            ex._pooled2[ex.cur] += 1.0
            state["fwd"] += 1
        elif name == "tt_ebc_backward_fused_peer":
            state["bwd"] += 1
        else:
            raise AssertionError(name)

    monkeypatch.setattr(N, "call", call)
    monkeypatch.setattr(N, "stream_ptr", lambda dev: 0)
    return state


def _forward(ex, train, scatter_add=True):
    values = torch.zeros(1, dtype=torch.int64)
    offsets = torch.zeros(2, dtype=torch.int32)
    anchor = torch.zeros(0, requires_grad=train)
    with torch.set_grad_enabled(train):
        out = S._PeerTwLookup.apply(ex, FakeLocalEbc(), (ex.cols, {}), ("f",), values, offsets, scatter_add, None, anchor)
    return out


def _run(ex, seq):
    """seq: 'T' = training step (forward + backward), 'E' = evaluation forward, 'C' = what CudaGraphTrainStep does before
    a replay followed by a replay of a step captured on buffer 1 (no Python runs inside a replay: fixed buffer, no flags)."""
    for op in seq:
        if op == "C":
            ex.clean()
            assert float(ex._pooled2[1].abs().max()) == 0.0, f"replay would add into a dirty buffer in {seq}"
            continue                                # the captured step leaves its buffer clear (it holds its own backward)
        out = _forward(ex, train=(op == "T"))
        assert torch.equal(out, torch.full_like(out, 2.0)), f"stale sums in the exchange buffer after {seq!r} at {op}"
        if op == "T":
            out.sum().backward()
            assert not ex.dirty[ex.cur] and float(ex._pooled2[ex.cur].abs().max()) == 0.0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6])
def test_scatter_add_buffer_is_clear_for_every_order_of_train_and_eval(fake_kernels, n):
    for seq in itertools.product("TE", repeat=n):
        ex = FakeExchange()
        fake_kernels["ex"] = ex
        _run(ex, seq + ("T", "T", "E", "T"))


def test_train_eval_train_takes_the_extra_barrier_only_when_needed(fake_kernels):
    ex = FakeExchange()
    fake_kernels["ex"] = ex
    _run(ex, "TTTT")
    assert ex.barriers == 4 * 2                      # a training step: rows landed + gradients staged, nothing else
    ex.barriers = 0
    _run(ex, "E")                                   # first eval forward after a backward: its buffer is clear already
    assert ex.barriers == 1
    _run(ex, "E")                                   # the second lands on a buffer a backward cleared: still no extra barrier
    assert ex.barriers == 2
    _run(ex, "E")                                   # the third reuses the first one's buffer: clear + one more barrier
    assert ex.barriers == 4
    ex.barriers = 0
    _run(ex, "T")                                   # buffer of the second eval forward: dirty -> 3 barriers
    _run(ex, "T")                                   # buffer of the third: dirty too (the round-2 single flag missed this one)
    assert ex.barriers == 6
    _run(ex, "T")
    assert ex.barriers == 8 and ex.dirty == [False, False]


def test_replays_after_eval_forwards_start_from_a_clear_buffer(fake_kernels):
    for evals in range(0, 5):
        ex = FakeExchange()
        fake_kernels["ex"] = ex
        _run(ex, "T")                               # eager warm-up step, lands on buffer 1 = the capture's buffer
        assert ex.cur == 1
        _run(ex, "C" + "E" * evals + "C" + "T" + "C")
        assert ex.dirty == [False, False]


def test_table_wise_replay_after_an_eval_forward_waits_for_the_readers(fake_kernels):
    """Store mode (table-wise) has nothing to clear, but a replay writes the ONE buffer it was captured with: if the last
    operation was an eager forward, a peer may still be reading that buffer -- clean() lines the ranks up, once."""
    ex = FakeExchange()
    fake_kernels["ex"] = ex
    out = _forward(ex, train=True, scatter_add=False)
    out.sum().backward()
    assert ex.barriers == 2 and ex.dirty == [False, False] and not ex.forward_open
    ex.clean()
    assert ex.barriers == 2                          # steady state of a training loop: nothing to do, no barrier
    _forward(ex, train=False, scatter_add=False)
    _forward(ex, train=False, scatter_add=False)
    assert ex.barriers == 4 and ex.forward_open and ex.dirty == [False, False]
    ex.clean()
    assert ex.barriers == 5 and not ex.forward_open
    ex.clean()
    assert ex.barriers == 5
