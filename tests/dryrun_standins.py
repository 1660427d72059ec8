"""Stand-ins that let GPU-only code run on CPU in the dry runs (tests/dryrun_bench.py, tests/dryrun_bench_world2.py): the
DEVICE ENTRY POINTS are replaced by the oracle / plain torch, the CUDA runtime primitives (events, streams, graphs, pinned
memory) by wall-clock / no-op fakes.  Test infrastructure only -- the product has no CPU path and never imports this.
``install()`` patches the modules PROCESS-WIDE: use it in a subprocess."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from oracle.ebc import TableSpec  # noqa: E402

calls = {}          # device entry point -> number of (faked) launches
live_adams = []     # FlatAdam instances, so that the faked tt_adam_flat_devstep finds its buffers


class OracleLookup(torch.autograd.Function):
    """EbcLookup's contract on CPU: pooled [B, sum D] (sum / mean pooling); backward applies the tagged row-wise Adagrad / Adam in
    place on ``grad * ebc._grad_scale`` and the weights get no .grad -- or, without an in-backward optimizer, hands the tables
    their dense gradient."""

    @staticmethod
    def forward(ctx, ebc, kjt_keys, values, offsets, batch, *anchors):
        import two_tower_recommender_model_b200 as tt
        specs = [TableSpec(c.name, c.num_embeddings, c.embedding_dim, list(c.feature_names),
                           "mean" if c.pooling == tt.PoolingType.MEAN else "sum") for c in ebc.embedding_bag_configs()]
        ws = [ebc.embedding_bags[s.name].weight.detach() for s in specs]
        lengths = (offsets[1:] - offsets[:-1]).to(torch.int32)
        n = int(offsets[-1])
        ctx.ebc, ctx.specs, ctx.keys, ctx.n = ebc, specs, list(kjt_keys), len(anchors)
        ctx.save_for_backward(values[:n], lengths)
        return oracle.ebc_forward(specs, ws, list(kjt_keys), values[:n], lengths)

    @staticmethod
    def backward(ctx, g):
        values, lengths = ctx.saved_tensors
        ebc = ctx.ebc
        grads = oracle.ebc_dense_grads(ctx.specs, ctx.keys, values, lengths, g * float(getattr(ebc, "_grad_scale", 1.0)))
        kind = ebc._in_backward_kind()
        if kind is None:
            return (None,) * 5 + tuple(grads)
        from two_tower_recommender_model_b200 import _native as N
        if kind == N.OPT_ROWWISE_ADAM:
            ebc._fused_step += 1          # the kernel advances its device-side counter; here the host-side one stands for it
        for s, gr in zip(ctx.specs, grads):
            w = ebc.embedding_bags[s.name].weight
            cfg = next(c for c in ebc.embedding_bag_configs() if c.name == s.name)
            st = ebc._state_for(cfg, w, kind)
            kw = w._optimizer_kwargs[0]
            if kind == N.OPT_ROWWISE_ADAM:
                hit = (gr != 0).any(dim=1).nonzero().flatten()
                b1, b2 = kw.get("betas", (0.9, 0.999))
                oracle.rowwise_adam_sparse(w.data, st["exp_avg"], st["exp_avg_sq"], hit, gr[hit], step=ebc._fused_step, lr=kw["lr"],
                                           beta1=b1, beta2=b2, eps=kw.get("eps", 1e-8))
            else:
                oracle.rowwise_adagrad_dense(w.data, st["sum"], gr, lr=kw["lr"], eps=kw.get("eps", 1e-10))
        return (None,) * 5 + (None,) * ctx.n


def _linear_act(x, w, b, relu):
    y = torch.nn.functional.linear(x, w, b)
    return torch.relu(y) if relu else y


class FusedTowers:
    """FusedTowersTC's contract in plain torch: tower t = relu(relu(x_t W1^T + b1) W2^T + b2) on the column window
    [cols[t], cols[t] + in_dim) of the pooled matrix; returns one output per tower plus the stacked bf16 copies."""

    @staticmethod
    def apply(pooled, cols, in_dim, grad_dst, *params):
        ys = []
        for t, c0 in enumerate(cols):
            w1, b1, w2, b2 = params[4 * t: 4 * t + 4]
            h = torch.relu(torch.nn.functional.linear(pooled[:, c0:c0 + in_dim], w1, b1))
            ys.append(torch.relu(torch.nn.functional.linear(h, w2, b2)))
        return (*ys, torch.stack([y.detach().bfloat16() for y in ys]))


class _TorchProxy:
    """``torch`` as bench.py sees it in a dry run: everything is the real module, except that a CUDA device is the CPU."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*args, **kwargs):
        return torch.device("cpu")


def run_bench_on_cpu(bench, tiny):
    """Points bench.py's own ``torch`` at the proxy, its workload at `tiny`, its process group at gloo and its probes at small
    sizes, so that ``bench.run_ours(args)`` -- the function the driver's launch reaches -- can run unchanged."""
    import torch.distributed as dist
    bench.torch = _TorchProxy()
    bench.CFG2 = dict(tiny)
    bench.CFG1 = dict(rows=[300, 200], dim=16, layers=[32, 16], batch=32, loss="bce", sparse_lr=0.01, dense_lr=0.001)
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo")
    probe, probe_sharded = bench.retrieval_probe, bench.retrieval_probe_sharded
    bench.retrieval_probe = lambda dev, n_items=3000, n_queries=64, d=64, k=100: probe(dev, min(n_items, 3000), min(n_queries, 64), d, k)
    bench.retrieval_probe_sharded = lambda dev, r, w: probe_sharded(dev, r, w, n_items=3001, q_per_rank=32, d=16, k=10)


def _dot_bce(q, c, labels):
    logits = (q * c).sum(dim=1)
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels.float()), logits.detach()


def _softmax_loss(q, c, temperature=1.0, precision="fp32", negatives="local", pg=None):
    logits = q @ c.t() / temperature
    return torch.nn.functional.cross_entropy(logits, torch.arange(q.shape[0])), logits.diagonal().detach()


def _from_id_columns(keys, ids, num_embeddings, row_range=None):
    """utils/model_training.py:43-61 on [F, B] id columns: id 0 -> empty bag, else id % rows; values keep capacity F * B."""
    from two_tower_recommender_model_b200.sparse.jagged_tensor import KeyedJaggedTensor
    F, B = ids.shape
    ne = torch.as_tensor(num_embeddings).tolist()
    vals, lens = [], []
    for f in range(F):
        keep = ids[f] != 0
        vals.append(ids[f][keep] % ne[f])
        lens.append(keep.to(torch.int32))
    v = torch.cat(vals)
    kjt = KeyedJaggedTensor(keys=list(keys), values=torch.cat([v, torch.zeros(F * B - v.numel(), dtype=torch.int64)]), lengths=torch.cat(lens))
    kjt._values_padded = True
    return kjt


def _bucketize(lengths, offsets, values, num_rows, num_features, batch, world):
    from oracle.kjt import block_bucketize_vectorized
    n = int(offsets[-1])
    nl, nv, unb = block_bucketize_vectorized(lengths, values[:n], torch.as_tensor(num_rows).tolist(), world, batch)
    return nl, oracle.lengths_to_offsets(nl).to(torch.int32), nv, unb


def _fake_call(name, *args):
    calls[name] = calls.get(name, 0) + 1
    if name == "tt_adam_flat_devstep":
        p, g, m, v, n, lr, b1, b2, eps, step_ptr, stream = args
        o = next(x for x in live_adams if x.flat_param.data_ptr() == p)
        o.step_dev += 1
        t = float(o.step_dev)
        o.exp_avg.mul_(b1).add_(o.flat_grad, alpha=1 - b1)
        o.exp_avg_sq.mul_(b2).addcmul_(o.flat_grad, o.flat_grad, value=1 - b2)
        o.flat_param.addcdiv_(o.exp_avg / (1 - b1 ** t), (o.exp_avg_sq / (1 - b2 ** t)).sqrt() + eps, value=-lr)
    elif name not in ("tt_ebc_forward", "tt_set_softmax_backward_mode"):       # launches / switches without an effect here
        raise AssertionError(f"unexpected device call {name}")


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class Graph:
    """A "captured" step is replayed by running it again: same effect on the static buffers as a real replay.  The body of the
    faked ``with torch.cuda.graph(...)`` EXECUTES (a real capture only records), so that execution stands for the first replay."""
    owner = None

    def __init__(self):
        self.fresh = True

    def replay(self):
        if self.fresh:
            self.fresh = False
            return
        Graph.owner._out = Graph.owner._step()


class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def install():
    import two_tower_recommender_model_b200 as tt
    import two_tower_recommender_model_b200.functional as Fn
    from two_tower_recommender_model_b200 import _native as N, retrieval, two_tower as tw_mod
    from two_tower_recommender_model_b200.modules import embedding_modules, mlp
    from two_tower_recommender_model_b200.sparse import jagged_tensor as jt

    N.require_cuda = lambda t, name: None
    N.stream_ptr = lambda dev: 0
    N.call = _fake_call
    N.timing_summary = lambda: {"tt_inbatch_softmax_forward_f32": {"ms": 1.0, "calls": 5}, "tt_inbatch_softmax_backward_f32": {"ms": 1.6, "calls": 5},
                                "tt_ebc_forward": {"ms": 0.03, "calls": 5}, "tt_ebc_backward_fused": {"ms": 0.1, "calls": 5}}
    embedding_modules.EbcLookup = OracleLookup
    mlp.linear_act = _linear_act
    tw_mod.dot_bce_loss = _dot_bce
    tw_mod.FusedTowersTC = FusedTowers
    tw_mod.in_batch_softmax_loss = _softmax_loss
    jt.KeyedJaggedTensor.from_id_columns = staticmethod(_from_id_columns)
    Fn.block_bucketize = _bucketize
    Fn.cast_bf16 = lambda x, **kw: x.bfloat16()
    retrieval.score_topk = lambda q, items, k, item_index_base=0, precision="fp32", items_bf16=None: tuple(
        a + (item_index_base if i else 0) for i, a in enumerate(oracle.exact_topk(q, items if items is not None else items_bf16.float(), k)))

    flat_init = tt.FlatAdam.__init__

    def adam_init(self, *a, **k):
        flat_init(self, *a, **k)
        live_adams.append(self)
    tt.FlatAdam.__init__ = adam_init
    graph_init = tt.CudaGraphTrainStep.__init__

    def gs_init(self, *a, **k):
        graph_init(self, *a, **k)
        Graph.owner = self
    tt.CudaGraphTrainStep.__init__ = gs_init

    torch.cuda.Stream = _Stream
    torch.cuda.current_stream = lambda device=None: _Stream()
    torch.cuda.stream = _Ctx
    torch.cuda.graph = _Ctx
    torch.cuda.CUDAGraph = Graph
    torch.cuda.Event = _Event
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.set_device = lambda d: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
