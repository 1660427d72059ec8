"""Optimizer state inside the checkpoint, host logic on CPU: ``include_optimizer_state(True)`` puts the
fused row-wise state under ``embedding_bags.<table>.sum`` (+ ``fused_optimizer_step``) in ``state_dict()``; a module that
loads it takes the SAME next step as the one that kept training, one that loads the reference's weights-only format
(utils/model_training.py:161-189) does not.  The device work (pooled lookup + fused row-wise Adagrad in the backward) is
replaced by the oracle stand-in of tests/test_reference_boundary.py; the kernels are held to the oracle on the GPU."""
import os
import sys

import pytest
import torch
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import two_tower_recommender_model_b200 as tt  # noqa: E402
from helpers import random_kjt  # noqa: E402
from test_reference_boundary import _OracleLookup  # noqa: E402

KEYS, ROWS, DIM, B, LR = ["u", "i"], [23, 31], 8, 16, 0.1


@pytest.fixture()
def oracle_device_work(monkeypatch):
    from two_tower_recommender_model_b200 import _native as N
    from two_tower_recommender_model_b200.modules import embedding_modules
    monkeypatch.setattr(embedding_modules, "EbcLookup", _OracleLookup)
    monkeypatch.setattr(N, "require_cuda", lambda t, name: None)


def _ebc(seed):
    torch.manual_seed(seed)
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=DIM, num_embeddings=ROWS[j], feature_names=[k]) for j, k in enumerate(KEYS)]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("cpu"))
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": LR})
    return ebc


def _step(ebc, s):
    v, l = random_kjt(KEYS, ROWS, B, 3, seed=40 + s, dup_pool=6)          # a small id pool: rows are hit in several steps
    out = ebc(tt.KeyedJaggedTensor.from_lengths_sync(KEYS, v, l)).values()
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(90 + s))
    (out * g).sum().backward()


def test_fused_optimizer_state_travels_in_the_state_dict(oracle_device_work):
    a = _ebc(0)
    for s in range(2):
        _step(a, s)
    plain = a.state_dict()
    assert set(plain) == {"embedding_bags.t_u.weight", "embedding_bags.t_i.weight"}          # the reference's format
    full = {k: v.clone() for k, v in a.include_optimizer_state(True).state_dict().items()}
    assert set(full) == set(plain) | {"embedding_bags.t_u.sum", "embedding_bags.t_i.sum", "fused_optimizer_step"}
    assert float(full["embedding_bags.t_u.sum"].max()) > 0 and full["embedding_bags.t_u.sum"].shape == (ROWS[0],)

    b = _ebc(1)                      # resumes WITH the optimizer state
    b.load_state_dict(full)
    c = _ebc(2)                      # resumes from a weights-only checkpoint: the accumulators restart at zero
    c.load_state_dict({k: v.clone() for k, v in plain.items()})
    for m in (a, b, c):
        _step(m, 2)
    for k in plain:
        assert torch.equal(a.state_dict()[k], b.state_dict()[k]), k
    for t in ("t_u", "t_i"):
        assert torch.equal(a.fused_optimizer_state()[t]["sum"], b.fused_optimizer_state()[t]["sum"])
    # the weights-only resume takes a different (larger: sqrt(sum) restarted) step on rows both runs had already visited
    assert not torch.equal(a.state_dict()["embedding_bags.t_u.weight"], c.state_dict()["embedding_bags.t_u.weight"])


def test_optimizer_state_keys_under_the_two_tower_prefix(oracle_device_work):
    """Through the TwoTower wrapper the entries sit under ``ebc.`` like the weights; a strict load of either format works."""
    ebc = _ebc(3)
    tower = tt.TwoTower(ebc, [8, 4], device=torch.device("cpu"))
    _step(tower.ebc, 0)
    tower.ebc.include_optimizer_state(True)
    sd = tower.state_dict()
    assert {"ebc.embedding_bags.t_u.sum", "ebc.embedding_bags.t_i.sum", "ebc.fused_optimizer_step"} <= set(sd)
    other = tt.TwoTower(_ebc(4), [8, 4], device=torch.device("cpu"))
    res = other.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(other.ebc.fused_optimizer_state()["t_i"]["sum"], tower.ebc.fused_optimizer_state()["t_i"]["sum"])
    weights_only = {k: v for k, v in sd.items() if k.endswith(".weight") or k.endswith(".bias")}
    res = tt.TwoTower(_ebc(5), [8, 4], device=torch.device("cpu")).load_state_dict(weights_only, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_flat_adam_state_dict_round_trip(monkeypatch):
    """FlatAdam keeps its moments and step in flat buffers outside ``Optimizer.state``; ``state_dict()`` carries them
    (also through ``KeyedOptimizerWrapper`` and ``torch.save``), so a resumed run takes the same next step.  The one
    device call (``tt_adam_flat_devstep``) is replaced by the same arithmetic in torch; tests/test_gpu_dense.py holds
    the kernel to ``torch.optim.Adam``."""
    import io

    from two_tower_recommender_model_b200 import _native as N
    live = []

    def fake_call(name, p, g, m, v, n, lr, b1, b2, eps, step_ptr, stream):
        assert name == "tt_adam_flat_devstep"
        o = next(x for x in live if x.flat_param.data_ptr() == p)
        assert (g, m, v, n, step_ptr) == (o.flat_grad.data_ptr(), o.exp_avg.data_ptr(), o.exp_avg_sq.data_ptr(), o.flat_param.numel(),
                                          o.step_dev.data_ptr())
        o.step_dev += 1
        t = float(o.step_dev)
        o.exp_avg.mul_(b1).add_(o.flat_grad, alpha=1 - b1)
        o.exp_avg_sq.mul_(b2).addcmul_(o.flat_grad, o.flat_grad, value=1 - b2)
        o.flat_param.addcdiv_(o.exp_avg / (1 - b1 ** t), (o.exp_avg_sq / (1 - b2 ** t)).sqrt() + eps, value=-lr)

    monkeypatch.setattr(N, "require_cuda", lambda t, name: None)
    monkeypatch.setattr(N, "call", fake_call)
    monkeypatch.setattr(N, "stream_ptr", lambda dev: 0)

    def build(seed):
        torch.manual_seed(seed)
        net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
        opt = tt.KeyedOptimizerWrapper(dict(net.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-2))
        live.append(opt._optimizer)
        return net, opt

    x = torch.randn(16, 4, generator=torch.Generator().manual_seed(3))

    def step(net, opt):
        opt.zero_grad()
        net(x).pow(2).mean().backward()
        opt.step()

    a, oa = build(0)
    want = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
    want.load_state_dict(a.state_dict())
    ow = torch.optim.Adam(want.parameters(), lr=1e-2)
    for _ in range(2):
        step(a, oa)
        step(want, ow)
    for k, v in want.state_dict().items():                         # the stand-in is Adam: FlatAdam's views / zero_grad bookkeeping hold
        torch.testing.assert_close(a.state_dict()[k], v, rtol=1e-6, atol=1e-7)
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    osd = oa.state_dict()
    assert osd["flat_adam"]["step"] == 2.0 and osd["flat_adam"]["exp_avg"].numel() == sum(p.numel() for p in a.parameters())
    buf = io.BytesIO()
    torch.save(osd, buf)
    buf.seek(0)
    b, ob = build(1)                                               # resumed: weights + optimizer state (through torch.save / load)
    b.load_state_dict(sd)
    ob.load_state_dict(torch.load(buf))
    assert ob.param_groups is ob._optimizer.param_groups and ob._optimizer.step_count == 2
    c, oc = build(2)                                               # resumed from the weights only
    c.load_state_dict(sd)
    for net, opt in ((a, oa), (b, ob), (c, oc)):
        step(net, opt)
    for k in sd:
        assert torch.equal(a.state_dict()[k], b.state_dict()[k]), k
    assert not torch.equal(a.state_dict()["0.weight"], c.state_dict()["0.weight"])      # Adam's moments restarted: another step
    with pytest.raises(ValueError):
        ob._optimizer.load_state_dict(dict(osd, flat_adam=dict(osd["flat_adam"], exp_avg=torch.zeros(3))))


def test_optimizer_state_of_the_wrong_size_is_refused(oracle_device_work):
    """The fused update indexes the accumulators by row: a checkpoint entry of another table size must fail the load (as a
    weight of the wrong size does), not become a buffer the kernel writes past."""
    a = _ebc(0)
    _step(a, 0)
    sd = {k: v.clone() for k, v in a.include_optimizer_state(True).state_dict().items()}
    sd["embedding_bags.t_u.sum"] = sd["embedding_bags.t_u.sum"][:-3].clone()
    b = _ebc(1)
    with pytest.raises(RuntimeError, match="size mismatch for embedding_bags.t_u.sum"):
        b.load_state_dict(sd)
    assert "t_u" not in b.fused_optimizer_state() or "sum" not in b.fused_optimizer_state()["t_u"]
    with pytest.raises(ValueError, match="t_i.sum"):
        b.load_fused_optimizer_state({"t_i": {"sum": torch.zeros(ROWS[1] + 1)}})
