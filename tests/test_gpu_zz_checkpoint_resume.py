"""SURVEY.md 8(f) N3 on the GPU: save -> reload -> identical next step.  The model checkpoint carries the fused row-wise
Adagrad accumulators (``include_optimizer_state(True)``), the dense optimizer's ``state_dict()`` carries FlatAdam's moments and
step; a run resumed from both takes the same third step as the run that never stopped, a run resumed from the reference's
weights-only format (utils/model_training.py:161-189) does not.  Host logic of the same round trip: tests/test_checkpoint_resume.py."""
import pytest
import torch
from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward

pytestmark = pytest.mark.gpu

CAT = ["user_id", "product_id"]


def test_save_reload_next_step_identical(cuda):
    import two_tower_recommender_model_b200 as tt
    emb, dim, layers, B, lr = [300, 200], 64, [128, 64], 512, 0.05

    def build(seed):
        torch.manual_seed(seed)
        ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=dim, num_embeddings=emb[i], feature_names=[c])
                                                for i, c in enumerate(CAT)], device=cuda)
        task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, layers, device=cuda))           # fp32 kernels, BCE: the reference's loss
        apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": lr})
        opt = tt.KeyedOptimizerWrapper(dict(task.named_parameters()), lambda p: tt.FlatAdam(p, lr=1e-2))
        return task, opt

    g = torch.Generator().manual_seed(5)
    data = [(torch.stack([torch.randint(0, 2 * emb[0], (B,), generator=g), torch.randint(0, 2 * emb[1], (B,), generator=g)]),
             torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)) for _ in range(3)]
    rows = torch.tensor(emb, device=cuda)

    def step(task, opt, i):
        ids, y = data[i]
        batch = tt.Batch(torch.zeros(1, device=cuda), tt.KeyedJaggedTensor.from_id_columns(CAT, ids.to(cuda), rows), y.to(cuda))
        opt.zero_grad()
        loss, _ = task(batch)
        loss.backward()
        opt.step()
        return float(loss.detach())

    a, oa = build(0)
    for i in range(2):
        step(a, oa, i)
    plain = {k: v.detach().clone() for k, v in a.state_dict().items()}                # the reference's weights-only format
    assert all(k.endswith(".weight") or k.endswith(".bias") for k in plain)
    a.two_tower.ebc.include_optimizer_state(True)
    full = {k: v.detach().clone() for k, v in a.state_dict().items()}
    assert {"two_tower.ebc.embedding_bags.t_user_id.sum", "two_tower.ebc.embedding_bags.t_product_id.sum",
            "two_tower.ebc.fused_optimizer_step"} == set(full) - set(plain)
    osd = oa.state_dict()
    assert float(osd["flat_adam"]["step"]) == 2.0

    b, ob = build(1)                     # resumed with model + optimizer state
    b.load_state_dict(full)
    ob.load_state_dict(osd)
    c, oc = build(2)                     # resumed from the weights only, fresh optimizers
    c.load_state_dict(plain)
    la, lb, lc = step(a, oa, 2), step(b, ob, 2), step(c, oc, 2)
    assert la == pytest.approx(lb, rel=1e-5) and la == pytest.approx(lc, rel=1e-5)    # same weights -> same loss at the step
    sa, sb, sc = a.state_dict(), b.state_dict(), c.state_dict()
    for k in plain:
        torch.testing.assert_close(sb[k], sa[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
    for t in ("t_user_id", "t_product_id"):
        torch.testing.assert_close(b.two_tower.ebc.fused_optimizer_state()[t]["sum"], a.two_tower.ebc.fused_optimizer_state()[t]["sum"],
                                   rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(ob._optimizer.exp_avg, oa._optimizer.exp_avg, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(ob._optimizer.exp_avg_sq, oa._optimizer.exp_avg_sq, rtol=1e-5, atol=1e-12)
    # without the optimizer state the third step is a different one: Adagrad's sqrt(sum) and Adam's moments restarted
    k_t, k_w = "two_tower.ebc.embedding_bags.t_user_id.weight", "two_tower.query_proj._mlp.0._linear.weight"
    assert float((sc[k_t] - sa[k_t]).abs().max()) > 1e-4 and float((sc[k_w] - sa[k_w]).abs().max()) > 1e-5
