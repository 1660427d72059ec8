"""Tower MLP, losses, dense Adam and top-k on CUDA vs the oracle / stock torch.
fp32 tolerances: activations rtol 1e-5 atol 1e-5 (sum order), losses rtol 1e-5."""
import pytest
import torch
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,relu", [(1024, 64, 128, True), (1024, 128, 64, True), (1000, 36, 20, False),
                                        (65, 5, 3, True), (1, 128, 1024, True), (4096, 1024, 512, True), (0, 8, 8, True)])
def test_linear_forward_backward(cuda, M, K, N, relu):
    from two_tower_recommender_model_b200.functional import linear_act
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    yr = F.linear(x, w, b)
    yr = torch.relu(yr) if relu else yr
    xd, wd, bd = (t.to(cuda).requires_grad_(True) for t in (x, w, b))
    y = linear_act(xd, wd, bd, relu)
    torch.testing.assert_close(y.cpu(), yr, rtol=1e-5, atol=1e-5)
    if M == 0:
        return
    y.backward(dy.to(cuda))
    # reference gradients in float64, with the ReLU mask taken from the device output: a
    # pre-activation within rounding of 0 may legitimately land on either side of it
    dz = (dy * (y.detach().cpu() > 0)) if relu else dy
    dz64, x64, w64 = dz.double(), x.double(), w.double()
    torch.testing.assert_close(xd.grad.cpu(), (dz64 @ w64).float(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(wd.grad.cpu(), (dz64.t() @ x64).float(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bd.grad.cpu(), dz64.sum(0).float(), rtol=1e-4, atol=1e-4)


def test_linear_strided_input(cuda):
    """x is a column window of a wider matrix (the pooled KeyedTensor)."""
    from two_tower_recommender_model_b200.functional import linear_act
    g = torch.Generator().manual_seed(0)
    big = torch.randn(300, 128, generator=g); w = torch.randn(32, 64, generator=g)
    want = torch.relu(F.linear(big[:, 64:], w))
    got = linear_act(big.to(cuda)[:, 64:], w.to(cuda), None, True)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)


def test_mlp_matches_oracle_and_names(cuda):
    import two_tower_recommender_model_b200 as tt
    mlp = tt.MLP(in_size=64, layer_sizes=[128, 64], device=cuda)
    names = [n for n, _ in mlp.named_parameters()]
    assert names == ["_mlp.0._linear.weight", "_mlp.0._linear.bias", "_mlp.1._linear.weight", "_mlp.1._linear.bias"]
    assert mlp._mlp[-1]._linear.out_features == 64
    x = torch.randn(500, 64)
    layers = [(p._linear.weight.detach().cpu(), p._linear.bias.detach().cpu()) for p in mlp._mlp]
    want = oracle.mlp_forward(x, layers)
    assert (want >= 0).all()  # ReLU after the last layer too
    torch.testing.assert_close(mlp(x.to(cuda)).cpu(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,d", [(1024, 64), (777, 36), (1, 8), (65536, 64)])
def test_dot_bce(cuda, B, d):
    from two_tower_recommender_model_b200.functional import dot_bce_loss
    g = torch.Generator().manual_seed(B)
    q = torch.randn(B, d, generator=g); c = torch.randn(B, d, generator=g); y = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
    qr, cr = q.clone().requires_grad_(True), c.clone().requires_grad_(True)
    logits_r = (qr * cr).sum(dim=1).squeeze()
    loss_r = F.binary_cross_entropy_with_logits(logits_r.reshape(-1), y.float())
    loss_o, logits_o = oracle.dot_bce_loss(q, c, y)
    torch.testing.assert_close(loss_o, loss_r.detach(), rtol=1e-6, atol=1e-7)
    loss_r.backward()
    qd, cd = q.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
    loss, logits = dot_bce_loss(qd, cd, y.to(cuda))
    torch.testing.assert_close(logits.cpu(), logits_r.detach().reshape(-1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(loss.cpu(), loss_r.detach(), rtol=1e-5, atol=1e-6)
    loss.backward()
    torch.testing.assert_close(qd.grad.cpu(), qr.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(cd.grad.cpu(), cr.grad, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("B,d,T", [(256, 64, 1.0), (1000, 64, 0.5), (130, 36, 1.0), (64, 256, 2.0), (2049, 128, 1.0)])
def test_in_batch_softmax(cuda, B, d, T):
    from two_tower_recommender_model_b200.functional import in_batch_softmax_loss
    g = torch.Generator().manual_seed(B + d)
    q = torch.rand(B, d, generator=g); c = torch.rand(B, d, generator=g)
    qr, cr = q.clone().requires_grad_(True), c.clone().requires_grad_(True)
    loss_r = F.cross_entropy((qr @ cr.t()) / T, torch.arange(B))
    loss_o, diag_o = oracle.in_batch_softmax_loss(q, c, T)
    torch.testing.assert_close(loss_o, loss_r.detach(), rtol=1e-5, atol=1e-6)
    loss_r.backward()
    qd, cd = q.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
    loss, diag = in_batch_softmax_loss(qd, cd, T)
    torch.testing.assert_close(loss.cpu(), loss_r.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(diag.cpu(), diag_o, rtol=1e-5, atol=1e-5)
    loss.backward()
    torch.testing.assert_close(qd.grad.cpu(), qr.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(cd.grad.cpu(), cr.grad, rtol=1e-4, atol=1e-7)


def test_flat_adam_matches_torch_adam(cuda):
    import two_tower_recommender_model_b200 as tt
    torch.manual_seed(0)
    ref = [torch.randn(17, 5, requires_grad=True), torch.randn(33, requires_grad=True)]
    mine = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=0.01)
    o_mine = tt.FlatAdam(mine, lr=0.01)
    for step in range(5):
        o_ref.zero_grad(); o_mine.zero_grad()
        gs = [torch.randn_like(p) for p in ref]
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        for p, g in zip(mine, gs):
            p.grad.copy_(g.to(cuda))
        o_ref.step(); o_mine.step()
        # oracle restatement agrees with torch too
    for p, m in zip(ref, mine):
        torch.testing.assert_close(m.detach().cpu(), p.detach(), rtol=1e-5, atol=1e-6)


def test_oracle_adam_restatement():
    torch.manual_seed(1)
    p = torch.randn(40, requires_grad=True); p2 = p.detach().clone()
    m = torch.zeros(40); v = torch.zeros(40)
    opt = torch.optim.Adam([p], lr=0.01)
    for step in range(1, 5):
        g = torch.randn(40)
        p.grad = g.clone(); opt.step()
        oracle.adam_step(p2, g, m, v, step, lr=0.01)
    torch.testing.assert_close(p.detach(), p2, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("Q,N,d,k", [(10, 1000, 32, 100), (130, 5000, 64, 100), (64, 50, 16, 100), (1, 1, 4, 5),
                                     (300, 70000, 64, 100), (5, 300, 7, 128)])
def test_topk_exact_grid(cuda, Q, N, d, k):
    """Entries on a 1/8 grid in [-1,1]: every dot product is exact in fp32 whatever the
    summation order, so indices must match the oracle bit for bit -- including ties."""
    from two_tower_recommender_model_b200.functional import score_topk
    g = torch.Generator().manual_seed(Q * 7 + N)
    q = torch.randint(-8, 9, (Q, d), generator=g).float() / 8
    it = torch.randint(-8, 9, (N, d), generator=g).float() / 8
    ws, wi = oracle.exact_topk(q, it, k)
    s, i = score_topk(q.to(cuda), it.to(cuda), k)
    kk = min(k, N)
    assert torch.equal(i.cpu()[:, :kk], wi)
    assert torch.equal(s.cpu()[:, :kk], ws)
    if kk < k:
        assert (i.cpu()[:, kk:] == -1).all()


def test_topk_random_normal_recall(cuda):
    from two_tower_recommender_model_b200.functional import score_topk
    g = torch.Generator().manual_seed(5)
    q = torch.randn(200, 64, generator=g); it = torch.randn(20000, 64, generator=g)
    ws, wi = oracle.exact_topk(q, it, 100)
    s, i = score_topk(q.to(cuda), it.to(cuda), 100)
    torch.testing.assert_close(s.cpu(), ws, rtol=1e-5, atol=1e-5)
    recall = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i.cpu(), wi)) / wi.numel()
    assert recall >= 0.999
