"""tcgen05 / TMA / TMEM kernels vs float64 math on the SAME bf16-rounded operands.
Tolerance: fp32 accumulation of exact bf16 products -> rtol 1e-4, atol 1e-4 * sqrt(K)-ish;
bf16 outputs add one bf16 rounding (rtol 8e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g) / K ** 0.5
    return a, b


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 128), (256, 256, 64), (1000, 100, 80), (4096, 128, 64),
                                   (333, 64, 1024), (65536, 64, 128), (128, 512, 256), (64, 16, 16), (129, 40, 72)])
def test_gemm_bf16_plain(cuda, M, N, K):
    from two_tower_recommender_model_b200.functional import cast_bf16, gemm_bf16
    a, b = _ops(M, N, K, M + N + K)
    ad, bd = cast_bf16(a.to(cuda)), cast_bf16(b.to(cuda))
    assert torch.equal(ad.cpu(), a.bfloat16()) and torch.equal(bd.cpu(), b.bfloat16())
    want = ad.double().cpu() @ bd.double().cpu().t()
    got = gemm_bf16(ad, bd)["f32"]
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-4, atol=1e-4)


def test_cast_transposed(cuda):
    from two_tower_recommender_model_b200.functional import cast_bf16
    x = torch.randn(1000, 72)
    r, t = cast_bf16(x.to(cuda), both=True)
    assert torch.equal(r.cpu(), x.bfloat16()) and torch.equal(t.cpu(), x.bfloat16().t())
    win = cast_bf16(x.to(cuda)[:, 8:40], transposed=True)
    assert torch.equal(win.cpu(), x[:, 8:40].bfloat16().t())


@pytest.mark.parametrize("M,N,K", [(1024, 128, 64), (777, 64, 128), (2048, 256, 512)])
def test_gemm_bf16_epilogues(cuda, M, N, K):
    from two_tower_recommender_model_b200.functional import cast_bf16, gemm_bf16
    a, b = _ops(M, N, K, 7 + M)
    g = torch.Generator().manual_seed(1)
    bias = torch.randn(N, generator=g)
    mask = torch.randn(M, N, generator=g)
    ad, bd = cast_bf16(a.to(cuda)), cast_bf16(b.to(cuda))
    base = ad.double().cpu() @ bd.double().cpu().t()
    r = gemm_bf16(ad, bd, bias=bias.to(cuda), relu=True, out_f32=True, out_bf16=True, out_bf16_t=True)
    want = torch.relu(base + bias.double())
    torch.testing.assert_close(r["f32"].cpu().double(), want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["bf16"].cpu().double(), want, rtol=8e-3, atol=1e-3)
    torch.testing.assert_close(r["bf16_t"].cpu().double(), want.t(), rtol=8e-3, atol=1e-3)
    r = gemm_bf16(ad, bd, mask=mask.to(cuda))
    torch.testing.assert_close(r["f32"].cpu().double(), base * (mask > 0), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,d,T", [(128, 64, 1.0), (256, 64, 1.0), (1000, 64, 0.5), (4096, 64, 1.0), (777, 128, 1.0),
                                   (300, 36, 2.0), (513, 256, 1.0), (640, 192, 1.0), (65, 8, 1.0), (2000, 48, 1.0),
                                   (8192, 64, 0.25), (3001, 20, 1.0), (200, 30, 1.0),
                                   (8192, 256, 1.0), (4100, 200, 0.5), (8200, 128, 1.0), (2050, 100, 1.0), (129, 256, 1.0),   # cfg4's width
                                   (1000, 64, 0.01), (4096, 64, 0.03)])   # small T: logits up to 400 -> the online-max forward path
def test_in_batch_softmax_tensor_core(cuda, B, d, T):
    """bf16 tensor-core softmax vs float64 math on the bf16-rounded q, c.
    Tolerance: loss/lse rtol 1e-4 (fp32 accumulate).  Gradients: dq = (P c - c_b)/(B T) is a
    difference of two O(|c|) terms and P is rounded to bf16 (2^-9 relative) before the second
    GEMM, so the error scales with the TERMS: atol = 6e-3 * max|c| / (B T), rtol 2e-2."""
    from two_tower_recommender_model_b200.functional import in_batch_softmax_loss
    g = torch.Generator().manual_seed(B + d)
    q = torch.rand(B, d, generator=g) * (2.0 / d ** 0.5)
    c = torch.rand(B, d, generator=g) * (2.0 / d ** 0.5)
    q[::7] = 0  # ReLU outputs: exact zeros
    qr = q.bfloat16().double().requires_grad_(True)
    cr = c.bfloat16().double().requires_grad_(True)
    s = (qr @ cr.t()) / T
    loss_r = torch.nn.functional.cross_entropy(s, torch.arange(B))
    loss_r.backward()
    qd, cd = q.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
    loss, diag = in_batch_softmax_loss(qd, cd, T, precision="bf16")
    torch.testing.assert_close(loss.cpu().double(), loss_r.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(diag.cpu().double(), s.detach().diagonal(), rtol=1e-4, atol=1e-5)
    loss.backward()
    atol = 6e-3 * float(max(q.abs().max(), c.abs().max())) / (B * T)
    torch.testing.assert_close(qd.grad.cpu().double(), qr.grad, rtol=2e-2, atol=atol)
    torch.testing.assert_close(cd.grad.cpu().double(), cr.grad, rtol=2e-2, atol=atol)


def test_in_batch_softmax_tc_matches_fp32_path(cuda):
    from two_tower_recommender_model_b200.functional import in_batch_softmax_loss
    g = torch.Generator().manual_seed(3)
    B, d = 2048, 64
    q = torch.rand(B, d, generator=g) * 0.3
    c = torch.rand(B, d, generator=g) * 0.3
    a = in_batch_softmax_loss(q.to(cuda), c.to(cuda), 1.0, precision="fp32")[0]
    b = in_batch_softmax_loss(q.to(cuda), c.to(cuda), 1.0, precision="bf16")[0]
    torch.testing.assert_close(a, b, rtol=1e-2, atol=1e-3)  # bf16 operand rounding


@pytest.mark.parametrize("M,N,K", [(64, 128, 65536), (128, 64, 65536), (100, 36, 5000), (1024, 512, 4096)])
def test_gemm_splitk_and_colsum(cuda, M, N, K):
    from two_tower_recommender_model_b200.functional import cast_bf16, colsum_bf16, gemm_bf16_splitk
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g); b = torch.randn(N, K, generator=g)
    ad, bd = cast_bf16(a.to(cuda)), cast_bf16(b.to(cuda))
    want = ad.double().cpu() @ bd.double().cpu().t()
    got = gemm_bf16_splitk(ad, bd)
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-4, atol=2e-3)
    cs = colsum_bf16(cast_bf16(a.t().contiguous().to(cuda)))       # [K, M] -> sums over K
    torch.testing.assert_close(cs.cpu().double(), a.bfloat16().double().sum(1), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(64, 128, 65536), (1024, 512, 4096), (256, 512, 5000), (20, 36, 777), (128, 1024, 8200), (8, 64, 130)])
def test_gemm_splitk_mn_major(cuda, M, N, K):
    """dW = dZ^T A from ROW-MAJOR dZ [K, M] and A [K, N] (both tcgen05 operands MN-major, no transposed copies) against
    float64 on the same bf16-rounded operands, and against the K-major split-K kernel fed the transposed copies."""
    from two_tower_recommender_model_b200.functional import cast_bf16, gemm_bf16_splitk, gemm_bf16_splitk_mn
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(K, M, generator=g); b = torch.randn(K, N, generator=g)
    ad, bd = cast_bf16(a.to(cuda)), cast_bf16(b.to(cuda))
    want = ad.double().cpu().t() @ bd.double().cpu()
    got = gemm_bf16_splitk_mn(ad, bd)
    torch.testing.assert_close(got.cpu().double(), want, rtol=1e-4, atol=2e-3 * (K / 4096) ** 0.5)
    ref = gemm_bf16_splitk(cast_bf16(a.to(cuda), transposed=True), cast_bf16(b.to(cuda), transposed=True))
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-3 * (K / 4096) ** 0.5)


def test_cast_gate(cuda):
    from two_tower_recommender_model_b200.functional import cast_bf16
    x = torch.randn(300, 40); gt = torch.randn(300, 40)
    r, t = cast_bf16(x.to(cuda), both=True, gate=gt.to(cuda))
    want = (x * (gt > 0)).bfloat16()
    assert torch.equal(r.cpu(), want) and torch.equal(t.cpu(), want.t())


@pytest.mark.parametrize("B,sizes", [(4096, [64, 128, 64]), (1000, [128, 1024, 512, 256]), (777, [36, 20, 8]), (65536, [64, 128, 64])])
def test_mlp_tensor_core(cuda, B, sizes):
    """bf16 tower (tcgen05) vs a float64 emulation of the SAME pipeline (operands and stored
    activations / dZ rounded to bf16 at the same points; ReLU gates taken from the bf16
    activations).  A plain fp32 tower is not a usable reference for gradients: a hidden unit
    whose pre-activation is within bf16 rounding of 0 legitimately gates differently.
    Tolerance: relative Frobenius error < 1e-2 (outputs 2e-3)."""
    import two_tower_recommender_model_b200 as tt
    torch.manual_seed(B)
    tc = tt.MLP(sizes[0], sizes[1:], device=cuda, precision="bf16")
    x = torch.randn(B, sizes[0])
    dy = torch.randn(B, sizes[-1])
    bf = lambda t: t.float().bfloat16().double()
    Ws = [p._linear.weight.detach().cpu() for p in tc._mlp]
    bs = [p._linear.bias.detach().cpu().double() for p in tc._mlp]
    acts = [bf(x)]
    pre_last = None
    for l, (W, b) in enumerate(zip(Ws, bs)):
        h = torch.relu(acts[-1] @ bf(W).t() + b)
        pre_last = h
        acts.append(bf(h))
    y_ref = pre_last                                   # last layer output is returned in fp32
    dz = bf(dy.double() * (y_ref > 0))
    dWs, dbs = [None] * len(Ws), [None] * len(Ws)
    for l in range(len(Ws) - 1, -1, -1):
        dWs[l] = dz.t() @ acts[l]
        dbs[l] = dz.sum(0)
        da = dz @ bf(Ws[l])
        if l > 0:
            dz = bf(da * (acts[l] > 0))
        else:
            dx_ref = da
    xt = x.to(cuda).requires_grad_(True)
    yt = tc(xt)
    rel = lambda a, b: float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))
    assert rel(yt.detach(), y_ref) < 2e-3
    yt.backward(dy.to(cuda))
    assert rel(xt.grad, dx_ref) < 1e-2, rel(xt.grad, dx_ref)
    for l, p in enumerate(tc._mlp):
        assert rel(p._linear.weight.grad, dWs[l]) < 1e-2, (l, rel(p._linear.weight.grad, dWs[l]))
        assert rel(p._linear.bias.grad, dbs[l]) < 1e-2, (l, rel(p._linear.bias.grad, dbs[l]))
    # and it stays close to the fp32 tower in the forward direction
    ref = tt.MLP(sizes[0], sizes[1:], device=cuda, precision="fp32")
    ref.load_state_dict(tc.state_dict())
    yr = ref(x.to(cuda))
    torch.testing.assert_close(yt.detach(), yr, rtol=3e-2, atol=3e-2 * float(yr.detach().abs().max()))


@pytest.mark.parametrize("Q,N,d,k", [(10, 1000, 32, 100), (130, 5000, 64, 100), (64, 50, 16, 100), (1, 1, 8, 5),
                                     (300, 70000, 64, 100), (5, 300, 24, 128), (1000, 200000, 64, 10)])
def test_topk_tensor_core_exact_grid(cuda, Q, N, d, k):
    """tcgen05 scoring + fused top-k on exact-arithmetic vectors (entries k/8 in [-1,1] are exact in
    bf16 and every dot product is exact in fp32): indices and scores must match the oracle bit for bit."""
    import oracle
    from two_tower_recommender_model_b200.functional import score_topk
    g = torch.Generator().manual_seed(Q * 7 + N)
    q = torch.randint(-8, 9, (Q, d), generator=g).float() / 8
    it = torch.randint(-8, 9, (N, d), generator=g).float() / 8
    ws, wi = oracle.exact_topk(q, it, k)
    s, i = score_topk(q.to(cuda), it.to(cuda), k, precision="bf16")
    kk = min(k, N)
    assert torch.equal(i.cpu()[:, :kk], wi)
    assert torch.equal(s.cpu()[:, :kk], ws)
    if kk < k:
        assert (i.cpu()[:, kk:] == -1).all()


def test_topk_tensor_core_random_normal(cuda):
    """Random-normal vectors: compare with the float64 oracle on the SAME bf16-rounded inputs:
    scores rtol 1e-5, recall@100 >= 0.999 (near-ties may swap under a different summation order)."""
    import oracle
    from two_tower_recommender_model_b200.functional import score_topk
    g = torch.Generator().manual_seed(5)
    q = torch.randn(256, 64, generator=g); it = torch.randn(50000, 64, generator=g)
    ws, wi = oracle.exact_topk(q.bfloat16().float(), it.bfloat16().float(), 100)
    s, i = score_topk(q.to(cuda), it.to(cuda), 100, precision="bf16")
    torch.testing.assert_close(s.cpu(), ws, rtol=1e-5, atol=1e-5)
    recall = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i.cpu(), wi)) / wi.numel()
    assert recall >= 0.999


@pytest.mark.parametrize("B,i,h,o", [(1000, 64, 128, 64), (4096, 32, 64, 32), (130, 8, 8, 8), (65536, 64, 128, 64), (777, 48, 96, 24)])
def test_fused_towers_match_per_layer_path(cuda, B, i, h, o):
    """One-launch fused towers (tt_towers_forward_fused / tt_towers_backward_fused) against the per-layer
    tcgen05 path (MlpTC: cast + tt_gemm_bf16 + split-K + colsum) on the same inputs.  Same bf16 operand
    rounding and fp32 accumulation in both, so activations agree to fp32 round-off (a bf16 re-rounding flip
    of h can move an output by ~2^-9 of one term: rtol 2e-3); gradients are long sums in different orders."""
    from two_tower_recommender_model_b200.functional import FusedTowersTC, MlpTC
    g = torch.Generator().manual_seed(B + i)
    pooled = (torch.randn(B, 2 * i, generator=g) * 0.5).to(cuda)
    params = []
    for t in range(2):
        params += [(torch.randn(h, i, generator=g) / i ** 0.5).to(cuda), (torch.randn(h, generator=g) * 0.1).to(cuda),
                   (torch.randn(o, h, generator=g) / h ** 0.5).to(cuda), (torch.randn(o, generator=g) * 0.1).to(cuda)]
    dys = [torch.randn(B, o, generator=g).to(cuda) for _ in range(2)]

    def run(fused):
        p = pooled.clone().requires_grad_(True)
        ps = [x.clone().requires_grad_(True) for x in params]
        if fused:
            ys = FusedTowersTC.apply(p, (0, i), i, None, *ps)[:2]
        else:
            ys = [MlpTC.apply(p.narrow(1, t * i, i), *ps[4 * t: 4 * t + 4]) for t in range(2)]
        torch.autograd.backward(list(ys), dys)
        return [y.detach() for y in ys], p.grad, [x.grad for x in ps]

    ya, dpa, ga = run(True)
    yb, dpb, gb = run(False)
    for a, b in zip(ya, yb):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(dpa, dpb, rtol=2e-2, atol=2e-2 * float(dpb.abs().max()))
    for k, (a, b) in enumerate(zip(ga, gb)):
        torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2 * float(b.abs().max()), msg=lambda m: f"param {k}: {m}")


@pytest.mark.parametrize("B,d", [(3000, 256), (1025, 136), (2048, 192)])
def test_in_batch_softmax_wide_kernels_match_ss_form(cuda, B, d):
    """64 < d <= 256: the TS-form backward (X resident in TMEM, streamed tile reused as MN-major B) against the SS-form
    kernels of round 1 on the same inputs.  Same math, same bf16 rounding of P, different accumulation order only."""
    from two_tower_recommender_model_b200 import functional as F
    g = torch.Generator().manual_seed(B)
    q = (torch.rand(B, d, generator=g) * (2.0 / d ** 0.5)).to(cuda)
    c = (torch.rand(B, d, generator=g) * (2.0 / d ** 0.5)).to(cuda)
    out = {}
    try:
        for wide in (True, False):
            F.set_softmax_wide(wide)
            qd, cd = q.clone().requires_grad_(True), c.clone().requires_grad_(True)
            loss, _ = F.in_batch_softmax_loss(qd, cd, 1.0, precision="bf16")
            loss.backward()
            out[wide] = (loss.detach(), qd.grad, cd.grad)
    finally:
        F.set_softmax_wide(True)
    assert torch.equal(out[True][0], out[False][0])                      # same forward kernel
    for a, b in zip(out[True][1:], out[False][1:]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))


@pytest.mark.parametrize("B,d", [(65536, 64), (32768, 256)])
def test_in_batch_softmax_full_size_closed_form(cuda, B, d):
    """BASELINE configs[1] size (B = 65536, d = 64) and cfg4's width (d = 256), where an O(B^2) reference is out of reach: with every candidate
    equal to one vector c0 the logits of a row are constant, so P = 1/B exactly and everything has a closed form:
        loss = log B,   dq = 0,   dc_j = (mean_i q_i - q_j) / (B T).
    Exercises forward + the one-pass backward (P.c, P^T.q, TMA reduce-add over 256 row blocks) at full size."""
    import math
    from two_tower_recommender_model_b200.functional import in_batch_softmax_loss
    T = 0.5
    g = torch.Generator().manual_seed(1)
    q = (torch.rand(B, d, generator=g) * 0.25).bfloat16().float()        # exactly representable operands
    c0 = (torch.rand(d, generator=g) * (16.0 / d)).bfloat16().float()
    c = c0.repeat(B, 1)
    qd, cd = q.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
    loss, diag = in_batch_softmax_loss(qd, cd, T, precision="bf16")
    loss.backward()
    assert abs(float(loss) - math.log(B)) < 1e-4 * math.log(B)
    torch.testing.assert_close(diag.cpu().double(), (q.double() @ c0.double()) / T, rtol=1e-5, atol=1e-6)
    scale = 1.0 / (B * T)
    # P is rounded to bf16 before the gradient products (1/65536 is exact), so the only error is fp32 accumulation
    assert float(qd.grad.abs().max()) <= 2e-3 * scale * float(c0.abs().max())
    want_dc = (q.double().mean(0, keepdim=True) - q.double()) * scale
    torch.testing.assert_close(cd.grad.cpu().double(), want_dc, rtol=1e-3, atol=2e-3 * scale * float(q.abs().max()))


def test_fused_towers_without_bias_and_unused_tower(cuda):
    """bias=False towers (null b1 / b2 pointers, no bias gradients) and a tower whose output receives no gradient."""
    from two_tower_recommender_model_b200.functional import FusedTowersTC, MlpTC
    g = torch.Generator().manual_seed(4)
    B, i, h, o = 515, 64, 128, 64
    pooled = (torch.randn(B, 2 * i, generator=g) * 0.5).to(cuda)
    ws = [(torch.randn(h, i, generator=g) / i ** 0.5).to(cuda), None, (torch.randn(o, h, generator=g) / h ** 0.5).to(cuda), None] * 2
    dy0 = torch.randn(B, o, generator=g).to(cuda)

    def run(fused):
        p = pooled.clone().requires_grad_(True)
        ps = [None if x is None else x.clone().requires_grad_(True) for x in ws]
        if fused:
            ys = FusedTowersTC.apply(p, (0, i), i, None, *ps)[:2]
        else:
            ys = [MlpTC.apply(p.narrow(1, t * i, i), *ps[4 * t: 4 * t + 4]) for t in range(2)]
        ys[0].backward(dy0)                      # tower 1 gets no gradient at all
        return ys[0].detach(), ys[1].detach(), p.grad, [None if x is None else x.grad for x in ps]

    a, b = run(True), run(False)
    torch.testing.assert_close(a[0], b[0], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(a[1], b[1], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(a[2][:, :i], b[2][:, :i], rtol=2e-2, atol=2e-2 * float(b[2].abs().max()))
    assert float(a[2][:, i:].abs().max()) == 0.0
    for k in (0, 2):
        torch.testing.assert_close(a[3][k], b[3][k], rtol=2e-2, atol=2e-2 * float(b[3][k].abs().max()))
    for k in (4, 6):                              # the unused tower: zero weight gradients (or none on the per-layer path)
        assert a[3][k] is None or float(a[3][k].abs().max()) == 0.0


@pytest.mark.parametrize("B,i,h,o", [(1000, 64, 128, 64), (8192, 64, 128, 64), (777, 48, 96, 24)])
def test_fused_towers_against_float64_emulation(cuda, B, i, h, o):
    """tt_towers_forward_fused / tt_towers_backward_fused against a float64 emulation of the SAME pipeline
    (x, h, dz2, dz1 rounded to bf16 where the kernels round them; ReLU gates from the rounded activations; fp32 y):
    outputs, dx, dW1, dW2, db1, db2 of both towers.  This pins the fused kernels to an oracle, not to the
    repo's own per-layer path.  Tolerance: relative Frobenius error 2e-3 on y, 1e-2 on gradients."""
    from two_tower_recommender_model_b200.functional import FusedTowersTC
    g = torch.Generator().manual_seed(3 * B + i)
    pooled = torch.randn(B, 2 * i, generator=g) * 0.5
    params = []
    for t in range(2):
        params += [torch.randn(h, i, generator=g) / i ** 0.5, torch.randn(h, generator=g) * 0.1,
                   torch.randn(o, h, generator=g) / h ** 0.5, torch.randn(o, generator=g) * 0.1]
    dys = [torch.randn(B, o, generator=g) for _ in range(2)]
    bf = lambda t: t.float().bfloat16().double()
    p = pooled.to(cuda).requires_grad_(True)
    ps = [x.to(cuda).requires_grad_(True) for x in params]
    ys = FusedTowersTC.apply(p, (0, i), i, None, *ps)[:2]
    torch.autograd.backward(list(ys), [d.to(cuda) for d in dys])
    rel = lambda a, b: float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))
    for t in range(2):
        W1, b1, W2, b2 = params[4 * t: 4 * t + 4]
        x = bf(pooled[:, t * i:(t + 1) * i])
        hh = bf(torch.relu(x @ bf(W1).t() + b1.double()))
        y = torch.relu(hh @ bf(W2).t() + b2.double())
        dz2 = bf(dys[t].double() * (y > 0))
        dW2, db2 = dz2.t() @ hh, dz2.sum(0)
        dz1 = bf((dz2 @ bf(W2)) * (hh > 0))
        dW1, db1 = dz1.t() @ x, dz1.sum(0)
        dx = dz1 @ bf(W1)
        assert rel(ys[t].detach(), y) < 2e-3, (t, rel(ys[t].detach(), y))
        assert rel(p.grad[:, t * i:(t + 1) * i], dx) < 1e-2, (t, "dx", rel(p.grad[:, t * i:(t + 1) * i], dx))
        for name, got, want in (("dW1", ps[4 * t].grad, dW1), ("db1", ps[4 * t + 1].grad, db1),
                                ("dW2", ps[4 * t + 2].grad, dW2), ("db2", ps[4 * t + 3].grad, db2)):
            assert rel(got, want) < 1e-2, (t, name, rel(got, want))
