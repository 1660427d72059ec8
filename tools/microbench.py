"""Per-kernel micro-benchmarks on one GPU (CUDA events, warm-up, L2-sized inputs noted).
    python tools/microbench.py softmax|ebc|towers|topk [--B 65536] [--d 64] [--iters 10]
Used for ncu captures (`ncu -k regex:... python tools/microbench.py softmax --iters 1`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import two_tower_recommender_model_b200 as tt  # noqa: E402
from two_tower_recommender_model_b200 import _native as N  # noqa: E402
from two_tower_recommender_model_b200 import functional as F  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if N._timing is not None:
        N.enable_timing(True)   # drop the warm-up calls (first-call attribute / module-load costs)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def softmax(args):
    dev = torch.device("cuda:0")
    B, d = args.B, args.d
    q = (torch.rand(B, d, device=dev) * 0.3).requires_grad_(True)
    c = (torch.rand(B, d, device=dev) * 0.3).requires_grad_(True)
    N.enable_timing(True)

    def run():
        loss, _ = F.in_batch_softmax_loss(q, c, 1.0, precision=args.precision)
        loss.backward()
    ms = timed(run, args.iters, warmup=args.warmup)
    torch.cuda.synchronize()
    summ = N.timing_summary()
    flops_f = 2.0 * B * B * d
    for k, v in summ.items():
        fl = flops_f if "forward" in k else (2 * flops_f * 2 if "backward" in k else 0)
        print(f"{k:42s} {v['ms']:9.4f} ms  x{v['calls']}" + (f"  {fl / v['ms'] / 1e9:8.1f} TFLOP/s (executed)" if fl else ""))
    print(f"fwd+bwd wall {ms:.4f} ms -> credited 6*B*B*d = {6.0 * B * B * d / ms / 1e9:.1f} TFLOP/s")


def ebc(args):
    dev = torch.device("cuda:0")
    from torch.distributed.optim import _apply_optimizer_in_backward as aob
    B, D, R, L = args.B, args.d, args.rows, args.L
    cfgs = [tt.EmbeddingBagConfig(name=f"t{i}", embedding_dim=D, num_embeddings=R, feature_names=[f"f{i}"],
                                  pooling=tt.PoolingType.MEAN if L > 1 else tt.PoolingType.SUM) for i in range(2)]
    e = tt.EmbeddingBagCollection(tables=cfgs, device=dev)
    aob(tt.RowWiseAdagrad, e.parameters(), {"lr": 0.01})
    lens = torch.full((2 * B,), L, dtype=torch.int32, device=dev)
    def draw():
        if args.zipf > 0:   # Zipf(alpha) ids expose hot-row handling in the fused update
            r = torch.arange(1, min(R, 1 << 22) + 1, device=dev, dtype=torch.float64)
            pmf = r.pow(-args.zipf)
            return torch.multinomial((pmf / pmf.sum()).float(), 2 * B * L, replacement=True)
        return torch.randint(0, R, (2 * B * L,), device=dev)
    kjts = [tt.KeyedJaggedTensor.from_lengths_sync(["f0", "f1"], draw(), lens) for _ in range(4)]
    go = torch.randn(B, 2 * D, device=dev)
    N.enable_timing(True)
    i = [0]

    def run():
        kt = e(kjts[i[0] % 4]); i[0] += 1
        kt.values().backward(go)
    timed(run, args.iters, warmup=args.warmup)
    torch.cuda.synchronize()
    per = dict(N.timing_summary())
    N.enable_timing(False)
    # the lookup alone, back to back through the C ABI (what bench.py reports as ebc_lookup)
    from ctypes import byref
    plan, total_dim = e._build_plan(tuple(kjts[0].keys()), B, with_state=False)
    vals = [k.values().contiguous() for k in kjts]
    offs = [k.offsets().to(torch.int32).contiguous() for k in kjts]
    pooled = torch.empty(B, total_dim, dtype=torch.float32, device=dev)
    sp = N.stream_ptr(dev)
    b2b = timed(lambda: [N.call("tt_ebc_forward", byref(plan), N.ptr(vals[j % 4]), N.ptr(offs[j % 4]), N.ptr(pooled), sp)
                         for j in range(40)], 5, warmup=2) / 40
    uniq = sum(int(torch.unique(k[f].values()).numel()) for k in kjts[:1] for f in ("f0", "f1"))
    fwd_bytes = 2 * (B * L * (8 + 4 * D) + 4 * B + 4 * B * D)
    bwd_bytes = 2 * (4 * B * D + 8 * B * L) + uniq * (8 * D + 8)
    for k, v in per.items():
        by = fwd_bytes if "forward" in k else bwd_bytes
        print(f"{k:42s} {v['ms'] * 1e3:9.1f} us  x{v['calls']}  {by / v['ms'] / 1e6:8.1f} GB/s algorithmic ({by / 1e6:.1f} MB; unique rows {uniq})")
    print(f"tt_ebc_forward back to back                {b2b * 1e3:9.2f} us  {fwd_bytes / b2b / 1e6:8.1f} GB/s algorithmic")


def towers(args):
    """Both towers (64 -> 128 -> 64), forward + backward.  precision bf16: the one-launch fused kernels
    (bytes per sample: fwd 4*64 + 2*256 + 4*64 + (bf16 y) = 1152 B... see towers_tcgen05.cu); fp32: CUDA-core path."""
    dev = torch.device("cuda:0")
    B = args.B
    N.enable_timing(True)
    if args.precision == "bf16":
        mlps = [tt.MLP(64, [128, 64], device=dev, precision="bf16") for _ in range(2)]
        params = []
        for m in mlps:
            for p in m._mlp:
                params += [p._linear.weight, p._linear.bias]
        pooled = torch.randn(B, 128, device=dev, requires_grad=True)
        dys = [torch.randn(B, 64, device=dev) for _ in range(2)]

        def run():
            ys = F.FusedTowersTC.apply(pooled, (0, 64), 64, None, *params)[:2]
            torch.autograd.backward(list(ys), dys)
        fwd_bytes = 2 * B * (4 * 64 + 2 * 64 + 2 * 128 + 4 * 64 + 2 * 64)
        bwd_bytes = 2 * B * (4 * 64 + 2 * 64 + 2 * 128 + 2 * 64 + 4 * 64)
    else:
        mlp = tt.MLP(64, [128, 64], device=dev)
        x = torch.randn(B, 64, device=dev, requires_grad=True)

        def run():
            mlp(x).sum().backward()
        fwd_bytes = bwd_bytes = 0
    ms = timed(run, args.iters, warmup=args.warmup)
    for k, v in N.timing_summary().items():
        by = fwd_bytes if "forward_fused" in k else (bwd_bytes if "backward_fused" in k else 0)
        print(f"{k:42s} {v['ms'] * 1e3:9.1f} us  x{v['calls']}" + (f"  {by / v['ms'] / 1e6:8.1f} GB/s algorithmic" if by else ""))
    print(f"towers fwd+bwd {ms:.3f} ms")


def mlp(args):
    """One tower of arbitrary widths through the per-layer tcgen05 GEMM path (cfg4: --din 128 --layers 1024,512,256),
    forward + backward, per-call times and the TFLOP/s of the GEMM calls (2*M*N*K each)."""
    dev = torch.device("cuda:0")
    B = args.B
    layers = [int(x) for x in args.layers.split(",")]
    m = tt.MLP(args.din, layers, device=dev, precision=args.precision)
    x = torch.randn(B, args.din, device=dev, requires_grad=True)
    dy = torch.randn(B, layers[-1], device=dev)
    N.enable_timing(True)

    def run():
        m(x).backward(dy)
    ms = timed(run, args.iters, warmup=args.warmup)
    dims = [args.din] + layers
    flops = sum(2.0 * B * a * b for a, b in zip(dims[:-1], dims[1:]))
    gemm_ms = 0.0
    for k, v in N.timing_summary().items():
        print(f"{k:42s} {v['ms'] * 1e3:9.1f} us  x{v['calls']}")
        if "gemm" in k:
            gemm_ms += v["ms"] * v["calls"] / max(args.iters, 1)
    print(f"mlp {dims} B={B}: fwd+bwd {ms:.3f} ms = {3 * flops / ms / 1e9:.1f} TFLOP/s over the whole call chain (3 GEMMs per layer); "
          f"GEMM calls alone {gemm_ms:.3f} ms = {3 * flops / max(gemm_ms, 1e-9) / 1e9:.1f} TFLOP/s")


def topk(args):
    dev = torch.device("cuda:0")
    q = torch.randn(args.Q, args.d, device=dev)
    it = torch.randn(args.rows, args.d, device=dev)
    itb = F.cast_bf16(it) if args.precision == "bf16" else None
    ms = timed(lambda: F.score_topk(q, it, args.k, precision=args.precision, items_bf16=itb), args.iters, warmup=1)
    print(f"top-{args.k} of {args.rows} items for {args.Q} queries: {ms:.3f} ms, {2.0 * args.Q * args.rows * args.d / ms / 1e9:.2f} TFLOP/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["softmax", "ebc", "towers", "mlp", "topk"])
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--L", type=int, default=1)
    ap.add_argument("--din", type=int, default=128)
    ap.add_argument("--layers", default="1024,512,256")
    ap.add_argument("--Q", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--zipf", type=float, default=0.0, help="ebc: draw ids from Zipf(alpha) instead of uniform")
    a = ap.parse_args()
    {"softmax": softmax, "ebc": ebc, "towers": towers, "mlp": mlp, "topk": topk}[a.what](a)
