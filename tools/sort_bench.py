"""Times tt_sort_pairs_u32 and the fused embedding backward at BASELINE configs[1] sizes with CUDA events around a
loop of back-to-back calls (what the launches cost inside a CUDA graph, without per-call event overhead)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from two_tower_recommender_model_b200 import _native as N
from two_tower_recommender_model_b200.functional import sort_pairs

dev = torch.device("cuda:0")
n, bits = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 25
keys = torch.randint(0, 20_000_000, (n,), device=dev, dtype=torch.int32)
vals = torch.arange(n, device=dev, dtype=torch.int32)
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
ws = N.workspace(N.load().tt_sort_pairs_workspace_bytes(n), dev)
sp = N.stream_ptr(dev)


def call():
    N.call("tt_sort_pairs_u32", N.ptr(keys), N.ptr(vals), N.ptr(ko), N.ptr(vo), n, bits, N.ptr(ws), ws.numel(), sp)


for _ in range(5):
    call()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    sp = N.stream_ptr(dev)
    with torch.cuda.graph(g, stream=s):
        for _ in range(20):
            call()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"sort of {n} pairs ({bits}-bit keys): {a.elapsed_time(b) / 100 * 1e3:.1f} us per sort (graph replay of 20 back-to-back sorts)")
k2, v2 = sort_pairs(keys, vals, bits)
assert torch.equal(k2, torch.sort(keys.to(torch.int64) & 0xffffffff, stable=True).values.to(torch.int32))
