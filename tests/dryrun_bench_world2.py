"""Dry run, on CPU with gloo and two processes, of bench.py's N > 1 flow: the sharded-vs-unsharded parity checks (table-wise and
row-wise, eager and through CudaGraphTrainStep), the headline block, every side block (row-wise, weak scaling, global-negatives
refusal recorded as an error entry, sharded retrieval) and the exit path -- with the device entry points and the CUDA runtime
primitives replaced by tests/dryrun_standins.py and the NCCL exchange standing in for the peer-memory one (which needs NVLink
symmetric memory).  What runs here is bench.py's own control flow and arithmetic (the replica's update rule, tolerances, max over
ranks, per-block bookkeeping); the kernels and the peer exchange are what the same run checks on the GPU box.

    python tests/dryrun_bench_world2.py      (spawns 2 ranks; run by tests/test_static_checks.py)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import dryrun_standins as S
    S.install()
    import bench
    from two_tower_recommender_model_b200 import _native as N
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu = torch.device("cpu")
    os.environ["TT_PARITY_BATCH"] = "32"
    tiny = dict(rows=[700, 500], dim=16, layers=[32, 16], batch=64, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001, precision="fp32")
    args = argparse.Namespace(gpus=world, steps=3, warmup=3, impl="ours", no_cpu_baseline=True, no_other_configs=True, exchange="nccl",
                              parity_only=False, parity_graph=True, no_graph=False)
    lib = N.load()
    parity = []
    for sh, gr in (("table_wise", False), ("row_wise", False), ("table_wise", True), ("row_wise", True)):
        p = bench.parity_check(world, rank, cpu, sh, "nccl", steps=4 if gr else 3, precision="fp32", graph=gr)
        assert p["ok"], p
        parity.append(p)
    G = tiny["batch"]
    main = bench.time_block(tiny, G // world, cpu, rank, world, rank, args, "table_wise", "nccl", lib, with_kernels=True)
    assert main["sharding"] == ["table_wise"] and main["cuda_graph"] and main["explain_error"] is None, main
    line = None
    if rank == 0:
        line = bench.headline(args, tiny, main, bench.peaks(), world, G, "dry run, 2 ranks", "strong", parity)
    probe = bench.retrieval_probe_sharded
    bench.retrieval_probe_sharded = lambda dev, r, w: probe(dev, r, w, n_items=3001, q_per_rank=32, d=16, k=10)
    bench.side_blocks(args, tiny, cpu, rank, world, rank, lib, G, line)
    if rank == 0:
        assert line["strong_row_wise"]["sharding"] == ["row_wise"] and line["strong_row_wise"]["value"] > 0, line["strong_row_wise"]
        assert line["weak"]["per_rank_batch"] == G and line["weak"]["global_batch"] == world * G, line["weak"]
        # global negatives run on the tcgen05 kernels only: on the fp32 stand-in path the block must end as an error ENTRY
        assert "error" in line["strong_global_negatives"] or line["strong_global_negatives"]["value"] > 0
        assert line["retrieval"]["queries_total"] == world * 32 and line["retrieval"]["items"] == 3001, line["retrieval"]
        print(json.dumps(line))
        print("bench world-2 dry run ok")
    bench.leave(world)          # flush + os._exit(0): what every rank of a real run ends with


if __name__ == "__main__":
    port = 29100 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(r, 2, port)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    for p in procs:
        if p.is_alive():
            p.terminate()
    sys.exit(0 if all(p.exitcode == 0 for p in procs) else 1)
