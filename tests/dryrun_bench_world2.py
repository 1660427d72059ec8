"""Dry run, on CPU with gloo and two processes, of bench.py's N > 1 run: ``run_ours(args)`` itself on every rank -- the
sharded-vs-unsharded parity checks (table-wise and row-wise, bf16-configured and fp32, eager and through CudaGraphTrainStep),
the headline block, every side block (row-wise, weak scaling, global negatives, sharded retrieval) and the exit path -- with
the device entry points and the CUDA runtime primitives replaced by tests/dryrun_standins.py and the NCCL exchange
(``--exchange nccl``) standing in for the peer-memory one, which needs NVLink symmetric memory.  What runs here is bench.py's
own control flow and arithmetic (the replica's update rule, tolerances, max over ranks, per-block bookkeeping); the kernels and
the peer exchange are what the same run checks on the GPU box.

    python tests/dryrun_bench_world2.py      (spawns 2 ranks; run by tests/test_static_checks.py)
"""
import argparse
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      TT_PARITY_BATCH="32")
    import dryrun_standins as S
    S.install()
    import bench
    tiny = dict(rows=[700, 500], dim=16, layers=[32, 16], batch=64, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001)
    S.run_bench_on_cpu(bench, tiny)
    args = argparse.Namespace(gpus=world, steps=3, warmup=3, impl="ours", no_cpu_baseline=True, no_other_configs=False, exchange="nccl",
                              parity_only=False, parity_graph=True, no_graph=False)
    if os.environ.get("DRYRUN_CFG3_SHARDED"):
        # configs[2] row-wise sharded (the last side block of an N > 1 run), called directly: tiny tables, fp32 towers, the NCCL
        # exchange (the peer-memory one needs NVLink symmetric memory)
        import torch
        import torch.distributed as dist
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_configs
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = run_configs.config3_sharded(torch.device("cpu"), rank, world, steps=3, warmup=2, batch=64, big_rows=5000, D=16, L=4,
                                          layers=(32, 16), precision="fp32", peer_exchange=False)
        assert out["sharding"] == ["row_wise"] and out["cuda_graph"] and out["per_rank_batch"] == 32 and out["value"] > 0, out
        assert out["ids_per_rank_per_step"] == 32 * 4 + 3 * 32 and 0 < out["last_loss"] < 20, out
        # configs[3] under the planner's sharding (the block before it): tiny tables, fp32 towers of three layers, row-wise Adam
        out4 = run_configs.config4_sharded(torch.device("cpu"), rank, world, steps=3, warmup=2, batch=64, rows=(900, 700), D=16,
                                           layers=(32, 24, 16), precision="fp32", peer_exchange=False)
        assert out4["sharding"] == ["table_wise"] and out4["cuda_graph"] and out4["per_rank_batch"] == 32 and out4["value"] > 0, out4
        assert 0 < out4["last_loss"] < 20 and out4["per_rank_tflops_credited"] >= 0, out4
        if rank == 0:
            import json
            print(json.dumps(out))
            print(json.dumps(out4))
        bench.leave(world)
    if os.environ.get("DRYRUN_FAIL_ROW_WISE"):
        # what a failed row-wise parity check does to the run: the table-wise headline stands, row-wise blocks are skipped
        # (the four eager parity modes and the configs[1] blocks are enough for that)
        args.parity_graph, args.no_other_configs = False, True
        real = bench.parity_check

        def parity_check(world, rank, dev, sharding, exchange, **kw):
            p = real(world, rank, dev, sharding, exchange, **kw)
            if sharding == "row_wise":
                p["ok"] = False
            return p
        bench.parity_check = parity_check
    bench.run_ours(args)        # every rank leaves through bench.leave(): flush + os._exit(0)
    raise AssertionError("run_ours returned on a multi-rank run")


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
    ok = all(p.exitcode == 0 for p in procs)
    print("bench world-2 dry run ok" if ok else f"exit codes {[p.exitcode for p in procs]}")
    sys.exit(0 if ok else 1)
