"""Dry run, on CPU with gloo and two processes, of the 2-GPU training worker of tests/test_gpu_multi.py (`_worker`: three
training steps of the sharded model through TrainPipelineSparseDist, losses and gathered tables against the oracle's two-rank
step) in the modes that have NOT run on GPUs yet -- column_wise, data_parallel, data_parallel_dense -- and, as a control, in
two that have (table_wise, row_wise).  Device entry points -> tests/dryrun_standins.py; `torch.device("cuda", r)` -> CPU;
NCCL -> gloo.  What this checks is the worker's flow and ITS NUMBERS against the oracle with the oracle-backed lookup in place
of the kernels, i.e. the host side of the sharded step; the kernels are what the same test checks on a 2-GPU box.

    python tests/dryrun_multi.py       (spawns 2 ranks; run by tests/test_static_checks.py)
"""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
MODES = ["table_wise", "row_wise", "column_wise", "data_parallel", "data_parallel_dense"]


class _Queue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def worker(rank, world, port):
    import torch.distributed as dist
    import dryrun_standins as S
    S.install()
    import test_gpu_multi as T
    T.torch = S._TorchProxy()
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo", rank=kw.get("rank"), world_size=kw.get("world_size"))
    real_exit = os._exit
    for i, mode in enumerate(MODES):
        q = _Queue()
        os._exit = lambda code: (_ for _ in ()).throw(SystemExit(code))      # the worker ends a failed rank with os._exit(1)
        try:
            T._worker(rank, world, port + i, mode, q)
        except SystemExit:
            sys.stderr.write("\n".join(q.items) + "\n")
            real_exit(1)
        finally:
            os._exit = real_exit
        if rank == 0:
            print(f"mode {mode} ok", flush=True)


if __name__ == "__main__":
    port = 29500 + os.getpid() % 300
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
    print("multi dry run ok" if ok else f"exit codes {[p.exitcode for p in procs]}")
    sys.exit(0 if ok else 1)
