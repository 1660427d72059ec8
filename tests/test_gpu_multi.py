"""Multi-GPU (NCCL) parity: a 2-rank table-wise / row-wise sharded two-tower trained through the
reference-facing API equals the unsharded CPU oracle fed the same per-rank batches.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu
CAT = ["user_id", "product_id"]
EMB, DIM, LAYERS, B, LR, STEPS = [1500, 900], 64, [128, 64], 512, 0.02, 4


def _raw(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return {"user_id": torch.randint(0, EMB[0] * 2, (B,), generator=g).tolist(),
            "product_id": torch.randint(0, EMB[1] * 2, (B,), generator=g).tolist(),
            "label": torch.randint(0, 2, (B,), generator=g).tolist()}


def _worker(rank, world, port, sharding, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import oracle
        from oracle.ebc import TableSpec
        import two_tower_recommender_model_b200 as tt
        from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

        specs = [TableSpec(f"t_{c}", EMB[i], DIM, [c]) for i, c in enumerate(CAT)]
        ref = oracle.OracleTwoTower(specs, LAYERS, loss="bce", sparse_lr=LR, dense_lr=LR, seed=3, dense_optimizer="sgd")
        ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=DIM, num_embeddings=EMB[i], feature_names=[c])
                                                for i, c in enumerate(CAT)], device=torch.device("meta"))
        task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, LAYERS, device=dev))
        apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": LR})
        # *_peer: output exchange fused into the lookup kernels (NVLink peer memory);
        # *_dense*: batches arrive as dense id columns (from_id_columns) -> single all-to-all input dist, no host sync
        peer = sharding.endswith("_peer")
        dense_ids = "_dense" in sharding
        sharding = "table_wise" if sharding.startswith("table_wise") else ("row_wise" if sharding.startswith("row_wise") else sharding)
        cons = {f"t_{c}": ParameterConstraints(sharding_types=[sharding]) for c in CAT} if sharding != "planner" else None
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world), constraints=cons).collective_plan(task, tt.get_default_sharders(), dist.GroupMember.WORLD)
        model = tt.DistributedModelParallel(module=task, device=dev, plan=plan, sharding_kwargs={"peer_exchange": True} if peer else None)
        model.module.two_tower.load_state_dict(ref.torchrec_state_dict())
        opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: torch.optim.SGD(p, lr=LR))
        pipe = tt.TrainPipelineSparseDist(model, opt, dev)

        class RawIds:
            """Host-side batch of raw id columns; the KJT is built on the device (as bench.py does)."""

            def __init__(self, b):
                self.ids = torch.tensor([b[c] for c in CAT], dtype=torch.int64).pin_memory()
                self.labels = torch.tensor(b["label"], dtype=torch.int32).pin_memory()

            def to(self, device, non_blocking=False):
                kjt = tt.KeyedJaggedTensor.from_id_columns(CAT, self.ids.to(device, non_blocking=non_blocking), torch.tensor(EMB))
                return tt.Batch(torch.zeros(1, device=device), kjt, self.labels.to(device, non_blocking=non_blocking))

        def transform(b):
            if dense_ids:
                return RawIds(b)
            v, l, y = oracle.transform_to_torchrec_batch(b, CAT, EMB)
            return tt.Batch(torch.zeros(1), tt.KeyedJaggedTensor.from_lengths_sync(CAT, v, l), y)

        model.train()
        it = map(transform, (_raw(rank, s) for s in range(STEPS)))
        for s in range(STEPS):
            per_rank = [oracle.transform_to_torchrec_batch(_raw(r, s), CAT, EMB) for r in range(world)]
            losses = ref.train_step_ranks(CAT, per_rank)
            loss_d, _, _ = pipe.progress(it)
            torch.testing.assert_close(loss_d.cpu(), losses[rank], rtol=1e-4, atol=1e-6)
        # gather the sharded tables exactly as utils/model_training.py:161-182 does
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        want = ref.torchrec_state_dict()
        sd = model.module.two_tower.state_dict()
        assert set(sd) == set(want)
        for k, t in sd.items():
            if isinstance(t, ShardedTensor):
                full = torch.zeros(t.size(), device=dev) if rank == 0 else None
                t.gather(0, full)
            else:
                full = t
            if rank == 0:
                torch.testing.assert_close(full.cpu(), want[k], rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


MODES = ["table_wise", "row_wise", "table_wise_peer", "table_wise_dense", "table_wise_dense_peer", "row_wise_dense_peer"]


@pytest.mark.parametrize("sharding", MODES)
def test_two_rank_sharded_training_matches_oracle(sharding):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29800 + os.getpid() % 100 + MODES.index(sharding)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sharding, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)
