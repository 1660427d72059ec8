#!/usr/bin/env python
"""Headline benchmark: two-tower train samples/s (BASELINE.json `metric`), plus the EBC
lookup HBM GB/s and the dominant kernel's roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: two 10M-row x 64 fp32 tables (user_id / product_id, 5.12 GB,
far larger than the 126 MB L2, uniform random ids, so no L2 flush is needed between iterations),
batch 65536, towers 64->128->64, in-batch softmax loss, row-wise Adagrad fused into the embedding
backward, Adam on the towers.  Synthetic ids, random-init weights (no network for datasets).

  N = 1 : the whole config on one GPU.
  N > 1 : `value` is the config AS STATED -- GLOBAL batch 65536 (per-rank 65536/N, "strong" scaling),
          tables sharded table-wise; the same with row-wise sharding and the weak-scaling run
          (per-rank batch 65536) are timed beside it (`strong_row_wise`, `weak`).  Before anything is
          timed the sharded path is CHECKED: 3 train steps of the same sharded module / exchange mode
          on small tables against an unsharded replica on rank 0 (`parity`).  A failed table-wise check ends the
          run with rc 4 and no number; a failed row-wise check leaves the (table-wise) headline standing, and the
          blocks that use row-wise sharding are recorded as skipped instead of timed (`parity_failed`).
          Last come configs[3] and configs[2] AS STATED, sharded over the N GPUs (`cfg4_sharded`, `cfg3_row_wise`;
          tools/run_configs.py), each under its own time limit: an error or a cut there costs that block only.

One "step" = forward + backward + both optimizers on one batch.
  value : samples/s with the batch already resident in HBM (max over ranks, CUDA events).
  e2e   : the same through the public API: every step copies that step's raw id columns + labels from
          pinned host memory, builds the KeyedJaggedTensor on the device, trains, and reads the loss back.
  roofline / kernels : per-kernel CUDA-event times from a second, instrumented pass.
  cpu_baseline : the oracle port of the reference's CPU path on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CAT = ["user_id", "product_id"]
# rank 0's line once the headline block is measured + the block in flight: what the watchdog prints if a later block wedges
_PARTIAL = {"line": None, "stage": "headline"}
# watchdog deadlines (time.time()): the whole run, and the side block in flight (shorter: a wedged side block is cut early)
_DEADLINE = {"run": None, "block": None}


def stage(name, limit_s=None):
    """Names the block in flight and (re)arms its own time limit; ``limit_s=None`` leaves only the run's limit."""
    _PARTIAL["stage"] = name
    _DEADLINE["block"] = None if limit_s is None else time.time() + limit_s


BLOCK_LIMIT_S = float(os.environ.get("TT_BENCH_BLOCK_S", "180"))
CFG2 = dict(rows=[10_000_000, 10_000_000], dim=64, layers=[128, 64], batch=65536, loss="in_batch_softmax",
            sparse_lr=0.01, dense_lr=0.001)
CFG3_SHARDED = dict(batch=65536, big_rows=100_000_000, D=128, L=20)      # configs[2]; the dry runs shrink it
CFG4_SHARDED = dict(batch=262144, rows=(10_000_000, 10_000_000), D=128, layers=(1024, 512, 256))      # configs[3]
CFG1 = dict(rows=[200_000, 50_000], dim=64, layers=[128, 64], batch=1024, loss="bce", sparse_lr=0.01, dense_lr=0.001)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def measured_traffic(kernel_key):
    """DRAM bytes per launch from the committed ncu capture (profiles/r02_dram_traffic.json, written by
    tools/ncu_traffic.py from an `ncu --set full` report); None when no capture exists for the key."""
    p = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    e = d.get(kernel_key)
    return (e["dram_bytes"], "profiles/r02_dram_traffic.json:" + kernel_key) if e else (None, None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.t_mark = [], None, index, None

    def mark(self):
        """Samples that arrive before this call (nvidia-smi start-up, warm-up) are dropped."""
        self.t_mark = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            if self.t_mark is not None and time.time() >= self.t_mark:
                self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- ours
def build_model(cfg, dev, sharding=None, exchange="peer", fused_sparse=True, dense_opt="flat_adam"):
    """The reference's init sequence (03_model_training.py:770-829) against this package."""
    from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
    import torch.distributed as dist
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
    eb = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=cfg["dim"], num_embeddings=cfg["rows"][i], feature_names=[c])
          for i, c in enumerate(CAT)]
    ebc = tt.EmbeddingBagCollection(tables=eb, device=torch.device("meta"))
    prec = cfg.get("precision", "bf16")
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, cfg["layers"], device=dev, precision=prec), loss=cfg["loss"], precision=prec,
                                negatives=cfg.get("negatives", "local"))
    if fused_sparse:
        apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": cfg["sparse_lr"]})
    multi = dist.is_initialized() and dist.get_world_size() > 1
    plan, kw = None, None
    if multi:
        W = dist.get_world_size()
        cons = None
        if sharding is not None:
            cons = {f"t_{c}": ParameterConstraints(sharding_types=[sharding]) for c in CAT}
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=W), constraints=cons
                                           ).collective_plan(task, tt.get_default_sharders(), dist.GroupMember.WORLD)
        kw = {"peer_exchange": True} if exchange == "peer" else None
    model = tt.DistributedModelParallel(module=task, device=dev, plan=plan, sharding_kwargs=kw)
    if dense_opt == "flat_adam":
        opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: tt.FlatAdam(p, lr=cfg["dense_lr"]))
    else:
        opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: torch.optim.SGD(p, lr=cfg["dense_lr"]))
    return model, opt


def plan_kinds(model):
    try:
        return sorted({ps.sharding_type for tables in model._plan.plan.values() for ps in tables.values()})
    except Exception:
        return None


class RawBatch:
    """One step's raw inputs in pinned host memory: id columns [F, B] int64 + labels [B] int32.
    ``.to(device)`` is the H2D copy followed by the device-side batch construction
    (KeyedJaggedTensor.from_id_columns = transform_to_torchrec_batch, utils/model_training.py:43-69)."""

    def __init__(self, ids, labels, rows):
        self.ids, self.labels, self.rows = ids, labels, rows

    def nbytes(self):
        return self.ids.numel() * 8 + self.labels.numel() * 4

    def to(self, device, non_blocking=False):
        import two_tower_recommender_model_b200 as tt
        ids = self.ids.to(device, non_blocking=non_blocking)
        labels = self.labels.to(device, non_blocking=non_blocking)
        kjt = tt.KeyedJaggedTensor.from_id_columns(CAT, ids, self.rows)
        return tt.Batch(dense_features=torch.zeros(1, device=device), sparse_features=kjt, labels=labels)


def make_raw_batches(n, cfg, seed, rows_dev, batch=None):
    g = torch.Generator().manual_seed(seed)
    B = batch or cfg["batch"]
    out = []
    for _ in range(n):
        ids = torch.stack([torch.randint(1, r, (B,), generator=g) for r in cfg["rows"]]).pin_memory()
        labels = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32).pin_memory()
        out.append(RawBatch(ids, labels, rows_dev))
    return out


def algorithmic_bytes(cfg, B, uniq_per_table):
    """SURVEY.md 8(d) conventions: 8-byte ids, 4-byte offsets, each gathered row once, each unique updated row
    one read + one write of weights and state.  Per rank the rows gathered / updated add up to one per-rank
    batch per table whatever the sharding (table-wise: an owner looks up the global batch of ITS table)."""
    D, L, F = cfg["dim"], 1, len(cfg["rows"])
    fwd = F * (B * L * (8 + 4 * D) + 4 * B + 4 * B * D)
    bwd = sum(4 * B * D + 8 * B * L + u * (8 * D + 8) for u in uniq_per_table[:F])
    return fwd, bwd


# ---- sharded-vs-unsharded parity, run by the driver's own N > 1 launches (the driver never runs pytest on > 1 GPU)
def parity_check(world, rank, dev, sharding, exchange, steps=3, precision="bf16", graph=False):
    """Trains `steps` steps of the SAME sharded module / exchange mode the timed run uses (small tables, in-batch
    softmax with per-rank negatives, fused row-wise Adagrad, deterministic softmax backward) and, on rank 0, an
    UNSHARDED replica of the same model fed every rank's batch: its table gradients are accumulated densely
    (no fused optimizer), divided by the world size like the tower gradients (TorchRec's gradient division) and
    applied with the row-wise Adagrad formula in plain torch.  Compared: every rank's loss at every step, the
    gathered tables (ShardedTensor.gather, utils/model_training.py:161-182) and the tower weights at the end."""
    import torch.distributed as dist
    import two_tower_recommender_model_b200 as tt
    small = dict(rows=[3001, 1777], dim=64, layers=[128, 64], batch=int(os.environ.get("TT_PARITY_BATCH", 256)), loss="in_batch_softmax",
                 sparse_lr=0.05, dense_lr=0.05, precision=precision)
    tt.functional.set_deterministic_softmax_backward(True)
    try:
        torch.manual_seed(1234)
        model, opt = build_model(small, dev, sharding=sharding, exchange=exchange, dense_opt="sgd")
        kinds = plan_kinds(model)
        ref = None
        # gather the sharded model's initial weights so that the unsharded replica starts from the same point
        sd0 = model.module.two_tower.state_dict()
        full0 = {}
        from torch.distributed._shard.sharded_tensor import ShardedTensor
        for k, t in sd0.items():
            if isinstance(t, ShardedTensor):
                out = torch.zeros(t.size(), device=dev) if rank == 0 else None
                t.gather(0, out)
                full0[k] = out
            else:
                full0[k] = t.detach().clone()
        if rank == 0:
            eb = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=small["dim"], num_embeddings=small["rows"][i], feature_names=[c])
                  for i, c in enumerate(CAT)]
            ebc = tt.EmbeddingBagCollection(tables=eb, device=dev)
            ref = tt.TwoTowerTrainTask(tt.TwoTower(ebc, small["layers"], device=dev, precision=precision), loss=small["loss"], precision=precision)
            ref.two_tower.load_state_dict(full0)
            state = {c: torch.zeros(small["rows"][i], device=dev) for i, c in enumerate(CAT)}
        rows_dev = torch.tensor(small["rows"], dtype=torch.int64, device=dev)
        B = small["batch"]
        max_loss_err = 0.0
        model.train()
        # graph=True: the sharded side runs through CudaGraphTrainStep (1 eager step, then capture + replays) -- the path the
        # timed blocks use; a replay reuses ONE exchange buffer index, which the eager loop never exercises
        gstep = tt.CudaGraphTrainStep(model, opt, CAT, small["rows"], B, dev, warmup_steps=1) if graph else None
        for s in range(steps):
            raws = [make_raw_batches(1, small, 7000 + 100 * s + r, rows_dev, B)[0] for r in range(world)]
            if graph:
                loss = gstep(raws[rank].ids, raws[rank].labels)[0].clone()
            else:
                opt.zero_grad()
                loss, _ = model(raws[rank].to(dev))
                loss.backward()
                model.sync_dense_grads()
                opt.step()
            losses = [torch.zeros((), device=dev) for _ in range(world)]
            dist.all_gather(losses, loss.detach())
            if rank == 0:
                dense = [p for n, p in ref.named_parameters() if "embedding_bags" not in n]
                tabs = [ref.two_tower.ebc.embedding_bags[f"t_{c}"].weight for c in CAT]
                for p in dense + tabs:
                    p.grad = None
                for r in range(world):
                    l_r, _ = ref(raws[r].to(dev))
                    l_r.backward()
                    max_loss_err = max(max_loss_err, abs(float(l_r) - float(losses[r])))
                with torch.no_grad():
                    for p in dense:
                        p -= small["dense_lr"] * p.grad / world
                    for c, w in zip(CAT, tabs):
                        g = w.grad / world                                   # gradient division (comm_ops default)
                        state[c] += g.pow(2).mean(dim=1)
                        w -= small["sparse_lr"] * g / (state[c].sqrt() + 1e-10).unsqueeze(1)
        sd = model.module.two_tower.state_dict()
        # Weights are compared per ROW against the distance the row travelled: the towers run in bf16, so a summation-order
        # difference of 1e-7 in one embedding row (e.g. a short run of duplicate ids summed by two groups) can flip the bf16
        # rounding of one activation at the next step, which changes that sample's gradient row -- and the next Adagrad step
        # of the rows it touches -- by one bf16 ulp (2^-9 = 0.2 %; observed up to 0.3 % of a row's displacement, 8e-4
        # absolute at lr 0.05).  A routing or reduction error (an occurrence lost, counted twice or scaled wrongly) moves a
        # row by >= 10 % of its displacement, so 2 % separates the two (error / (displacement + 1e-4) per row).
        max_w_err, max_w, max_rel, worst = 0.0, 0.0, 0.0, None
        for k, t in sd.items():
            if isinstance(t, ShardedTensor):
                out = torch.zeros(t.size(), device=dev) if rank == 0 else None
                t.gather(0, out)
            else:
                out = t
            if rank == 0:
                want = ref.two_tower.state_dict()[k]
                w0 = full0[k]
                e2 = (out - want).reshape(out.shape[0], -1) if out.dim() >= 1 else (out - want).reshape(1, -1)
                d2 = (want - w0).reshape(e2.shape)
                err_row = e2.norm(dim=1)
                disp_row = d2.norm(dim=1)
                rel = err_row / (disp_row + 1e-4)      # 1e-4: floor under which a row's error is rounding, whatever it moved
                if float(rel.max()) > max_rel:
                    i = int(rel.argmax())
                    worst = {"key": k, "row": i, "rel_err_of_row_update": float(rel[i]), "abs_err": float(e2[i].abs().max()),
                             "rows_over_1e-4_abs": int((e2.abs().max(dim=1).values > 1e-4).sum()), "rows": int(e2.shape[0])}
                max_rel = max(max_rel, float(rel.max()))
                max_w_err = max(max_w_err, float(e2.abs().max()))
                max_w = max(max_w, float(want.abs().max()))
        tol_loss, tol_rel, tol_abs = 2e-4, 2e-2, 2e-3
        if precision == "fp32":      # exact-fp32 towers and softmax: no re-rounding, only summation order is left
            tol_loss, tol_rel, tol_abs = 2e-5, 1e-3, 2e-5
        res = {"mode": f"{'+'.join(kinds or [])}/{exchange}" + ("/cuda_graph" if graph else "/eager"), "precision": precision, "world": world,
               "steps": steps, "batch_per_rank": B,
               "rows": small["rows"], "max_abs_err_loss": max_loss_err, "max_rel_err_row_update": max_rel,
               "max_abs_err_weights": max_w_err, "weights_max_abs": max_w, "worst_entry": worst,
               "tol_loss": tol_loss, "tol_rel_row_update": tol_rel, "tol_abs_weights": tol_abs,
               "reference": "unsharded replica on rank 0 (same kernels, dense table gradients / world, row-wise Adagrad in torch); "
                            "row errors are measured against the row's displacement (bf16 towers: one-ulp flips, see bench.py)"}
        ok = torch.tensor([1 if (rank != 0 or (max_loss_err <= tol_loss and max_rel <= tol_rel and max_w_err <= tol_abs)) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["ok"] = bool(ok.item())
        del model, opt, ref, gstep
        torch.cuda.empty_cache()
        return res
    finally:
        tt.functional.set_deterministic_softmax_backward(False)


def ebc_lookup_alone(model, kjts, B, dev):
    """EBC lookup alone (BASELINE's second metric): 40 back-to-back lookups over rotating batches, one event pair around
    the loop (the lookup is a ~15 us kernel; tables are 5 GB of random rows, nothing is L2-resident).  Returns ms per call."""
    from ctypes import byref
    from two_tower_recommender_model_b200 import _native as N
    nb = len(kjts)
    ebc_mod = model.module.two_tower.ebc
    plan, total_dim = ebc_mod._build_plan(tuple(kjts[0].keys()), B, with_state=False)
    vals = [k.values().contiguous() for k in kjts]
    offs = [k.offsets().to(torch.int32).contiguous() for k in kjts]
    pooled = torch.empty(B, total_dim, dtype=torch.float32, device=dev)
    sp = N.stream_ptr(dev)
    for i in range(4):
        N.call("tt_ebc_forward", byref(plan), N.ptr(vals[i % nb]), N.ptr(offs[i % nb]), N.ptr(pooled), sp)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(40):
        N.call("tt_ebc_forward", byref(plan), N.ptr(vals[i % nb]), N.ptr(offs[i % nb]), N.ptr(pooled), sp)
    a1.record()
    torch.cuda.synchronize()
    return a0.elapsed_time(a1) / 40


def time_block(cfg, B, dev, rank, world, local, args, sharding, exchange, lib, with_kernels):
    """Builds the model for one (per-rank batch, sharding) point, captures the step as a CUDA graph and times the
    device-resident and the end-to-end loops.  Returns a dict (max over ranks applied by the caller)."""
    import torch.distributed as dist
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N
    model, opt = build_model(cfg, dev, sharding=sharding, exchange=exchange)
    model.train()
    rows_dev = torch.tensor(cfg["rows"], dtype=torch.int64, device=dev)
    nb = 4
    raw = make_raw_batches(nb, cfg, 1234 + rank, rows_dev, B)
    resident = [b.to(dev) for b in raw]
    torch.cuda.synchronize()

    def step(batch):
        opt.zero_grad()
        loss, out = model(batch)
        loss.backward()
        sync = getattr(model, "sync_dense_grads", None)
        if sync is not None:
            sync()
        opt.step()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph
    launches_per_step = None
    graph_step = None
    if use_graph:
        graph_step = tt.CudaGraphTrainStep(model, opt, CAT, cfg["rows"], B, dev, warmup_steps=3)
        dev_raw = [(b.ids.to(dev), b.labels.to(dev)) for b in raw]
        for i in range(3):
            graph_step(*dev_raw[i % nb])
        l_before = lib.tt_kernel_launch_count()
        graph_step(*dev_raw[3 % nb])          # capture (+ first replay)
        launches_per_step = int(lib.tt_kernel_launch_count() - l_before)
        torch.cuda.synchronize()

    # the sampler starts BEFORE the warm-up (nvidia-smi needs ~0.1 s to deliver its first line) and keeps what arrives
    # from the start of the timed region on
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        graph_step(*dev_raw[i % nb]) if use_graph else step(resident[i % nb])
    barrier()
    sampler.mark()
    l0 = lib.tt_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        graph_step(*dev_raw[i % nb]) if use_graph else step(resident[i % nb])
    e1.record()
    barrier()
    launches = launches_per_step * args.steps if use_graph else lib.tt_kernel_launch_count() - l0
    ms_value = e0.elapsed_time(e1) / args.steps

    total = args.warmup + args.steps
    if use_graph:
        e2e_api = "CudaGraphTrainStep(ids_pinned, labels_pinned) + float(loss)"
        for i in range(args.warmup):
            float(graph_step(raw[i % nb].ids, raw[i % nb].labels)[0])
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        for i in range(args.steps):
            last = float(graph_step(raw[i % nb].ids, raw[i % nb].labels)[0])  # device -> host read of the loss
        t1.record()
        barrier()
    else:
        e2e_api = "TrainPipelineSparseDist.progress(iterator of pinned raw batches) + float(loss)"
        pipe = tt.TrainPipelineSparseDist(model, opt, dev)
        it = iter(raw[i % nb] for i in range(total))
        for _ in range(args.warmup):
            float(pipe.progress(it)[0])
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        for _ in range(args.steps):
            last = float(pipe.progress(it)[0])
        t1.record()
        barrier()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    t = torch.tensor([ms_value, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e = t.tolist()
    # K steps of a few ms give nvidia-smi (20 ms period) only a handful of samples: after the timed regions the SAME step keeps
    # replaying under the sampler until it has covered ~0.5 s of this load.  The count is derived from the ALL-REDUCED time,
    # so it is the same on every rank (the sharded step has collectives: a rank-local count deadlocks)
    extra = min(5000, max(0, int(500.0 / max(ms_e2e, 1e-3)) - 2 * args.steps))
    for i in range(extra):
        graph_step(*dev_raw[i % nb]) if use_graph else step(resident[i % nb])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = ("device-resident + end-to-end timed steps + %d further replays of the same step "
                            "(the sampler is started before the warm-up)" % extra)

    # value / e2e are measured at this point; what follows explains them (per-call times, the lookup alone).  Every rank
    # runs the same eager steps, so a failure here is either common to all ranks or a fault that ends the run anyway.
    per_call, explain_error = {}, None
    if with_kernels:
        try:
            N.enable_timing(True)
            for i in range(min(args.steps, 5)):
                step(resident[i % nb])
            torch.cuda.synchronize()
            per_call = N.timing_summary()
        except Exception as e:      # noqa: BLE001 -- recorded in the line (`explain_error`), the measured value stands
            explain_error = f"per-call pass: {type(e).__name__}: {e}"[:300]
        finally:
            N.enable_timing(False)

    ebc_only_ms = None
    if with_kernels and world == 1 and explain_error is None:
        try:
            ebc_only_ms = ebc_lookup_alone(model, [resident[i].sparse_features for i in range(nb)], B, dev)
        except Exception as e:      # noqa: BLE001
            explain_error = f"lookup-alone loop: {type(e).__name__}: {e}"[:300]
    uniq = [int(torch.unique(resident[0].sparse_features[c].values()[:B]).numel()) for c in CAT]
    out = {"ms_value": ms_value, "ms_e2e": ms_e2e, "launches": int(launches), "clocks": clocks, "per_call": per_call,
           "ebc_only_ms": ebc_only_ms, "uniq": uniq, "last_loss": last, "e2e_api": e2e_api, "h2d": raw[0].nbytes(),
           "sharding": plan_kinds(model) if world > 1 else None, "cuda_graph": bool(use_graph), "batch": B,
           "explain_error": explain_error}
    del model, opt, graph_step, resident, raw
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist
    from two_tower_recommender_model_b200 import _native as N

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(CFG2)
    lib = N.load()
    pk = peaks()
    G = cfg["batch"]                     # configs[1]: batch 65536

    parity = []
    if world == 1:
        main = time_block(cfg, G, dev, rank, world, local, args, None, None, lib, with_kernels=True)
        workload = ("BASELINE configs[1] on 1 GPU: 2 tables 10M x 64 fp32, batch 65536, MLP 64-128-64, in-batch softmax, "
                    "fused row-wise Adagrad, Adam")
        scaling = "strong"
    else:
        if G % world != 0:
            raise SystemExit("world size must divide 65536")
        # sharded-path parity first, on exactly the (sharding, exchange) pairs that are timed below
        modes = [("table_wise", "bf16", False), ("row_wise", "bf16", False), ("table_wise", "fp32", False), ("row_wise", "fp32", False)]
        if args.parity_graph:      # opt-in: the sharded side replays a captured graph (exercised by tests/test_gpu_multi.py instead)
            modes += [("table_wise", "fp32", True), ("row_wise", "fp32", True)]
        failed = set()
        for sh, prec, gr in modes:
            p = parity_check(world, rank, dev, sh, args.exchange, steps=4 if gr else 3, precision=prec, graph=gr)
            parity.append(p)
            if not p["ok"]:            # the same on every rank (all-reduced inside parity_check)
                failed.add(sh)
        if args.parity_only:
            if rank == 0:
                print(json.dumps({"parity": parity}))
            leave(world, 0 if not failed else 4)
            return
        if "table_wise" in failed:
            # the headline's own sharding failed its check: no number is better than a number without parity
            if rank == 0:
                print(json.dumps({"metric": "two-tower train samples/s", "error": "sharded parity check failed", "parity": parity}))
            leave(world, 4)
            return
        args.parity_failed = sorted(failed)      # a sharding that failed is not timed (side_blocks records why)
        main = time_block(cfg, G // world, dev, rank, world, local, args, "table_wise", args.exchange, lib, with_kernels=True)
        workload = ("BASELINE configs[1] on %d GPUs as stated: 2 tables 10M x 64 fp32 table-wise sharded, GLOBAL batch %d "
                    "(per-rank %d), MLP 64-128-64, in-batch softmax (per-rank negatives), fused row-wise Adagrad, Adam" % (world, G, G // world))
        scaling = "strong"
    line = None
    if rank == 0:
        line = headline(args, cfg, main, pk, world, G, workload, scaling, parity)
        _PARTIAL["line"] = line          # from here on a wedged side block costs that block, not the headline (see watchdog)
    if world > 1:
        side_blocks(args, cfg, dev, rank, world, local, lib, G, line)
    if rank != 0:
        leave(world)
        return
    finish(args, cfg, dev, world, line)


def side_blocks(args, cfg, dev, rank, world, local, lib, G, line):
    """N > 1: the blocks timed BESIDE the headline (every rank runs them, rank 0 records them).  A block that raises on
    every rank is recorded as an error and the run goes on; one that wedges is cut by the watchdog, which still prints
    the headline line with an `incomplete` entry."""
    def record(name, fn):
        stage(name, BLOCK_LIMIT_S)
        try:
            out = fn()
        except Exception as e:      # noqa: BLE001 -- recorded, not hidden: the line says which block failed and why
            out = {"error": f"{type(e).__name__}: {e}"[:400]}
            sys.stderr.write(f"bench.py: block {name} failed on rank {rank}: {out['error']}\n")
        if line is not None:
            line[name] = out

    def strong_row_wise():
        srw = time_block(cfg, G // world, dev, rank, world, local, args, "row_wise", args.exchange, lib, with_kernels=False)
        return {"value": round(G / (srw["ms_value"] * 1e-3), 1), "ms_per_step": round(srw["ms_value"], 4),
                "e2e": round(G / (srw["ms_e2e"] * 1e-3), 1), "global_batch": G, "per_rank_batch": G // world,
                "sharding": srw["sharding"], "last_loss": srw["last_loss"]}

    def weak():
        wk = time_block(cfg, G, dev, rank, world, local, args, None, args.exchange, lib, with_kernels=False)
        return {"value": round(world * G / (wk["ms_value"] * 1e-3), 1), "ms_per_step": round(wk["ms_value"], 4),
                "e2e": round(world * G / (wk["ms_e2e"] * 1e-3), 1), "global_batch": G * world, "per_rank_batch": G,
                "sharding": wk["sharding"], "last_loss": wk["last_loss"], "note": "per-rank batch fixed at 65536 (round 1's headline)"}

    def strong_global_negatives():
        sgl = time_block(dict(cfg, negatives="global"), G // world, dev, rank, world, local, args, "table_wise", args.exchange, lib,
                         with_kernels=False)
        return {"value": round(G / (sgl["ms_value"] * 1e-3), 1), "ms_per_step": round(sgl["ms_value"], 4),
                "e2e": round(G / (sgl["ms_e2e"] * 1e-3), 1), "global_batch": G, "per_rank_batch": G // world,
                "sharding": sgl["sharding"], "last_loss": sgl["last_loss"],
                "note": "every rank's candidates are negatives for every rank's queries (all-gather + reduce-scatter): the SAME loss "
                        "function as the 1-GPU run at global batch 65536; per-rank logits flops = 1/N of the 1-GPU step"}

    bad = set(getattr(args, "parity_failed", []) or [])
    # weak scaling lets the planner choose: row-wise once the ranks outnumber the two tables
    weak_kind = "row_wise" if world > len(cfg["rows"]) else "table_wise"

    def not_timed(kind):
        return lambda: {"skipped": f"the {kind} parity check of this run failed (see `parity`): not timed"}

    record("strong_row_wise", not_timed("row_wise") if "row_wise" in bad else strong_row_wise)
    record("weak", not_timed(weak_kind) if weak_kind in bad else weak)
    record("strong_global_negatives", strong_global_negatives)
    record("retrieval", lambda: retrieval_probe_sharded(dev, rank, world))

    def cfg3_row_wise():
        # configs[2] as stated (row-wise sharded 100M-row tables, L = 20) at this N.  LAST: this block and the one before it
        # have not been timed on GPUs before -- if one wedges, the watchdog cuts it and every block above is already in the line
        if args.exchange != "peer":
            return {"skipped": "needs the peer-memory exchange (the sync-free multi-hot input dist is what makes the step capturable)"}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_configs
        return run_configs.config3_sharded(dev, rank, world, steps=max(args.steps, 5), warmup=args.warmup, **CFG3_SHARDED)

    def cfg4_sharded():
        # configs[3] as stated (B = 262144 global, bf16 towers 1024-512-256, d = 256 logits, row-wise Adam) at this N: id-column
        # batches, so every piece of it is a path the blocks above already ran -- the combination is what is new
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_configs
        return run_configs.config4_sharded(dev, rank, world, steps=max(args.steps, 5), warmup=args.warmup,
                                           peer_exchange=(args.exchange == "peer"), **CFG4_SHARDED)

    if not args.no_other_configs:
        record("cfg4_sharded", cfg4_sharded)
        record("cfg3_row_wise", cfg3_row_wise)
    stage("done")


def headline(args, cfg, main, pk, world, G, workload, scaling, parity):
    """The JSON line of the headline block (configs[1]); the other blocks are added to it afterwards."""
    B = main["batch"]
    per_call = main["per_call"]
    fwd_bytes, bwd_bytes = algorithmic_bytes(cfg, B, main["uniq"])
    d_out = cfg["layers"][-1]
    logit_flops = 6.0 * B * B * d_out
    kernels = {}

    def add(name, work, unit, peak, bound):
        if name in per_call and per_call[name]["ms"] > 0:
            ach = work / (per_call[name]["ms"] * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
            kernels[name] = {"ms": round(per_call[name]["ms"], 4), "achieved": round(ach, 2), "unit": unit, "peak": peak,
                             "frac": round(ach / peak, 4), "bound": bound}

    add("tt_ebc_forward", fwd_bytes, "GB/s", pk["hbm"], "hbm")
    add("tt_ebc_backward_fused", bwd_bytes, "GB/s", pk["hbm"], "hbm")
    add("tt_ebc_forward_peer", fwd_bytes, "GB/s", pk["hbm"], "hbm")          # world > 1: rows leave / gradients arrive over NVLink
    add("tt_ebc_backward_fused_peer", bwd_bytes, "GB/s", pk["hbm"], "hbm")
    sm_ms = sum(per_call.get(n, {"ms": 0})["ms"] for n in ("tt_inbatch_softmax_forward_f32", "tt_inbatch_softmax_backward_f32",
                                                          "tt_inbatch_softmax_forward_bf16", "tt_inbatch_softmax_backward_bf16"))
    roof = None
    if sm_ms > 0:
        ach = logit_flops / (sm_ms * 1e-3) / 1e12
        traffic, traffic_src = measured_traffic("softmax_fwd_bwd_B%d_d%d" % (B, d_out))
        roof = {"kernel": "in-batch softmax: tc_softmax_fwd_kernel + tc_softmax_bwd_fused_kernel (tcgen05)", "bound": "tensor",
                "achieved": round(ach, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4),
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"] + " (sustained bf16)", "ms": round(sm_ms, 4),
                "flops_credited": "6*B*B*d (recomputation of S in the backward is not credited)",
                "note": "at d=64 the forward is MUFU(ex2)-bound (1 ex2 per 64 MACs) and the one-pass backward is bound by shared-memory "
                        "operand bandwidth of its N=64 tcgen05.mma; see DESIGN.md section 4"}
    total_batch = G
    line = {
        "metric": "two-tower train samples/s", "value": round(total_batch / (main["ms_value"] * 1e-3), 1), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(main["ms_value"], 4),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "bf16 tower + logits GEMMs (fp32 accumulate, fp32 master weights) + f32 embeddings/optimizers", "data": "synthetic",
        "config": {"workload": workload, "per_rank_batch": B, "global_batch": total_batch, "cuda_graph": main["cuda_graph"],
                   "exchange": (args.exchange if world > 1 else None), "sharding": main["sharding"],
                   "l2": "tables 5.12 GB >> 126 MB L2, random ids; no flush needed"},
        "e2e": {"value": round(total_batch / (main["ms_e2e"] * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(main["ms_e2e"], 4),
                # whole job: every rank copies its own share of the batch in and reads its own loss back
                "h2d_bytes_per_step": main["h2d"] * world, "d2h_bytes_per_step": 4 * world,
                "h2d_bytes_per_step_per_rank": main["h2d"], "last_loss": main["last_loss"], "api": main["e2e_api"]},
        "gpu_launches": main["launches"], "clocks": main["clocks"], "roofline": roof, "kernels": kernels,
        "calls_ms": {k: round(v["ms"], 4) for k, v in sorted(per_call.items(), key=lambda kv: -kv[1]["ms"])},
        "ebc_lookup_gbs": kernels.get("tt_ebc_forward", kernels.get("tt_ebc_forward_peer", {})).get("achieved"),
    }
    if main.get("explain_error"):
        line["explain_error"] = main["explain_error"]
    if parity:
        line["parity"] = parity
        if getattr(args, "parity_failed", None):
            line["parity_failed"] = list(args.parity_failed)
    if main["ebc_only_ms"]:
        ms = main["ebc_only_ms"]
        line["ebc_lookup"] = {"gbs": round(fwd_bytes / (ms * 1e-3) / 1e9, 1), "us": round(ms * 1e3, 2), "peak_gbs": pk["hbm"],
                              "frac": round(fwd_bytes / (ms * 1e-3) / 1e9 / pk["hbm"], 4), "bytes": int(fwd_bytes),
                              "how": "40 back-to-back tt_ebc_forward launches over 4 rotating batches, CUDA events around the loop"}
        line["ebc_lookup_gbs"] = line["ebc_lookup"]["gbs"]
    return line


def finish(args, cfg, dev, world, line):
    """Rank 0: the single-GPU side blocks (retrieval probes, configs[2] / configs[3], CPU baselines), then the line."""
    def record(name, fn):
        """A side block that raises becomes an `error` entry: the headline line is printed whatever happens after it."""
        try:
            line[name] = fn()
        except Exception as e:      # noqa: BLE001 -- recorded in the line, configs[1] stays the value
            line[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            sys.stderr.write(f"bench.py: block {name} failed: {line[name]['error']}\n")
            try:
                if torch.cuda.is_available():
                    torch.cuda.empty_cache()
            except Exception:       # noqa: BLE001 -- a sticky CUDA error: the later blocks will record it too, the line still prints
                pass

    if world == 1 and not args.no_cpu_baseline:
        # first: `cpu_baseline` is part of the bench contract, the blocks after it are extras
        stage("cpu_baseline", BLOCK_LIMIT_S)
        record("cpu_baseline", lambda: cpu_baseline(cfg, steps=1, warmup=1))
        record("cpu_baseline_cfg1", cpu_baseline_cfg1)
    if world == 1:
        stage("retrieval", BLOCK_LIMIT_S)
        record("retrieval", lambda: retrieval_probe(dev))
        record("retrieval_large", lambda: retrieval_probe(dev, n_items=10_000_000, n_queries=131072))
    if world == 1 and not args.no_other_configs:
        # BASELINE configs[2] and configs[3] at full size on this GPU (eager steps, CUDA events); configs[1] stays the `value`
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        torch.cuda.empty_cache()
        for name in ("cfg3", "cfg4"):
            stage(name, BLOCK_LIMIT_S)

            def run(name=name):
                import run_configs
                return {"cfg3": run_configs.config3, "cfg4": run_configs.config4}[name]()
            record(name, run)
    stage("done")
    _PARTIAL["line"] = None
    print(json.dumps(line))
    if world > 1:
        # ranks still stuck in a side block that rank 0 left through an error entry learn from this file, when their own
        # watchdog fires, that the line is out (see watchdog())
        try:
            open(_marker_path(), "w").close()
        except OSError:
            pass
    leave(world)


def leave(world, rc=0):
    """Multi-rank exit: the captured graphs hold NCCL kernels and tearing the communicator down under them can
    block, so flush and exit the process without the (purely cosmetic) process-group teardown."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        try:
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001 -- a device fault inside a recorded side block must not change the exit code
            sys.stderr.write(f"bench.py: synchronize at exit failed: {type(e).__name__}: {e}\n")
            sys.stderr.flush()
        os._exit(rc)
    if rc:
        sys.exit(rc)


def retrieval_probe(dev, n_items=2_000_000, n_queries=16384, d=64, k=100):
    """Secondary number (BASELINE configs[4] shape): top-100 by dot product over a resident bf16 corpus of
    RANDOM-NORMAL vectors (trained-like: distinct scores, the candidate path is exercised), tcgen05 scoring
    with the top-k fused in the epilogue.  The large call is one GPU's share of configs[4] on 8 GPUs."""
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator(device=dev).manual_seed(7)
    items = torch.randn(n_items, d, device=dev, generator=g)
    queries = torch.randn(n_queries, d, device=dev, generator=g)
    index = tt.BruteForceIndex(items, precision="bf16")
    del items
    index.search(queries[:1024], k)
    index.search(queries, k, query_chunk=1 << 17)           # same shape as the timed call: allocations settled
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s, _ = index.search(queries, k, query_chunk=1 << 17)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    tf = 2.0 * n_queries * n_items * d / (ms * 1e-3) / 1e12
    pk = peaks()
    return {"queries_per_s": round(n_queries / (ms * 1e-3), 1), "ms": round(ms, 3), "items": n_items, "queries": n_queries,
            "k": k, "d": d, "tflops": round(tf, 1), "frac_of_sustained_bf16_peak": round(tf / pk["tf_sustained"], 4),
            "data": "randn", "top1_score_mean": round(float(s[:, 0].mean()), 3), "dtype": "bf16 scoring, f32 accumulate"}


def retrieval_probe_sharded(dev, rank, world, n_items=10_000_000, q_per_rank=131072, d=64, k=100):
    """BASELINE configs[4] on N GPUs: the corpus lives sharded (each rank holds the embeddings of a contiguous id
    range), is all-gathered once as bf16, and every rank answers its own 131072 queries (8 x 131072 = 1M queries on
    8 GPUs).  Timed: the search (max over ranks); the one-off all-gather is reported separately."""
    import torch.distributed as dist
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    per = -(-n_items // world)
    n_local = max(0, min(per, n_items - rank * per))
    local = torch.randn(n_local, d, device=dev, generator=g)
    queries = torch.randn(q_per_rank, d, device=dev, generator=g)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    index = tt.BruteForceIndex.from_sharded(local, precision="bf16")
    torch.cuda.synchronize()
    gather_s = time.perf_counter() - t0
    del local
    index.search(queries[:1024], k)
    index.search(queries, k, query_chunk=1 << 17)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s, _ = index.search(queries, k, query_chunk=1 << 17)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tf = 2.0 * q_per_rank * n_items * d / (ms * 1e-3) / 1e12
    pk = peaks()
    return {"queries_per_s_total": round(world * q_per_rank / (ms * 1e-3), 1), "queries_per_s_per_gpu": round(q_per_rank / (ms * 1e-3), 1),
            "ms": round(ms, 3), "items": n_items, "queries_total": world * q_per_rank, "k": k, "d": d,
            "tflops_per_gpu": round(tf, 1), "frac_of_sustained_bf16_peak": round(tf / pk["tf_sustained"], 4), "data": "randn",
            "corpus_all_gather_s": round(gather_s, 3), "how": "queries sharded over ranks, bf16 corpus all-gathered once"}


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_baseline(cfg, steps, warmup, sample_batch=None, budget_s=None):
    """Oracle port of the reference's unsharded CPU path (dense [R,D] embedding gradient + row-wise Adagrad over
    the whole table each step, as nn.EmbeddingBag + a grad hook do) at the FULL configs[1] batch.  ``budget_s``
    bounds the run: the first step sizes it (it is a warm-up step when the budget leaves room for one) -- as many
    of the `steps` timed steps as fit, then as many of the `warmup` steps as the rest allows."""
    import oracle
    from oracle.ebc import TableSpec
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    specs = [TableSpec(f"t_{c}", cfg["rows"][i], cfg["dim"], [c]) for i, c in enumerate(CAT)]
    m = oracle.OracleTwoTower(specs, cfg["layers"], loss="softmax" if cfg["loss"] != "bce" else "bce",
                              sparse_lr=cfg["sparse_lr"], dense_lr=cfg["dense_lr"], seed=0)
    g = torch.Generator().manual_seed(0)
    Bs = min(sample_batch or cfg["batch"], cfg["batch"])
    steps0, warmup0, all_dt, elapsed = steps, warmup, [], 0.0
    i = 0
    while i < warmup + steps:
        vals = torch.cat([torch.randint(1, r, (Bs,), generator=g) for r in cfg["rows"]])
        lens = torch.ones(2 * Bs, dtype=torch.int32)
        y = torch.randint(0, 2, (Bs,), generator=g, dtype=torch.int32)
        t0 = time.perf_counter()
        m.train_step(CAT, vals, lens, y)
        dt = time.perf_counter() - t0
        all_dt.append(dt)
        elapsed += dt
        i += 1
        if budget_s is not None and i <= 2:      # the (cold) first step sizes the run, the second one corrects it
            steps, warmup = bounded_steps(steps0, warmup0, budget_s - elapsed, dt, done=i)
    times = all_dt[warmup:warmup + steps]
    sec = sum(times) / len(times)
    return {"value": round(Bs / sec, 1), "unit": "samples/s", "cores": cores, "kind": "port", "ms_per_step": round(sec * 1e3, 2),
            "steps": steps, "warmup": warmup,
            "sample": f"{Bs} of the {cfg['batch']}-sample batch per step (in-batch negatives = {Bs}), full 10M-row tables, "
                      f"{steps} timed step(s) after {warmup} warm-up"}


def bounded_steps(steps, warmup, remaining_s, step_s, done=0):
    """How many timed / warm-up steps fit when `done` steps are behind us, `remaining_s` of the budget are left and
    a step takes `step_s`: timed steps first (at least one), then warm-up steps from what is left."""
    afford = done + int(max(0.0, remaining_s) / max(step_s, 1e-3))
    afford = max(1, afford)
    steps = max(1, min(steps, afford))
    return steps, max(0, min(warmup, afford - steps))


def cpu_baseline_cfg1(steps=30, warmup=5):
    """configs[0] (B 1024, 200k / 50k rows, BCE, CPU) -- the reference's own CPU-runnable case --
    plus the separate timing of transform_to_torchrec_batch (utils/model_training.py:43-69), whose Python loop is a
    large share of the reference's real step."""
    import oracle
    from oracle.ebc import TableSpec
    cfg = CFG1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    specs = [TableSpec(f"t_{c}", cfg["rows"][i], cfg["dim"], [c]) for i, c in enumerate(CAT)]
    m = oracle.OracleTwoTower(specs, cfg["layers"], loss="bce", sparse_lr=cfg["sparse_lr"], dense_lr=cfg["dense_lr"], seed=0)
    g = torch.Generator().manual_seed(1234)
    B = cfg["batch"]
    t_step, t_tr = [], []
    for i in range(warmup + steps):
        raw = {c: torch.randint(1, cfg["rows"][j], (B,), generator=g).tolist() for j, c in enumerate(CAT)}
        raw["label"] = torch.randint(0, 2, (B,), generator=g).tolist()
        t0 = time.perf_counter()
        vals, lens, y = oracle.transform_to_torchrec_batch(raw, CAT, cfg["rows"])
        t1 = time.perf_counter()
        m.train_step(CAT, vals, lens, y)
        t2 = time.perf_counter()
        if i >= warmup:
            t_tr.append(t1 - t0)
            t_step.append(t2 - t1)
    st, tr = sum(t_step) / len(t_step), sum(t_tr) / len(t_tr)
    return {"value": round(B / (st + tr), 1), "unit": "samples/s", "cores": cores, "kind": "port", "batch": B,
            "ms_per_step_model": round(st * 1e3, 3), "ms_per_step_transform_to_torchrec_batch": round(tr * 1e3, 3),
            "sample": f"configs[0]: B={B}, tables {cfg['rows']} x {cfg['dim']}, BCE, {steps} timed steps after {warmup} warm-up, "
                      "transform (Python loop of the reference) + train step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    cfg = dict(CFG2)
    # the FULL configs[1] batch per step; as many of the asked-for steps / warm-up steps as a few minutes of CPU time hold
    # (16 host cores: ~8 s per step, so --steps 20 --warmup 5 runs as asked)
    budget_s = float(os.environ.get("TT_REFERENCE_BUDGET_S", "240"))
    cb = cpu_baseline(cfg, steps=max(1, args.steps), warmup=max(0, args.warmup), budget_s=budget_s)
    steps, warmup = cb["steps"], cb["warmup"]
    line = {"impl": "reference", "metric": "two-tower train samples/s", "value": cb["value"], "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] (CPU port of the reference's unsharded TorchRec path; torchrec/fbgemm "
                                   "are not installable here)", "per_rank_batch": cfg["batch"], "global_batch": cfg["batch"],
                       "steps_requested": args.steps, "warmup_requested": args.warmup,
                       "steps_note": "bounded so that the CPU run ends within a few minutes (TT_REFERENCE_BUDGET_S, %d s)" % budget_s},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _marker_path():
    return os.path.join("/tmp", "tt_bench_partial_%s" % os.environ.get("MASTER_PORT", "0"))


def watchdog(why=None):
    """The run overran its limit (a wedged collective or capture) -- or, with `why`, was told to end.  If the headline block had been measured, rank 0
    still prints its line -- with an `incomplete` entry naming the block that was cut -- and the ranks exit 0; before
    that point there is nothing to report and the exit code is 3.  The other ranks wait 15 s longer than rank 0 and
    learn the outcome from a marker file."""
    rank = int(os.environ.get("RANK", 0))
    marker = _marker_path()
    sys.stderr.write("bench.py: %s on rank %d during block '%s'\n" % ("watchdog expired" if why is None else why, rank, _PARTIAL["stage"]))
    rc = 3
    if rank == 0:
        line = _PARTIAL["line"]
        if line is not None:
            line = dict(line)
            line["incomplete"] = {"cut_block": _PARTIAL["stage"],
                                  "reason": ("watchdog: the block did not finish within its limit" if why is None else why) +
                                  "; the headline block (value / e2e / roofline / parity) had completed before it started"}
            try:
                print(json.dumps(line))
                open(marker, "w").close()
                rc = 0
            except Exception:       # noqa: BLE001 -- a block was mutating the dict: nothing printable
                rc = 3
    elif os.path.exists(marker) and time.time() - os.path.getmtime(marker) < 600:
        rc = 0
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(rc)


def start_watchdog(run_limit_s):
    """A daemon thread that ends the process (through ``watchdog``) once the run's limit, or the limit of the side block in
    flight (``stage``), has passed.  Ranks other than 0 wait 15 s longer so that rank 0 can print first."""
    grace = 0 if int(os.environ.get("RANK", 0)) == 0 else 15
    _DEADLINE["run"] = time.time() + run_limit_s
    if grace == 0:
        try:
            os.remove(_marker_path())       # a marker left by an earlier run on the same port says nothing about this one
        except OSError:
            pass

    # SIGTERM is what torchrun sends the surviving ranks when one rank dies.  A Python-level handler only runs once the main
    # thread is back in the interpreter -- not while it waits inside a collective -- so the signal is routed to a pipe
    # (signal.set_wakeup_fd: written by the C-level handler at once) that the watchdog thread polls: rank 0 still prints the
    # headline line, if it has one, before the process goes.
    term_r = None
    try:
        import signal
        term_r, term_w = os.pipe()
        os.set_blocking(term_r, False)
        os.set_blocking(term_w, False)
        signal.signal(signal.SIGTERM, lambda signum, frame: None)
        signal.set_wakeup_fd(term_w, warn_on_full_buffer=False)
    except (ValueError, OSError, AttributeError):      # not the main thread (tests import this module), or no pipes
        term_r = None

    def poll():
        while True:
            time.sleep(0.5)
            now = time.time() - grace
            blk = _DEADLINE["block"]
            if now > _DEADLINE["run"] or (blk is not None and now > blk):
                watchdog()
            if term_r is not None:
                try:
                    got = os.read(term_r, 64)
                except (BlockingIOError, OSError):
                    got = b""
                if got and signal.SIGTERM in got:
                    watchdog("SIGTERM (torchrun ends the surviving ranks when one rank dies)")

    threading.Thread(target=poll, daemon=True).start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs[2] / configs[3] blocks (102 GB of tables)")
    ap.add_argument("--exchange", default=os.environ.get("TT_EXCHANGE", "peer"), choices=["nccl", "peer"],
                    help="N>1: output exchange by NCCL all-to-all, or fused into the lookup kernels over NVLink peer memory")
    ap.add_argument("--parity-only", action="store_true", help="N > 1: run the sharded-vs-unsharded parity checks and exit")
    ap.add_argument("--parity-graph", action="store_true", help="N > 1: also check the sharded model through CudaGraphTrainStep replays")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly (N>1: through TrainPipelineSparseDist) instead of replaying a CUDA graph")
    args = ap.parse_args()
    # a wedged collective / capture must not hold the box: the default run takes about a minute
    start_watchdog(float(os.environ.get("TT_BENCH_WATCHDOG_S", "600")))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
