#!/usr/bin/env python
"""Headline benchmark: two-tower train samples/s (BASELINE.json `metric`), plus the EBC
lookup HBM GB/s and the dominant kernel's roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE.json configs[1] on ONE GPU -- two 10M-row x 64 fp32 tables
(user_id / product_id, 5.12 GB, far larger than the 126 MB L2, uniform random ids, so no L2
flush is needed between iterations), per-rank batch 65536, towers 64->128->64, in-batch
softmax loss, row-wise Adagrad fused into the embedding backward, Adam on the towers.
Synthetic ids, random-init weights (no network for datasets).

One "step" = forward + backward + both optimizers on one batch.
  value : samples/s with the batch already resident in HBM (max over ranks, CUDA events).
  e2e   : the same through the public API (TrainPipelineSparseDist.progress): every step
          copies that step's raw id columns + labels from pinned host memory, builds the
          KeyedJaggedTensor on the device, trains, and reads the loss back to the host.
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
CFG2 = dict(rows=[10_000_000, 10_000_000], dim=64, layers=[128, 64], batch=65536, loss="in_batch_softmax",
            sparse_lr=0.01, dense_lr=0.001)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
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
def build_model(cfg, dev):
    from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
    import two_tower_recommender_model_b200 as tt
    eb = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=cfg["dim"], num_embeddings=cfg["rows"][i], feature_names=[c])
          for i, c in enumerate(CAT)]
    ebc = tt.EmbeddingBagCollection(tables=eb, device=torch.device("meta"))
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, cfg["layers"], device=dev, precision=cfg.get("precision", "bf16")), loss=cfg["loss"], precision=cfg.get("precision", "bf16"))
    apply_optimizer_in_backward(tt.RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": cfg["sparse_lr"]})
    peer = cfg.get("exchange", "nccl") == "peer" and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    plan = None
    if peer and os.environ.get("TT_BENCH_SHARDING") and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints   # diagnostics: force a sharding type
        cons = {f"t_{c}": ParameterConstraints(sharding_types=[os.environ["TT_BENCH_SHARDING"]]) for c in CAT}
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=torch.distributed.get_world_size()), constraints=cons
                                           ).collective_plan(task, tt.get_default_sharders(), torch.distributed.GroupMember.WORLD)
    if not peer and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        # the NCCL exchange path is benchmarked table-wise (its row-wise input dist needs a host sync per step)
        from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
        cons = {f"t_{c}": ParameterConstraints(sharding_types=["table_wise"]) for c in CAT}
        plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=torch.distributed.get_world_size()), constraints=cons
                                           ).collective_plan(task, tt.get_default_sharders(), torch.distributed.GroupMember.WORLD)
    model = tt.DistributedModelParallel(module=task, device=dev, plan=plan, sharding_kwargs={"peer_exchange": True} if peer else None)
    opt = tt.KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: tt.FlatAdam(p, lr=cfg["dense_lr"]))
    return model, opt


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


def make_raw_batches(n, cfg, seed, rows_dev):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ids = torch.stack([torch.randint(1, r, (cfg["batch"],), generator=g) for r in cfg["rows"]]).pin_memory()
        labels = torch.randint(0, 2, (cfg["batch"],), generator=g, dtype=torch.int32).pin_memory()
        out.append(RawBatch(ids, labels, rows_dev))
    return out


def algorithmic_bytes(cfg, uniq_per_table, world=1):
    """SURVEY.md 8(d) conventions: 8-byte ids, 4-byte offsets, each gathered row once, each
    unique updated row one read + one write of weights and state.  world > 1: table-wise (2 ranks, 2
    tables) rank 0 owns ONE table and looks up the GLOBAL batch of its feature; row-wise (more ranks than
    tables) every rank scans the offsets of the global batch and serves ~1/world of the ids of BOTH tables.
    Either way the rows a rank gathers / updates add up to one per-rank batch per table."""
    D, L, F = cfg["dim"], 1, len(cfg["rows"])
    B = cfg["batch"]
    scan = 4 * B * F * (world if world > F else 1)          # offsets of every bag this rank walks over
    fwd = F * (B * L * (8 + 4 * D) + 4 * B * D) + scan
    bwd = sum(4 * B * D + 8 * B * L + u * (8 * D + 8) for u in uniq_per_table[:F]) + scan
    return fwd, bwd


def run_ours(args):
    import torch.distributed as dist
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200 import _native as N

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(CFG2)
    cfg["exchange"] = args.exchange
    lib = N.load()
    model, opt = build_model(cfg, dev)
    model.train()
    rows_dev = torch.tensor(cfg["rows"], dtype=torch.int64, device=dev)
    nb = 4
    raw = make_raw_batches(nb, cfg, 1234 + rank, rows_dev)
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
    graph_step = None
    launches_per_step = None
    if use_graph:
        # single GPU: the whole step (device KJT build, lookup, towers, loss, backward + fused update,
        # Adam) is one CUDA graph; warm-up calls run eagerly on real batches, the 4th call captures
        graph_step = tt.CudaGraphTrainStep(model, opt, CAT, cfg["rows"], cfg["batch"], dev, warmup_steps=3)
        dev_raw = [(b.ids.to(dev), b.labels.to(dev)) for b in raw]
        for i in range(3):
            graph_step(*dev_raw[i % nb])
        l_before = lib.tt_kernel_launch_count()
        graph_step(*dev_raw[3 % nb])          # capture (+ first replay)
        launches_per_step = int(lib.tt_kernel_launch_count() - l_before)
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for i in range(args.warmup):
        graph_step(*dev_raw[i % nb]) if use_graph else step(resident[i % nb])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.tt_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        graph_step(*dev_raw[i % nb]) if use_graph else step(resident[i % nb])
    e1.record()
    barrier()
    launches = launches_per_step * args.steps if use_graph else lib.tt_kernel_launch_count() - l0
    ms_value = e0.elapsed_time(e1) / args.steps

    # ---- e2e: pinned host -> device -> step -> loss to host
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
            last = float(pipe.progress(it)[0])  # .item(): device -> host read of the loss
        t1.record()
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = t0.elapsed_time(t1) / args.steps

    # ---- instrumented pass: per-library-call CUDA events
    N.enable_timing(True)
    for i in range(min(args.steps, 5)):
        step(resident[i % nb])
    torch.cuda.synchronize()
    per_call = N.timing_summary()
    N.enable_timing(False)

    # ---- EBC lookup alone (BASELINE's second metric): 40 back-to-back lookups over rotating batches, one event pair
    # around the loop, so the figure is the kernel's duration and not one launch + event overhead (the lookup is a
    # 20 us kernel at this size; tables are 5 GB of random rows, nothing is L2-resident between launches)
    ebc_only_ms = None
    if world == 1:
        from ctypes import byref
        ebc_mod = model.module.two_tower.ebc
        kj = [resident[i].sparse_features for i in range(nb)]
        plan, total_dim = ebc_mod._build_plan(tuple(kj[0].keys()), cfg["batch"], with_state=False)   # built once: the loop is launches only
        vals = [k.values().contiguous() for k in kj]
        offs = [k.offsets().to(torch.int32).contiguous() for k in kj]
        pooled = torch.empty(cfg["batch"], total_dim, dtype=torch.float32, device=dev)
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
        ebc_only_ms = a0.elapsed_time(a1) / 40

    t = torch.tensor([ms_value, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e = t.tolist()
    if rank != 0:
        leave(world)
        return
    pk = peaks()
    B = cfg["batch"]
    uniq = [int(torch.unique(resident[0].sparse_features[c].values()[:B]).numel()) for c in CAT]
    fwd_bytes, bwd_bytes = algorithmic_bytes(cfg, uniq, world)
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
        roof = {"kernel": "in-batch softmax: tc_softmax_fwd_kernel + tc_softmax_bwd_fused_kernel (tcgen05)", "bound": "tensor",
                "achieved": round(ach, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4),
                # dram__bytes_read+write per step of the two launches, from the ncu --set full capture
                # profiles/r01_ncu_softmax_final2_summary.txt (17.3 MB forward + 50.8 MB one-pass backward)
                "traffic": 6.82e7, "peak_source": pk["source"] + " (sustained bf16)", "ms": round(sm_ms, 4),
                "flops_credited": "6*B*B*d (recomputation of S in the backward is not credited)",
                "note": "at d=64 the forward is MUFU(ex2)-bound (1 ex2 per 64 MACs) and the one-pass backward is bound by shared-memory "
                        "operand bandwidth of its N=64 tcgen05.mma (ncu: tc+lsu smem wavefronts ~90%); see DESIGN.md section 4"}
    line = {
        "metric": "two-tower train samples/s", "value": round(world * B / (ms_value * 1e-3), 1), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_value, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tower + logits GEMMs (fp32 accumulate, fp32 master weights) + f32 embeddings/optimizers", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] on %d GPU(s): 2 tables 10M x 64 fp32, per-rank batch 65536, MLP 64-128-64, "
                               "in-batch softmax, fused row-wise Adagrad, Adam" % world,
                   "per_rank_batch": B, "global_batch": B * world, "cuda_graph": bool(use_graph),
                   "exchange": (cfg["exchange"] if world > 1 else None),
                   "sharding": (None if world == 1 else ("row_wise" if (world > len(CAT) and cfg["exchange"] == "peer") else "table_wise")), "l2": "tables 5.12 GB >> 126 MB L2, random ids; no flush needed"},
        "e2e": {"value": round(world * B / (ms_e2e * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": raw[0].nbytes(), "d2h_bytes_per_step": 4, "last_loss": last, "api": e2e_api},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels,
        "calls_ms": {k: round(v["ms"], 4) for k, v in sorted(per_call.items(), key=lambda kv: -kv[1]["ms"])},
        "ebc_lookup_gbs": kernels.get("tt_ebc_forward", kernels.get("tt_ebc_forward_peer", {})).get("achieved"),
    }
    if ebc_only_ms:
        line["ebc_lookup"] = {"gbs": round(fwd_bytes / (ebc_only_ms * 1e-3) / 1e9, 1), "us": round(ebc_only_ms * 1e3, 2), "peak_gbs": pk["hbm"],
                              "frac": round(fwd_bytes / (ebc_only_ms * 1e-3) / 1e9 / pk["hbm"], 4), "bytes": int(fwd_bytes),
                              "how": "40 back-to-back tt_ebc_forward launches over 4 rotating batches, CUDA events around the loop"}
        line["ebc_lookup_gbs"] = line["ebc_lookup"]["gbs"]
    if world == 1:
        line["retrieval"] = retrieval_probe(dev)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, steps=2, warmup=1)
    print(json.dumps(line))
    leave(world)


def leave(world):
    """Multi-rank exit: the captured graphs hold NCCL kernels and tearing the communicator down under them can
    block, so flush and exit the process without the (purely cosmetic) process-group teardown."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def retrieval_probe(dev, n_items=2_000_000, n_queries=16384, d=64, k=100):
    """Secondary number (BASELINE configs[4] shape, scaled to one GPU and a few hundred ms): top-100 by
    dot product over a resident bf16 corpus, tcgen05 scoring with the top-k fused in the epilogue."""
    import two_tower_recommender_model_b200 as tt
    g = torch.Generator(device=dev).manual_seed(7)
    items = torch.randn(n_items, d, device=dev, generator=g)
    queries = torch.randn(n_queries, d, device=dev, generator=g)
    index = tt.BruteForceIndex(items, precision="bf16")
    index.search(queries[:1024], k)
    index.search(queries, k)           # same shape as the timed call: workspace allocation / launch attributes settled
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    index.search(queries, k)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return {"queries_per_s": round(n_queries / (ms * 1e-3), 1), "ms": round(ms, 3), "items": n_items, "queries": n_queries,
            "k": k, "d": d, "tflops": round(2.0 * n_queries * n_items * d / (ms * 1e-3) / 1e12, 1), "dtype": "bf16 scoring, f32 accumulate"}


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_baseline(cfg, steps, warmup, sample_batch=8192):
    """Oracle port of the reference's unsharded CPU path (dense [R,D] embedding gradient +
    row-wise Adagrad over the whole table each step, as nn.EmbeddingBag + a grad hook do)."""
    import oracle
    from oracle.ebc import TableSpec
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    specs = [TableSpec(f"t_{c}", cfg["rows"][i], cfg["dim"], [c]) for i, c in enumerate(CAT)]
    m = oracle.OracleTwoTower(specs, cfg["layers"], loss="softmax" if cfg["loss"] != "bce" else "bce",
                              sparse_lr=cfg["sparse_lr"], dense_lr=cfg["dense_lr"], seed=0)
    g = torch.Generator().manual_seed(0)
    Bs = min(sample_batch, cfg["batch"])
    times = []
    for i in range(warmup + steps):
        vals = torch.cat([torch.randint(1, r, (Bs,), generator=g) for r in cfg["rows"]])
        lens = torch.ones(2 * Bs, dtype=torch.int32)
        y = torch.randint(0, 2, (Bs,), generator=g, dtype=torch.int32)
        t0 = time.perf_counter()
        m.train_step(CAT, vals, lens, y)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": round(Bs / sec, 1), "unit": "samples/s", "cores": cores, "kind": "port", "ms_per_step": round(sec * 1e3, 2),
            "sample": f"{Bs} of the {cfg['batch']}-sample batch per step (in-batch negatives = {Bs}), full 10M-row tables, "
                      f"{steps} timed steps after {warmup} warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    cfg = dict(CFG2)
    steps = max(1, min(args.steps, 3))
    cb = cpu_baseline(cfg, steps=steps, warmup=1)
    line = {"impl": "reference", "metric": "two-tower train samples/s", "value": cb["value"], "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] (CPU port of the reference's unsharded TorchRec path; torchrec/fbgemm "
                                   "are not installable here)", "per_rank_batch": cfg["batch"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default=os.environ.get("TT_EXCHANGE", "peer"), choices=["nccl", "peer"],
                    help="N>1: table-wise output exchange by NCCL all-to-all, or fused into the lookup kernels over NVLink peer memory")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly (N>1: through TrainPipelineSparseDist) instead of replaying a CUDA graph")
    args = ap.parse_args()
    # a wedged collective / capture must not hold the box: the default run takes well under two minutes
    wd = threading.Timer(float(os.environ.get("TT_BENCH_WATCHDOG_S", "900")), lambda: (sys.stderr.write("bench.py: watchdog expired\n"), os._exit(3)))
    wd.daemon = True
    wd.start()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
