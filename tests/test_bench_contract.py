"""bench.py's host logic, on CPU: the JSON line carries every key the bench contract names, the side blocks of an
N > 1 run are recorded one by one (a failing block becomes an `error` entry, the others stay), and a run cut by the
watchdog still prints the headline line once that block is measured.  No kernel runs here: `time_block` and the
retrieval probe are replaced by fakes with the shape of their real results."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=10, warmup=3, impl="ours", no_cpu_baseline=True, no_other_configs=True, exchange="peer",
             parity_only=False, parity_graph=False, no_graph=False)
    d.update(kw)
    return argparse.Namespace(**d)


def _block(ms=2.8, B=65536, sharding=None):
    return {"ms_value": ms, "ms_e2e": ms * 1.02, "launches": 470, "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 9},
            "per_call": {"tt_inbatch_softmax_forward_bf16": {"ms": 0.9, "calls": 5}, "tt_inbatch_softmax_backward_bf16": {"ms": 1.3, "calls": 5},
                         "tt_ebc_forward": {"ms": 0.03, "calls": 5}, "tt_ebc_backward_fused": {"ms": 0.08, "calls": 5}},
            "ebc_only_ms": 0.0125, "uniq": [65300, 65310], "last_loss": 11.09, "e2e_api": "CudaGraphTrainStep", "h2d": 2 * B * 8 + B * 4,
            "sharding": sharding, "cuda_graph": True, "batch": B}


def test_headline_line_has_every_contract_key():
    line = bench.headline(_args(), dict(bench.CFG2), _block(), bench.peaks(), 1, 65536, "BASELINE configs[1] on 1 GPU", "strong", [])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in line, k
    assert line["metric"] == "two-tower train samples/s" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None                 # BASELINE.md publishes no number for this metric
    assert "workload" in line["config"] and "model" not in line["config"]
    assert abs(line["value"] - 65536 / 2.8e-3) < 1.0 and line["ms_per_step"] == 2.8
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 65536 * 8 + 65536 * 4 and line["e2e"]["value"] < line["value"]
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["achieved"] - 6.0 * 65536 ** 2 * 64 / 2.2e-3 / 1e12) < 0.1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert line["gpu_launches"] > 0
    assert line["ebc_lookup"]["frac"] == pytest.approx(line["ebc_lookup"]["gbs"] / line["ebc_lookup"]["peak_gbs"], abs=1e-3)
    json.dumps(line)


def test_side_blocks_are_recorded_one_by_one_and_a_failing_block_does_not_take_the_others(monkeypatch):
    calls = []

    def fake_time_block(cfg, B, dev, rank, world, local, args, sharding, exchange, lib, with_kernels):
        calls.append((B, sharding, cfg.get("negatives", "local")))
        if sharding == "row_wise":
            raise RuntimeError("peer buffer rendezvous failed")
        return _block(ms=0.5, B=B, sharding=[sharding or "table_wise"])

    monkeypatch.setattr(bench, "time_block", fake_time_block)
    monkeypatch.setattr(bench, "retrieval_probe_sharded", lambda dev, rank, world: {"queries_per_s_total": 1.0})
    line = {"value": 1.0}
    bench.side_blocks(_args(gpus=4), dict(bench.CFG2), None, 0, 4, 0, None, 65536, line)
    assert calls == [(16384, "row_wise", "local"), (65536, None, "local"), (16384, "table_wise", "global")]
    assert "error" in line["strong_row_wise"] and "rendezvous" in line["strong_row_wise"]["error"]
    assert line["weak"]["value"] == pytest.approx(4 * 65536 / 0.5e-3, rel=1e-6) and line["weak"]["global_batch"] == 4 * 65536
    assert line["strong_global_negatives"]["value"] == pytest.approx(65536 / 0.5e-3, rel=1e-6)
    assert line["retrieval"] == {"queries_per_s_total": 1.0}
    assert bench._PARTIAL["stage"] == "done"
    # the other ranks run the same blocks and record nothing
    calls.clear()
    bench.side_blocks(_args(gpus=4), dict(bench.CFG2), None, 3, 4, 3, None, 65536, None)
    assert len(calls) == 3


def test_a_sharding_that_failed_its_parity_check_is_not_timed(monkeypatch):
    """N > 1: a failed row-wise parity check leaves the table-wise headline standing; the blocks that would run row-wise
    (strong_row_wise always, weak once the ranks outnumber the two tables) are recorded as skipped, on every rank alike."""
    calls = []

    def fake_time_block(cfg, B, dev, rank, world, local, args, sharding, exchange, lib, with_kernels):
        calls.append((B, sharding, cfg.get("negatives", "local")))
        return _block(ms=0.5, B=B, sharding=[sharding or "table_wise"])

    monkeypatch.setattr(bench, "time_block", fake_time_block)
    monkeypatch.setattr(bench, "retrieval_probe_sharded", lambda dev, rank, world: {"queries_per_s_total": 1.0})
    for world, weak_timed in ((8, False), (2, True)):
        calls.clear()
        line = {"value": 1.0}
        bench.side_blocks(_args(gpus=world, parity_failed=["row_wise"]), dict(bench.CFG2), None, 0, world, 0, None, 65536, line)
        assert "skipped" in line["strong_row_wise"] and "row_wise" in line["strong_row_wise"]["skipped"]
        assert ("skipped" not in line["weak"]) == weak_timed
        assert "value" in line["strong_global_negatives"] and line["retrieval"] == {"queries_per_s_total": 1.0}
        assert [c[1] for c in calls] == ([None, "table_wise"] if weak_timed else ["table_wise"])
    # and the headline line names the failed sharding next to the parity records
    parity = [{"mode": "table_wise/peer/eager", "ok": True}, {"mode": "row_wise/peer/eager", "ok": False}]
    line = bench.headline(_args(gpus=8, parity_failed=["row_wise"]), dict(bench.CFG2), _block(B=8192, sharding=["table_wise"]), bench.peaks(), 8, 65536,
                          "configs[1] on 8 GPUs", "strong", parity)
    assert line["parity_failed"] == ["row_wise"] and line["parity"] == parity and line["value"] > 0


@pytest.mark.parametrize("have_line", [True, False])
def test_watchdog_prints_the_headline_once_it_exists(have_line, tmp_path):
    code = (
        "import sys, json; sys.path.insert(0, %r); import bench\n"
        "bench._PARTIAL['stage'] = 'weak'\n"
        "bench._PARTIAL['line'] = {'metric': 'two-tower train samples/s', 'value': 1.5} if %r else None\n"
        "bench.watchdog()\n" % (ROOT, have_line))
    env = dict(os.environ, RANK="0", MASTER_PORT="0%d" % os.getpid())
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    marker = "/tmp/tt_bench_partial_0%d" % os.getpid()
    if have_line:
        assert p.returncode == 0, p.stderr
        line = json.loads(p.stdout.strip().splitlines()[-1])
        assert line["value"] == 1.5 and line["incomplete"]["cut_block"] == "weak"
        # a non-zero rank learns the outcome from the marker
        p2 = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(env, RANK="1"), timeout=300)
        assert p2.returncode == 0 and p2.stdout.strip() == ""
        os.remove(marker)
    else:
        assert p.returncode == 3 and p.stdout.strip() == ""
        assert not os.path.exists(marker)
    assert "watchdog expired on rank 0 during block 'weak'" in p.stderr


def test_reference_arm_line_shape(monkeypatch, capsys):
    """`--impl reference`: the oracle port timed on the host cores; same metric / unit / config keys as our arm."""
    monkeypatch.setattr(bench, "cpu_baseline", lambda cfg, steps, warmup, sample_batch=None, budget_s=None: {
        "value": 8000.0, "unit": "samples/s", "cores": 16, "kind": "port", "ms_per_step": 8192.0, "sample": "x",
        "steps": steps, "warmup": warmup})
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "1")
    bench.run_reference(_args(impl="reference", steps=4, warmup=1))
    line = json.loads(capsys.readouterr().out.strip())
    assert line["impl"] == "reference" and line["metric"] == "two-tower train samples/s" and line["unit"] == "samples/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 4 and line["warmup"] == 1 and line["higher_is_better"] is True
    # ranks other than 0 print nothing
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    bench.run_reference(_args(impl="reference"))
    assert capsys.readouterr().out == ""


def test_reference_arm_honours_steps_and_warmup_within_its_budget():
    """The CPU arm runs the steps / warm-up steps it is asked for when they fit its time budget (so that the driver's
    `steps_match` / `warmup_match` hold), fewer otherwise: timed steps first, at least one."""
    assert bench.bounded_steps(20, 5, 240.0 - 8.0, 8.0, done=1) == (20, 5)          # 16 host cores, ~8 s per step
    assert bench.bounded_steps(20, 5, 240.0 - 21.0, 21.0, done=1) == (11, 0)        # a slower host: fewer steps, no warm-up
    assert bench.bounded_steps(20, 5, 0.0, 300.0, done=1) == (1, 0)                 # one step already overran: it is the result
    assert bench.bounded_steps(4, 3, 45.0, 10.0, done=2) == (4, 2)
    tiny = dict(rows=[50, 40], dim=8, layers=[16, 8], batch=32, loss="in_batch_softmax", sparse_lr=0.01, dense_lr=0.001)
    full = bench.cpu_baseline(tiny, steps=3, warmup=2, budget_s=1000.0)
    assert (full["steps"], full["warmup"]) == (3, 2) and full["value"] > 0 and "3 timed step(s) after 2 warm-up" in full["sample"]
    cut = bench.cpu_baseline(tiny, steps=3, warmup=2, budget_s=0.0)
    assert (cut["steps"], cut["warmup"]) == (1, 0)
    plain = bench.cpu_baseline(tiny, steps=1, warmup=1)                              # our arm's `cpu_baseline` block: no budget
    assert (plain["steps"], plain["warmup"]) == (1, 1)


def test_a_wedged_side_block_is_cut_at_its_own_limit():
    """The polling watchdog: the run's limit is far away, the side block in flight has its own short one -- the process ends
    there, rank 0 prints the headline with the block named; a finished block (stage("done")) disarms it."""
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "bench._PARTIAL['line'] = {'metric': 'two-tower train samples/s', 'value': 2.5}\n"
        "bench.start_watchdog(300.0)\n"
        "bench.stage('strong_row_wise', 60.0); time.sleep(1.2); bench.stage('weak', 1.0)\n"
        "time.sleep(30)\n"
        "print('not reached')\n" % ROOT)
    env = dict(os.environ, RANK="0", MASTER_PORT="1%d" % os.getpid())
    t0 = __import__("time").time()
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    took = __import__("time").time() - t0
    marker = "/tmp/tt_bench_partial_1%d" % os.getpid()
    if os.path.exists(marker):
        os.remove(marker)
    assert p.returncode == 0 and "not reached" not in p.stdout and took < 25, (p.returncode, took, p.stderr[-300:])
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["value"] == 2.5 and line["incomplete"]["cut_block"] == "weak"
    code_ok = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "bench.start_watchdog(300.0)\n"
        "bench.stage('weak', 1.0); bench.stage('done'); time.sleep(3)\n"
        "print('finished')\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code_ok], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "finished"



def test_single_gpu_side_blocks_that_raise_become_error_entries(monkeypatch, capsys):
    """N = 1: the CPU baselines, the retrieval probes and the cfg3 / cfg4 blocks run AFTER the headline is measured; one that
    raises (host out of memory, a device fault) is recorded as an `error` entry and the line -- headline included -- still prints."""
    import types
    calls = []

    def cpu_fail(cfg, steps, warmup, sample_batch=None, budget_s=None):
        raise MemoryError("host out of memory")

    def probe(dev, n_items=2_000_000, n_queries=16384, d=64, k=100):
        calls.append(n_items)
        if n_items > 2_000_000:
            raise RuntimeError("CUDA error: an illegal memory access was encountered")
        return {"queries_per_s": 1.0}

    fake_cfgs = types.ModuleType("run_configs")
    fake_cfgs.config3 = lambda: {"ms_per_step": 4.4}
    fake_cfgs.config4 = lambda: (_ for _ in ()).throw(RuntimeError("out of memory"))
    monkeypatch.setitem(sys.modules, "run_configs", fake_cfgs)
    monkeypatch.setattr(bench, "cpu_baseline", cpu_fail)
    monkeypatch.setattr(bench, "cpu_baseline_cfg1", lambda: {"value": 96000.0})
    monkeypatch.setattr(bench, "retrieval_probe", probe)
    monkeypatch.setattr(bench.torch.cuda, "empty_cache", lambda: None)
    line = bench.headline(_args(), dict(bench.CFG2), _block(), bench.peaks(), 1, 65536, "BASELINE configs[1] on 1 GPU", "strong", [])
    bench.finish(_args(no_cpu_baseline=False, no_other_configs=False), dict(bench.CFG2), None, 1, line)
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert out["value"] == line["value"] and out["roofline"]["frac"] > 0
    assert "MemoryError" in out["cpu_baseline"]["error"] and out["cpu_baseline_cfg1"] == {"value": 96000.0}
    assert out["retrieval"] == {"queries_per_s": 1.0} and "illegal memory access" in out["retrieval_large"]["error"]
    assert out["cfg3"] == {"ms_per_step": 4.4} and "out of memory" in out["cfg4"]["error"]
    assert calls == [2_000_000, 10_000_000] and bench._PARTIAL["stage"] == "done"


def test_sigterm_from_torchrun_still_gets_the_headline_out():
    """When one rank dies torchrun sends the others SIGTERM.  Rank 0 -- possibly waiting inside a collective, where no Python
    signal handler would run -- still prints the headline line it holds (the signal reaches the watchdog thread through a
    wake-up pipe) and exits 0; a rank that holds no line exits 3.  Either way the process ends at once."""
    import signal
    import time
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "bench._PARTIAL['line'] = %s\n"
        "bench.start_watchdog(300.0); bench.stage('cfg3_row_wise', 200.0)\n"
        "print('ready', flush=True)\n"
        "time.sleep(60)\n"
        "print('not reached')\n")
    for have_line in (True, False):
        env = dict(os.environ, RANK="0", MASTER_PORT="2%d" % os.getpid())
        p = subprocess.Popen([sys.executable, "-c", code % (ROOT, "{'metric': 'two-tower train samples/s', 'value': 7.5}" if have_line else "None")],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        assert p.stdout.readline().strip() == "ready"
        t0 = time.time()
        p.send_signal(signal.SIGTERM)
        out, err = p.communicate(timeout=60)
        assert time.time() - t0 < 20 and "not reached" not in out, err[-300:]
        marker = "/tmp/tt_bench_partial_2%d" % os.getpid()
        if os.path.exists(marker):
            os.remove(marker)
        if have_line:
            line = json.loads(out.strip().splitlines()[-1])
            assert p.returncode == 0 and line["value"] == 7.5 and line["incomplete"]["cut_block"] == "cfg3_row_wise"
            assert "SIGTERM" in line["incomplete"]["reason"]
        else:
            assert p.returncode == 3 and out.strip() == ""
