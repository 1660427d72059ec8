"""Static checks over EVERY Python file of the repo -- in particular over the code that only runs on a GPU box (the `-m gpu`
tests, bench.py's timed blocks, tools/), which the CPU suite never executes:

* no name is read that is not bound in an enclosing scope (a NameError waiting on the GPU box);
* every call into the C ABI -- ``N.call("tt_x", ...)`` and ``lib.tt_x(...)`` -- passes exactly as many arguments as the ctypes
  signature in ``_native.SIGNATURES`` has (which tests/test_host_logic.py holds to include/tt_b200.h)."""
import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _python_files():
    pats = ["*.py", "tests/*.py", "tests/golden/*.py", "tools/*.py", "oracle/*.py", "two_tower_recommender_model_b200/**/*.py"]
    out = []
    for p in pats:
        out += glob.glob(os.path.join(ROOT, p), recursive=True)
    return sorted(set(out))


def _bound_names(node):
    """Names bound directly in the scope of `node` (module, class, function or lambda), not in nested scopes."""
    names = set()

    class Binder(ast.NodeVisitor):
        def visit_FunctionDef(self, n):
            names.add(n.name)
        visit_AsyncFunctionDef = visit_FunctionDef

        def visit_ClassDef(self, n):
            names.add(n.name)

        def visit_Lambda(self, n):
            pass

        def visit_Import(self, n):
            for a in n.names:
                names.add((a.asname or a.name).split(".")[0])

        def visit_ImportFrom(self, n):
            for a in n.names:
                names.add(a.asname or a.name)

        def visit_Name(self, n):
            if isinstance(n.ctx, (ast.Store, ast.Del)):
                names.add(n.id)

        def visit_Global(self, n):
            names.update(n.names)
        visit_Nonlocal = visit_Global

        def visit_ExceptHandler(self, n):
            if n.name:
                names.add(n.name)
            self.generic_visit(n)

        def visit_arg(self, n):
            names.add(n.arg)

    body = node.body if isinstance(node.body, list) else [node.body]
    b = Binder()
    for stmt in body:
        b.visit(stmt)            # comprehension targets count for the enclosing scope: looser than Python, never stricter
    if hasattr(node, "args"):
        a = node.args
        for x in a.posonlyargs + a.args + a.kwonlyargs:
            names.add(x.arg)
        for x in (a.vararg, a.kwarg):
            if x:
                names.add(x.arg)
    return names


class _Unbound(ast.NodeVisitor):
    def __init__(self):
        self.scopes = [set(dir(builtins)) | {"__file__", "__name__", "__doc__"}]
        self.problems = []

    def _scoped(self, node, inner):
        self.scopes.append(_bound_names(node))
        for n in inner:
            self.visit(n)
        self.scopes.pop()

    def visit_Module(self, n):
        self._scoped(n, n.body)

    def visit_FunctionDef(self, n):
        for d in n.decorator_list + n.args.defaults + [x for x in n.args.kw_defaults if x]:
            self.visit(d)
        self._scoped(n, n.body)
    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_Lambda(self, n):
        self._scoped(n, [n.body])

    def visit_ClassDef(self, n):
        for b in n.bases + n.decorator_list:
            self.visit(b)
        self._scoped(n, n.body)

    def visit_Name(self, n):
        if isinstance(n.ctx, ast.Load) and not any(n.id in s for s in self.scopes):
            self.problems.append((n.lineno, n.id))


def test_no_unbound_names_in_any_python_file():
    files = _python_files()
    assert len(files) > 60 and any(f.endswith("test_gpu_tc.py") for f in files) and any(f.endswith("bench.py") for f in files)
    problems = []
    for fn in files:
        with open(fn) as f:
            tree = ast.parse(f.read(), fn)
        v = _Unbound()
        v.visit(tree)
        problems += [f"{os.path.relpath(fn, ROOT)}:{line}: {name}" for line, name in v.problems]
    assert not problems, "names read but never bound:\n" + "\n".join(problems)


def test_the_checker_sees_an_unbound_name():
    v = _Unbound()
    v.visit(ast.parse("import os\ndef f(a):\n    b = a + 1\n    return os.path.join(b, missing)\n"))
    assert v.problems == [(4, "missing")]


def test_every_c_abi_call_site_matches_the_ctypes_signature():
    from two_tower_recommender_model_b200._native import SIGNATURES
    problems, called = [], set()
    for fn in _python_files():
        with open(fn) as f:
            tree = ast.parse(f.read(), fn)
        for n in ast.walk(tree):
            if not isinstance(n, ast.Call) or not isinstance(n.func, ast.Attribute) or any(isinstance(a, ast.Starred) for a in n.args):
                continue
            if n.func.attr == "call" and n.args and isinstance(n.args[0], ast.Constant) and str(n.args[0].value).startswith("tt_"):
                name, nargs = n.args[0].value, len(n.args) - 1
            elif n.func.attr.startswith("tt_"):
                name, nargs = n.func.attr, len(n.args)
            else:
                continue
            called.add(name)
            where = f"{os.path.relpath(fn, ROOT)}:{n.lineno}"
            if name not in SIGNATURES:
                problems.append(f"{where}: {name} is not an entry point")
            elif len(SIGNATURES[name][1]) != nargs:
                problems.append(f"{where}: {name} called with {nargs} arguments, its signature has {len(SIGNATURES[name][1])}")
    assert not problems, "\n".join(problems)
    assert len(called) >= 50          # the scan found the call sites (54 entry points, tt_adam_flat is for C hosts only)


def test_package_and_oracle_attributes_used_by_gpu_only_code_exist():
    """``tt.X.Y`` / ``oracle.X`` chains written in tests, bench.py, tools and the driver entry resolve to real attributes, and every
    ``from <project module> import name`` names something that exists (static attributes of modules and classes only)."""
    import importlib
    import sys
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    import two_tower_recommender_model_b200 as tt
    import two_tower_recommender_model_b200._native  # noqa: F401
    import two_tower_recommender_model_b200.functional  # noqa: F401
    roots = {"tt": tt, "oracle": oracle}
    module_type = type(os)
    problems = set()
    own = ("two_tower_recommender_model_b200", "oracle", "helpers", "run_configs", "bench", "test_reference_boundary", "test_gpu_multi")
    for fn in _python_files():
        rel = os.path.relpath(fn, ROOT)
        with open(fn) as f:
            tree = ast.parse(f.read(), fn)
        pkg = rel[:-3].replace(os.sep, ".").rsplit(".", 1)[0] if rel.startswith("two_tower_recommender_model_b200") else None
        if pkg is not None and rel.endswith("__init__.py"):
            pkg = rel[:-len("/__init__.py")].replace(os.sep, ".")
        for n in ast.walk(tree):
            if isinstance(n, ast.Attribute) and pkg is None:
                parts, base = [], n
                while isinstance(base, ast.Attribute):
                    parts.append(base.attr)
                    base = base.value
                if isinstance(base, ast.Name) and base.id in roots:
                    obj = roots[base.id]
                    for i, p in enumerate(reversed(parts)):
                        if not hasattr(obj, p):
                            if isinstance(obj, (module_type, type)):
                                problems.add(f"{rel}:{n.lineno}: {base.id}.{'.'.join(list(reversed(parts))[:i + 1])}")
                            break
                        obj = getattr(obj, p)
            elif isinstance(n, ast.ImportFrom):
                mod = n.module or ""
                if n.level:
                    if pkg is None:
                        continue
                    base_parts = pkg.split(".")
                    base_parts = base_parts[:len(base_parts) - (n.level - 1)]
                    mod = ".".join(base_parts + ([mod] if mod else []))
                if not mod.startswith(own):
                    continue
                try:
                    m = importlib.import_module(mod)
                except Exception as e:          # noqa: BLE001
                    problems.add(f"{rel}:{n.lineno}: cannot import {mod}: {e}")
                    continue
                for a in n.names:
                    if a.name != "*" and not hasattr(m, a.name):
                        try:
                            importlib.import_module(mod + "." + a.name)
                        except Exception:       # noqa: BLE001
                            problems.add(f"{rel}:{n.lineno}: {mod} has no {a.name}")
    assert not problems, "\n".join(sorted(problems))


def test_calls_into_the_package_bind_to_the_signatures():
    """Every call in tests / bench.py / tools / the driver entry whose callee resolves statically to a function or class of
    this package or of oracle/ (``tt.X(...)``, names brought in by ``from <project module> import``) binds to that callee's
    signature: positional count and keyword names.  ~800 call sites, most of them in code that only runs on a GPU."""
    import importlib
    import inspect
    import sys

    import torch
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    import two_tower_recommender_model_b200 as tt
    import two_tower_recommender_model_b200._native  # noqa: F401
    import two_tower_recommender_model_b200.functional  # noqa: F401
    own = ("two_tower_recommender_model_b200", "oracle", "helpers", "bench")
    module_type = type(os)

    def resolve(node, env):
        parts = []
        while isinstance(node, ast.Attribute):
            parts.append(node.attr)
            node = node.value
        if not isinstance(node, ast.Name) or node.id not in env:
            return None
        obj = env[node.id]
        for p in reversed(parts):
            if not isinstance(obj, (module_type, type)) or not hasattr(obj, p):
                return None
            obj = getattr(obj, p)
        return obj

    problems, checked = [], 0
    for fn in _python_files():
        rel = os.path.relpath(fn, ROOT)
        if rel.startswith(("two_tower_recommender_model_b200", "oracle")):
            continue                      # the package's own internals are executed by the CPU suite through stand-ins
        with open(fn) as f:
            tree = ast.parse(f.read(), fn)
        env = {"tt": tt, "oracle": oracle}
        for n in ast.walk(tree):
            if isinstance(n, ast.ImportFrom) and not n.level and (n.module or "").startswith(own):
                try:
                    m = importlib.import_module(n.module)
                except Exception:          # noqa: BLE001 -- reported by the import check above
                    continue
                for a in n.names:
                    if hasattr(m, a.name):
                        env[a.asname or a.name] = getattr(m, a.name)
        for n in ast.walk(tree):
            if not isinstance(n, ast.Call) or any(isinstance(a, ast.Starred) for a in n.args) or any(k.arg is None for k in n.keywords):
                continue
            obj = resolve(n.func, env)
            if obj is None or not callable(obj) or (isinstance(obj, type) and issubclass(obj, torch.autograd.Function)):
                continue
            try:
                sig = inspect.signature(obj)
            except (TypeError, ValueError):
                continue
            checked += 1
            try:
                sig.bind(*[None] * len(n.args), **{k.arg: None for k in n.keywords})
            except TypeError as e:
                problems.append(f"{rel}:{n.lineno}: {ast.unparse(n.func)}: {e}")
    assert not problems, "\n".join(problems)
    assert checked > 500


def test_unverified_gpu_tests_dry_run_on_cpu():
    """tests/dryrun_zz.py: the GPU tests that have not run on a GPU yet, executed on CPU with the device entry points replaced by
    the oracle (in a subprocess: the stand-ins are patched process-wide)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_zz.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    for done in ("checkpoint_resume dry run ok", "corpus_sharded dry run ok", "golden batch construction ok", "golden train steps ok",
                 "golden corpus ok", "golden raytune ok", "fbgemm vector ok"):
        assert done in r.stdout, done


def test_bench_single_gpu_flow_dry_run_on_cpu():
    """tests/dryrun_bench.py: bench.py's run_ours(args) itself (default variant: the step behind CudaGraphTrainStep) and time_block
    in its eager variant, on a tiny configuration, device entry points replaced by the oracle, CUDA events / streams / graphs by fakes.  Checked here: it runs through and the line it prints has
    the contract's keys and every side block of the N = 1 run."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_bench.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "bench dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    line = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith('{"metric"')))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "cpu_baseline_cfg1", "retrieval", "retrieval_large",
              "kernels", "calls_ms", "ebc_lookup"):
        assert k in line, k
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 2 * 64 * 8 + 64 * 4
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0 and "error" not in line["cpu_baseline_cfg1"]
    assert line["retrieval"]["queries_per_s"] > 0 and "error" not in line["retrieval_large"]
    assert line["roofline"]["bound"] == "tensor" and "explain_error" not in line and line["config"]["cuda_graph"] is True


def test_bench_multi_rank_flow_dry_run_on_cpu_world2_gloo():
    """tests/dryrun_bench_world2.py: bench.py's run_ours(args) itself on two gloo ranks -- parity checks (table-wise / row-wise,
    bf16-configured and fp32, eager and through CudaGraphTrainStep), the headline block, every side block, the exit path -- with
    the device entry points replaced by the oracle and the NCCL exchange standing in for the peer-memory one."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_bench_world2.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "bench world-2 dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    line = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith('{"metric"')))
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["config"]["sharding"] == ["table_wise"]
    assert [(p["mode"], p["precision"]) for p in line["parity"]] == [
        ("table_wise/nccl/eager", "bf16"), ("row_wise/nccl/eager", "bf16"), ("table_wise/nccl/eager", "fp32"), ("row_wise/nccl/eager", "fp32"),
        ("table_wise/nccl/cuda_graph", "fp32"), ("row_wise/nccl/cuda_graph", "fp32")]
    assert all(p["ok"] for p in line["parity"]) and "parity_failed" not in line
    assert line["e2e"]["d2h_bytes_per_step"] == 8 and line["e2e"]["h2d_bytes_per_step"] == 2 * line["e2e"]["h2d_bytes_per_step_per_rank"]
    for k in ("strong_row_wise", "weak", "strong_global_negatives", "retrieval", "cfg4_sharded", "cfg3_row_wise"):
        assert k in line, k
    assert "skipped" in line["cfg3_row_wise"]          # configs[2] sharded needs the peer-memory exchange; its own dry run is below


def test_bench_multi_rank_run_with_a_failed_row_wise_parity_check_dry_run():
    """The same two-rank dry run with the row-wise parity results forced to `ok: false`: the run still ends with rc 0 and the
    table-wise headline; `parity_failed` names the sharding, the row-wise block is recorded as skipped, the others are timed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_bench_world2.py")], capture_output=True, text=True, cwd=ROOT,
                       timeout=900, env=dict(os.environ, DRYRUN_FAIL_ROW_WISE="1"))
    assert r.returncode == 0 and "bench world-2 dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    line = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith('{"metric"')))
    assert line["parity_failed"] == ["row_wise"] and line["value"] > 0 and line["config"]["sharding"] == ["table_wise"]
    assert "skipped" in line["strong_row_wise"] and "value" in line["weak"] and "retrieval" in line


def test_driver_smoke_entry_dry_run_on_cpu():
    """tests/dryrun_smoke.py: ``__graft_entry__.smoke()`` -- what the driver runs on the GPU box before the bench -- with the device
    entry points replaced by the oracle: its own flow and its comparisons against the oracle's numbers."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_smoke.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "smoke dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    for leg in ("smoke[bce]", "smoke[in_batch_softmax]", "smoke[bf16 tcgen05]", "top-100 indices bit-exact"):
        assert leg in r.stdout, leg


def test_config3_and_config4_sharded_blocks_dry_run_on_cpu_world2_gloo():
    """tools/run_configs.py::config3_sharded / config4_sharded -- configs[2] row-wise sharded and configs[3] under the planner's
    sharding (row-wise Adam, three-layer towers), the last two side blocks of an N > 1 bench run -- on two gloo ranks with tiny
    tables, fp32 towers and the NCCL exchange: its model construction, the multi-hot batches in their
    fixed-capacity buffers, the CudaGraphTrainStep.step_kjt loop, the max over ranks and the block's result dict."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_bench_world2.py")], capture_output=True, text=True, cwd=ROOT,
                       timeout=900, env=dict(os.environ, DRYRUN_CFG3_SHARDED="1"))
    assert r.returncode == 0 and "bench world-2 dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    out = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith('{"config": 3')))
    assert out["sharding"] == ["row_wise"] and out["cuda_graph"] is True and out["global_batch"] == 64 and out["value"] > 0
    out4 = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith('{"config": 4')))
    assert out4["sharding"] == ["table_wise"] and out4["cuda_graph"] is True and out4["global_batch"] == 64 and out4["value"] > 0


def test_two_gpu_training_worker_dry_run_on_cpu_world2_gloo():
    """tests/dryrun_multi.py: the training worker of tests/test_gpu_multi.py on two gloo ranks with the oracle-backed lookup in
    place of the kernels -- in the modes that have not run on GPUs yet (column_wise, data_parallel, data_parallel_dense) and two
    that have; the worker's own assertions (losses of every step, gathered tables against the oracle) are the check."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_multi.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "multi dry run ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    for mode in ("table_wise", "row_wise", "column_wise", "data_parallel", "data_parallel_dense"):
        assert f"mode {mode} ok" in r.stdout, mode
