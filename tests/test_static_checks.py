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
