"""CPU: host-side logic of the package (containers, planner, tagging, shim) and the C-ABI
library's presence -- no compute calls."""
import ctypes
import os
import re

import pytest
import torch

import two_tower_recommender_model_b200 as tt
from two_tower_recommender_model_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from two_tower_recommender_model_b200.build import build_library
    build_library()
    lib = N.load()
    assert lib.tt_abi_version() == N.TT_ABI_VERSION and lib.tt_build_arch() == b"sm_100a"
    header = open(os.path.join(ROOT, "include", "tt_b200.h")).read()
    declared = set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    raw = ctypes.CDLL(N.library_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in tt_b200.h but not exported"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(N.EbcPlan) == 16 + 8 + 32 * (8 * 5 + 4 * 5)
    assert ctypes.sizeof(N.SparseOptimizer) == 40   # 8 x 4 bytes + the device step pointer


def test_no_cpu_fallback():
    cfgs = [tt.EmbeddingBagConfig(name="t", embedding_dim=8, num_embeddings=10, feature_names=["a"])]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("cpu"))
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(["a"], torch.tensor([1, 2]), torch.tensor([1, 1], dtype=torch.int32))
    with pytest.raises(N.NativeLibraryError):
        ebc(kjt)
    with pytest.raises(N.NativeLibraryError):
        tt.MLP(8, [4])(torch.zeros(2, 8))
    with pytest.raises(RuntimeError):
        tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))(kjt)


def test_kjt_container_cpu():
    keys = ["user_id", "product_id"]
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(keys, torch.tensor([1, 3, 10, 4]), torch.tensor([1, 0, 1, 1, 1, 0], dtype=torch.int32))
    assert kjt.keys() == keys and kjt.stride() == 3 and kjt.length_per_key() == [2, 2] and kjt.offset_per_key() == [0, 2, 4]
    assert kjt.offsets().tolist() == [0, 1, 1, 2, 3, 4, 4] and kjt.offsets().dtype == torch.int32
    d = kjt.to_dict()
    assert d["product_id"].values().tolist() == [10, 4] and d["product_id"].lengths().tolist() == [1, 1, 0]
    assert d["product_id"].offsets().tolist() == [0, 1, 2, 2]
    assert kjt["user_id"].values().tolist() == [1, 3]
    p = kjt.permute([1, 0])
    assert p.keys() == ["product_id", "user_id"] and p.values().tolist() == [10, 4, 1, 3] and p.lengths().tolist() == [1, 1, 0, 1, 0, 1]
    a, b = kjt.split([1, 1])
    assert a.keys() == ["user_id"] and b.values().tolist() == [10, 4]
    # the ctor form of 03_model_training.py:1081-1085
    k2 = tt.create_keyed_jagged_tensor(4, keys, "product_id", device="cpu")
    assert k2.values().tolist() == [0, 1, 2, 3] and k2.lengths().tolist() == [0] * 4 + [1] * 4 and k2.length_per_key() == [0, 4]
    with pytest.raises(ValueError):
        tt.create_keyed_jagged_tensor(4, keys, "nope", device="cpu")


def test_keyed_tensor():
    kt = tt.KeyedTensor(["a", "b", "c"], [2, 3, 1], torch.arange(12.0).view(2, 6))
    assert kt["b"].tolist() == [[2, 3, 4], [8, 9, 10]] and kt.columns(["b", "c"]) == (2, 4) and kt.columns(["a", "c"]) == (-1, -1)
    assert set(kt.to_dict()) == {"a", "b", "c"}


def test_ebc_surface_meta_tags_and_state_dict_keys():
    from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
    cfgs = [tt.EmbeddingBagConfig(name=f"t_{c}", embedding_dim=8, num_embeddings=n, feature_names=[c])
            for c, n in (("user_id", 10), ("product_id", 20))]
    ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
    assert ebc.embedding_bag_configs()[0].embedding_dim == 8 and ebc.embedding_bag_configs()[1].feature_names == ["product_id"]
    assert all(p.device.type == "meta" for p in ebc.parameters())
    apply_optimizer_in_backward(tt.RowWiseAdagrad, ebc.parameters(), {"lr": 0.02})
    ebc.materialize(torch.device("cpu"))
    assert list(ebc.state_dict().keys()) == ["embedding_bags.t_user_id.weight", "embedding_bags.t_product_id.weight"]
    w = ebc.embedding_bags["t_user_id"].weight
    assert w.device.type == "cpu" and w._optimizer_classes[0] is tt.RowWiseAdagrad and w._optimizer_kwargs[0] == {"lr": 0.02}
    assert float(w.abs().max()) <= (1 / 10) ** 0.5 + 1e-7  # U(-1/sqrt(R), 1/sqrt(R))
    spec = ebc._sparse_optimizer_spec(advance_step=False)
    assert spec.kind == N.OPT_ROWWISE_ADAGRAD and abs(spec.lr - 0.02) < 1e-9 and abs(spec.eps - 1e-10) < 1e-16
    two = tt.TwoTower(ebc, [16, 8])
    names = set(dict(two.named_parameters()))
    assert "ebc.embedding_bags.t_user_id.weight" in names and "query_proj._mlp.1._linear.bias" in names
    opt = tt.KeyedOptimizerWrapper(dict(tt.TwoTowerTrainTask(two).named_parameters()), lambda p: torch.optim.Adam(p, lr=0.1))
    assert all("embedding_bags" not in k for k in opt.params) and opt.param_groups[0]["lr"] == 0.1


def test_two_tower_asserts_like_reference():
    mk = lambda n, d: tt.EmbeddingBagConfig(name=n, embedding_dim=d, num_embeddings=5, feature_names=[n])
    with pytest.raises(AssertionError):
        tt.TwoTower(tt.EmbeddingBagCollection(tables=[mk("a", 8)]), [4])
    with pytest.raises(AssertionError):
        tt.TwoTower(tt.EmbeddingBagCollection(tables=[mk("a", 8), mk("b", 4)]), [4])


def test_planner():
    from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints
    mk = lambda n, r: tt.EmbeddingBagConfig(name=n, embedding_dim=64, num_embeddings=r, feature_names=[n])
    ebc = tt.EmbeddingBagCollection(tables=[mk("t_a", 1000), mk("t_b", 5000), mk("t_c", 10)], device=torch.device("meta"))
    holder = torch.nn.ModuleDict({"ebc": ebc})
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(local_world_size=8, world_size=8, compute_device="cuda"), batch_size=64,
                                       storage_reservation=tt.HeuristicalStorageReservation(percentage=0.05)).plan(holder, tt.get_default_sharders())
    p = plan.plan["ebc"]
    assert list(p) == ["t_a", "t_b", "t_c"] and all(v.sharding_type == "table_wise" for v in p.values())
    assert p["t_b"].ranks == [0] and p["t_a"].ranks == [1] and p["t_c"].ranks == [2]  # largest first, least-loaded rank
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=8), constraints={"t_b": ParameterConstraints(sharding_types=["row_wise"])}).plan(holder)
    assert plan.plan["ebc"]["t_b"].sharding_type == "row_wise" and plan.plan["ebc"]["t_b"].block_size == 625
    # a table that cannot fit one GPU's budget goes row-wise by itself
    big = torch.nn.ModuleDict({"ebc": tt.EmbeddingBagCollection(tables=[mk("t_big", 100_000_000)], device=torch.device("meta"))})
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=4, hbm_cap=8 * 2 ** 30)).plan(big)
    assert plan.plan["ebc"]["t_big"].sharding_type == "row_wise"
    assert "t_big" in str(plan)
    # column-wise: as many shards as divide D (at most W), on distinct ranks, least-loaded first
    cw = torch.nn.ModuleDict({"ebc": tt.EmbeddingBagCollection(tables=[
        tt.EmbeddingBagConfig(name="t_x", embedding_dim=36, num_embeddings=5000, feature_names=["x"]), mk("t_y", 100)], device=torch.device("meta"))})
    plan = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=8), constraints={"t_x": ParameterConstraints(sharding_types=["column_wise"]),
                                                                                        "t_y": ParameterConstraints(sharding_types=["column_wise"])}).plan(cw)
    px, py = plan.plan["ebc"]["t_x"], plan.plan["ebc"]["t_y"]
    assert px.sharding_type == "column_wise" and px.ranks == [0, 1, 2, 3, 4, 5]           # 36 = 6 x 6; 7 and 8 do not divide it
    assert py.ranks == list(range(8)) and 36 % len(px.ranks) == 0 and 64 % len(py.ranks) == 0
    # data-parallel: a replica on every rank (dense gradient all-reduce), counted against every rank's budget
    dp = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=2), constraints={"t_x": ParameterConstraints(sharding_types=["data_parallel"])}).plan(cw)
    assert dp.plan["ebc"]["t_x"].sharding_type == "data_parallel" and dp.plan["ebc"]["t_x"].ranks == [0, 1]
    assert dp.plan["ebc"]["t_x"].compute_kernel == "dense" and dp.plan["ebc"]["t_y"].sharding_type == "table_wise"
    with pytest.raises(RuntimeError):                # a replica that does not fit one GPU
        tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=2, hbm_cap=5000 * 36 * 4), constraints={
            "t_x": ParameterConstraints(sharding_types=["data_parallel"])}).plan(cw)
    with pytest.raises(NotImplementedError):
        tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=2), constraints={"t_x": ParameterConstraints(sharding_types=["table_row_wise"])}).plan(cw)


def test_shim_resolves_reference_imports():
    tt.install_torchrec_shim()
    from torchrec.distributed import TrainPipelineSparseDist  # noqa: F401
    from torchrec.distributed.model_parallel import DistributedModelParallel, get_default_sharders  # noqa: F401
    from torchrec.inference.state_dict_transform import state_dict_gather, state_dict_to_device  # noqa: F401
    from torchrec.modules.embedding_configs import EmbeddingBagConfig  # noqa: F401
    from torchrec.modules.embedding_modules import EmbeddingBagCollection  # noqa: F401
    from torchrec.optim.keyed import KeyedOptimizerWrapper  # noqa: F401
    from torchrec.optim.rowwise_adagrad import RowWiseAdagrad  # noqa: F401
    from torchrec.sparse.jagged_tensor import KeyedJaggedTensor  # noqa: F401
    from torchrec.datasets.utils import Batch  # noqa: F401
    from torchrec.modules.mlp import MLP  # noqa: F401
    from torchrec.distributed.comm import get_local_size  # noqa: F401
    from torchrec.distributed.planner import EmbeddingShardingPlanner, Topology  # noqa: F401
    from torchrec.distributed.planner.storage_reservations import HeuristicalStorageReservation  # noqa: F401
    assert EmbeddingBagCollection is tt.EmbeddingBagCollection


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "two_tower_recommender_model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/tt_b200.h must compile as C99 on its own (no C++, no torch, no CUDA headers)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "tt_b200.h")
    r = subprocess.run(["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_planner_prefers_row_wise_when_ranks_outnumber_tables():
    """Two big tables on 8 ranks: table-wise would leave six ranks without embedding work -> row-wise; with as many
    tables as ranks, or tables too small to split, table-wise (largest first onto the least-loaded rank)."""
    import two_tower_recommender_model_b200 as tt
    from two_tower_recommender_model_b200.distributed.planner import ParameterConstraints

    def plan_for(world, rows, constraints=None):
        cfgs = [tt.EmbeddingBagConfig(name=f"t{i}", embedding_dim=64, num_embeddings=r, feature_names=[f"f{i}"]) for i, r in enumerate(rows)]
        ebc = tt.EmbeddingBagCollection(tables=cfgs, device=torch.device("meta"))
        p = tt.EmbeddingShardingPlanner(topology=tt.Topology(world_size=world), constraints=constraints).plan(torch.nn.ModuleDict({"ebc": ebc}))
        return {k: (v.sharding_type, v.ranks) for k, v in p.plan["ebc"].items()}

    p8 = plan_for(8, [10_000_000, 10_000_000])
    assert all(st == "row_wise" and ranks == list(range(8)) for st, ranks in p8.values())
    p2 = plan_for(2, [10_000_000, 5_000_000])
    assert p2["t0"] == ("table_wise", [0]) and p2["t1"] == ("table_wise", [1])
    small = plan_for(8, [134, 21])
    assert all(st == "table_wise" for st, _ in small.values())
    forced = plan_for(8, [10_000_000, 10_000_000], {"t0": ParameterConstraints(sharding_types=["table_wise"])})
    assert forced["t0"][0] == "table_wise" and forced["t1"][0] == "row_wise"


def test_c_host_links_and_calls_the_library(tmp_path):
    """A plain-C host (no Python, no torch) includes include/tt_b200.h, links libtt_b200.so and calls the entry points that
    need no device: ABI version, build architecture, workspace queries, and the argument check of a compute entry point
    (a NULL plan must come back as an error code with a message, not a crash).  This is the binding INTEGRATION.md section 3
    shows for a C / C++ host."""
    import shutil
    import subprocess
    from two_tower_recommender_model_b200 import _native as N
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    N.load()
    lib = N.library_path()
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "tt_b200.h"
int main(void) {
  if (tt_abi_version() != TT_ABI_VERSION) { printf("abi %d != %d\n", tt_abi_version(), TT_ABI_VERSION); return 1; }
  if (strcmp(tt_build_arch(), "sm_100a") != 0) { printf("arch %s\n", tt_build_arch()); return 2; }
  if (tt_kjt_offsets_workspace_bytes(1 << 20) == 0 || tt_ebc_backward_workspace_bytes(131072) == 0 ||
      tt_topk_workspace_bytes(1024, 100000, 100) == 0) { printf("workspace query returned 0\n"); return 3; }
  int rc = tt_ebc_forward(NULL, NULL, NULL, NULL, NULL);
  if (rc >= 0) { printf("NULL plan accepted (rc %d)\n", rc); return 4; }
  const char* msg = tt_last_error();
  if (msg == NULL || msg[0] == 0) { printf("no error text\n"); return 5; }
  printf("ok abi=%d arch=%s rc=%d msg=%s\n", tt_abi_version(), tt_build_arch(), rc, msg);
  return 0;
}
''')
    exe = tmp_path / "host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    lib, "-Wl,-rpath," + os.path.dirname(lib)], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith(f"ok abi={N.TT_ABI_VERSION} arch=sm_100a")


def test_entry_points_survive_null_and_zero_arguments():
    """Every entry point of the C ABI called with NULL pointers and zero sizes (in a child process: a crash must not take
    pytest down): workspace queries return a size, compute entry points come back with an error code and a message or with
    'nothing to do' -- none may crash.  (Found this way: tt_topk_workspace_bytes(0, ..) and
    tt_gemm_bf16_splitk_workspace_bytes(0, ..) divided by zero, i.e. an EMPTY query batch killed the process.)"""
    import json
    import subprocess
    import sys
    child = r'''
import ctypes, json, sys
sys.path.insert(0, %r)
from two_tower_recommender_model_b200 import _native as N
lib = N.load()
for name, (res, args) in N.SIGNATURES.items():
    if not args:
        continue
    vals = [0.0 if a is ctypes.c_float else (None if (a in (ctypes.c_void_p, ctypes.c_char_p) or "LP_" in a.__name__) else 0) for a in args]
    print("CALL", name, flush=True)
    rc = getattr(lib, name)(*vals)
    msg = lib.tt_last_error()
    print("DONE", json.dumps({"name": name, "rc": int(rc), "msg": msg.decode() if msg else ""}), flush=True)
''' % ROOT
    p = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, timeout=600)
    lines = p.stdout.strip().splitlines()
    assert p.returncode == 0, f"crashed in {lines[-1] if lines else '?'}: rc {p.returncode}\n{p.stderr[-400:]}"
    done = [json.loads(l[5:]) for l in lines if l.startswith("DONE ")]
    assert len(done) == sum(1 for _r, a in N.SIGNATURES.values() if a)
    for d in done:
        if d["name"].endswith("_workspace_bytes"):
            assert d["rc"] > 0, d
        elif d["name"].startswith("tt_set_") or d["name"] == "tt_adam_flat":
            assert d["rc"] == 0, d                     # mode setters; Adam over zero elements is a no-op
        else:
            assert d["rc"] < 0 and d["msg"], d         # bad arguments are reported, with text


def test_retrieval_with_an_empty_query_batch():
    """No queries -> empty [0, k] results, no kernel launch (so this runs without a GPU)."""
    items = torch.randn(50, 8)
    for index in (tt.BruteForceIndex(items), tt.CorpusShardedIndex(items, first_id=0)):
        s, i = index.search(torch.zeros(0, 8), 10)
        assert s.shape == (0, 10) and i.shape == (0, 10) and s.dtype == torch.float32 and i.dtype == torch.int64



def test_torchrec_docstring_examples(monkeypatch):
    """The examples TorchRec publishes in its own docstrings (torchrec 0.7 / 1.0, the versions the reference pins at
    requirements.txt:1; not installable here, quoted from their documentation), on this package's containers:
    KeyedJaggedTensor -- keys [Feature0, Feature1], bags [V0 V1] [] [V2] / [V3] [V4] [V5 V6 V7]  ->  lengths [2 0 1 1 1 3],
    offsets [0 2 2 3 4 5 8], offset_per_key [0 3 8];  EmbeddingBagCollection -- tables t1 (dim 3, feature f1) and t2 (dim 4,
    feature f2) on the same bags  ->  pooled values of shape [3, 7], keys [f1, f2], offset_per_key [0 3 7]."""
    vals = torch.arange(8)
    by_lengths = tt.KeyedJaggedTensor(keys=["Feature0", "Feature1"], values=vals, lengths=torch.tensor([2, 0, 1, 1, 1, 3], dtype=torch.int32))
    by_offsets = tt.KeyedJaggedTensor(keys=["Feature0", "Feature1"], values=vals, offsets=torch.tensor([0, 2, 2, 3, 4, 5, 8], dtype=torch.int32))
    for kjt in (by_lengths, by_offsets):
        assert kjt.lengths().tolist() == [2, 0, 1, 1, 1, 3] and kjt.offsets().tolist() == [0, 2, 2, 3, 4, 5, 8]
        assert kjt.offset_per_key() == [0, 3, 8] and kjt.length_per_key() == [3, 5] and kjt.stride() == 3
        assert kjt["Feature0"].values().tolist() == [0, 1, 2] and kjt["Feature1"].lengths().tolist() == [1, 1, 3]
        assert kjt.to_dict()["Feature1"].values().tolist() == [3, 4, 5, 6, 7]
        # the same container permuted / split by key (KeyedJaggedTensor.permute / .split: whole keys move, bags keep their order)
        p = kjt.permute([1, 0])
        assert p.keys() == ["Feature1", "Feature0"] and p.values().tolist() == [3, 4, 5, 6, 7, 0, 1, 2]
        assert p.lengths().tolist() == [1, 1, 3, 2, 0, 1] and p.offset_per_key() == [0, 5, 8]
        a, b = kjt.split([1, 1])
        assert a.keys() == ["Feature0"] and a.values().tolist() == [0, 1, 2] and a.lengths().tolist() == [2, 0, 1]
        assert b.keys() == ["Feature1"] and b.values().tolist() == [3, 4, 5, 6, 7] and b.lengths().tolist() == [1, 1, 3]
    # the EmbeddingBagCollection example; the one device call is replaced by the oracle (tests only)
    import oracle
    from oracle.ebc import TableSpec
    from two_tower_recommender_model_b200 import _native as N
    from two_tower_recommender_model_b200.modules import embedding_modules

    class Lookup(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ebc, kjt_keys, values, offsets, batch, *anchors):
            specs = [TableSpec(c.name, c.num_embeddings, c.embedding_dim, list(c.feature_names)) for c in ebc.embedding_bag_configs()]
            ws = [ebc.embedding_bags[s.name].weight.detach() for s in specs]
            return oracle.ebc_forward(specs, ws, list(kjt_keys), values, (offsets[1:] - offsets[:-1]).to(torch.int32))

    monkeypatch.setattr(embedding_modules, "EbcLookup", Lookup)
    monkeypatch.setattr(N, "require_cuda", lambda t, name: None)
    ebc = tt.EmbeddingBagCollection(tables=[
        tt.EmbeddingBagConfig(name="t1", embedding_dim=3, num_embeddings=10, feature_names=["f1"]),
        tt.EmbeddingBagConfig(name="t2", embedding_dim=4, num_embeddings=10, feature_names=["f2"])], device=torch.device("cpu"))
    features = tt.KeyedJaggedTensor(keys=["f1", "f2"], values=vals, offsets=torch.tensor([0, 2, 2, 3, 4, 5, 8], dtype=torch.int32))
    with torch.no_grad():
        pooled = ebc(features)
    assert tuple(pooled.values().shape) == (3, 7) and pooled.keys() == ["f1", "f2"] and pooled.offset_per_key() == [0, 3, 7]
    w1, w2 = ebc.embedding_bags["t1"].weight.detach(), ebc.embedding_bags["t2"].weight.detach()
    torch.testing.assert_close(pooled["f1"], torch.stack([w1[0] + w1[1], torch.zeros(3), w1[2]]))          # sum pooling, empty bag -> zeros
    torch.testing.assert_close(pooled["f2"], torch.stack([w2[3], w2[4], w2[5] + w2[6] + w2[7]]))


def test_a_replaced_loss_fn_is_called_like_the_reference_does(monkeypatch):
    """utils/model_training.py:129-140: the task's loss is ``self.loss_fn(logits, labels.float())``.  The plain mean BCE is
    computed by the fused kernel; a module put in its place (here ``pos_weight``) must be honoured, not silently ignored."""
    import oracle
    from oracle.ebc import TableSpec
    from two_tower_recommender_model_b200 import _native as N
    from two_tower_recommender_model_b200 import two_tower as two_tower_mod
    from two_tower_recommender_model_b200.modules import embedding_modules, mlp

    class Lookup(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ebc, kjt_keys, values, offsets, batch, *anchors):
            specs = [TableSpec(c.name, c.num_embeddings, c.embedding_dim, list(c.feature_names)) for c in ebc.embedding_bag_configs()]
            ws = [ebc.embedding_bags[s.name].weight.detach() for s in specs]
            return oracle.ebc_forward(specs, ws, list(kjt_keys), values, (offsets[1:] - offsets[:-1]).to(torch.int32))

    fused_calls = []

    def fused_bce(q, c, labels):
        fused_calls.append(1)
        logits = (q * c).sum(dim=1)
        return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels.float()), logits

    monkeypatch.setattr(embedding_modules, "EbcLookup", Lookup)
    monkeypatch.setattr(mlp, "linear_act", lambda x, w, b, relu: torch.relu(torch.nn.functional.linear(x, w, b)) if relu else torch.nn.functional.linear(x, w, b))
    monkeypatch.setattr(two_tower_mod, "dot_bce_loss", fused_bce)
    monkeypatch.setattr(N, "require_cuda", lambda t, name: None)
    torch.manual_seed(0)
    ebc = tt.EmbeddingBagCollection(tables=[tt.EmbeddingBagConfig(name=f"t_{k}", embedding_dim=8, num_embeddings=20, feature_names=[k])
                                            for k in ("u", "i")], device=torch.device("cpu"))
    task = tt.TwoTowerTrainTask(tt.TwoTower(ebc, [8, 4], device=torch.device("cpu")))
    kjt = tt.KeyedJaggedTensor.from_lengths_sync(["u", "i"], torch.arange(12) % 20, torch.ones(12, dtype=torch.int32))
    batch = tt.Batch(torch.zeros(1), kjt, torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.int32))
    with torch.no_grad():
        plain, (_, logits, labels) = task(batch)
        assert fused_calls == [1] and task._loss_fn_is_plain_bce()
        task.loss_fn = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(3.0))
        weighted, (_, logits_w, _) = task(batch)
    assert fused_calls == [1]                                   # the fused kernel was not used for the replaced loss
    torch.testing.assert_close(logits_w, logits)
    want = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels.float(), pos_weight=torch.tensor(3.0))
    torch.testing.assert_close(weighted, want)
    assert abs(float(weighted) - float(plain)) > 1e-4


def test_build_digest_does_not_depend_on_where_the_tree_lies(tmp_path):
    """The GPU box runs from a scratch copy of the tree: the library built here must count as current there (no recompile
    in smoke()), and a changed source must not."""
    import importlib.util
    import shutil
    from two_tower_recommender_model_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = tmp_path / "elsewhere"
    shutil.copytree(os.path.join(root, "two_tower_recommender_model_b200", "csrc"), dst / "two_tower_recommender_model_b200" / "csrc")
    shutil.copytree(os.path.join(root, "include"), dst / "include")
    shutil.copy(os.path.join(root, "two_tower_recommender_model_b200", "build.py"), dst / "two_tower_recommender_model_b200" / "build.py")
    spec = importlib.util.spec_from_file_location("build_elsewhere", dst / "two_tower_recommender_model_b200" / "build.py")
    other = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(other)
    assert other.CSRC != build.CSRC and other._digest() == build._digest()
    with open(dst / "two_tower_recommender_model_b200" / "csrc" / "common.cuh", "a") as f:
        f.write("\n// changed\n")
    assert other._digest() != build._digest()
