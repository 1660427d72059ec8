"""Makes ``import torchrec...`` resolve to this package, so that
/root/reference/utils/model_training.py:20-41 imports unchanged:

    import two_tower_recommender_model_b200 as tt
    tt.install_torchrec_shim()
    from torchrec.modules.embedding_modules import EmbeddingBagCollection   # -> ours

Only the names the reference imports are provided.  If a real ``torchrec`` is
importable the shim refuses unless ``force=True``.
"""
import importlib
import sys
import types

_MAP = {
    "torchrec": ".",
    "torchrec.distributed": ".distributed",
    "torchrec.distributed.model_parallel": ".distributed.model_parallel",
    "torchrec.distributed.comm": ".distributed.comm",
    "torchrec.distributed.planner": ".distributed.planner",
    "torchrec.distributed.planner.storage_reservations": ".distributed.planner.storage_reservations",
    "torchrec.distributed.train_pipeline": ".distributed.train_pipeline",
    "torchrec.inference": ".inference",
    "torchrec.inference.state_dict_transform": ".inference.state_dict_transform",
    "torchrec.modules": ".modules",
    "torchrec.modules.embedding_configs": ".modules.embedding_configs",
    "torchrec.modules.embedding_modules": ".modules.embedding_modules",
    "torchrec.modules.mlp": ".modules.mlp",
    "torchrec.optim": ".optim",
    "torchrec.optim.keyed": ".optim.keyed",
    "torchrec.optim.rowwise_adagrad": ".optim.rowwise_adagrad",
    "torchrec.sparse": ".sparse",
    "torchrec.sparse.jagged_tensor": ".sparse.jagged_tensor",
    "torchrec.datasets": ".datasets",
    "torchrec.datasets.utils": ".datasets.utils",
}


def install_torchrec_shim(force: bool = False) -> None:
    if "torchrec" in sys.modules and not getattr(sys.modules["torchrec"], "__tt_b200_shim__", False) and not force:
        raise RuntimeError("a real torchrec is already imported; pass force=True to shadow it")
    pkg = __name__.rsplit(".", 1)[0]
    for alias, rel in _MAP.items():
        target = importlib.import_module(pkg if rel == "." else pkg + rel)
        if alias == "torchrec":
            mod = types.ModuleType("torchrec")
            mod.__dict__.update({k: v for k, v in target.__dict__.items() if not k.startswith("__")})
            mod.__path__ = []  # a package
            mod.__tt_b200_shim__ = True
            sys.modules[alias] = mod
        else:
            sys.modules[alias] = target
    # attribute access (torchrec.distributed.X) follows sys.modules
    for alias in _MAP:
        if "." in alias:
            parent, child = alias.rsplit(".", 1)
            try:
                setattr(sys.modules[parent], child, sys.modules[alias])
            except Exception:
                pass
