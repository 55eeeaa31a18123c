"""Import the UNMODIFIED reference (danielgordon10/vince) on CPU.  TEST INFRASTRUCTURE ONLY.

In the dev container the reference is mounted read-only at /root/reference; elsewhere (the GPU
box) the byte copies staged by oracle/build_ref.py under oracle/_ref/reference are used (git-ignored,
travels with the gpurun snapshot).  It is used by
  * oracle/make_golden.py   - to generate tests/golden/*.npz from the real reference
  * tests/test_oracle_vs_reference.py - to pin oracle/vince_oracle.py to the reference
    (skipped automatically when neither location exists)
  * bench.py --impl reference / cpu_baseline - the reference's own CPU implementation timed on the
    host cores (cpu_baseline.kind = "reference")
Never imported by vince_b200/ or by the `-m gpu` parity tests.

The reference needs two un-installable packages (`dg_util`, `efficientnet_pytorch`);
oracle/dg_util_shim provides import-level stand-ins.
"""
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref", "reference")        # oracle/build_ref.py: byte copies of the hot-path modules


def _default_root():
    if os.path.isfile("/root/reference/models/vince_model.py"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = os.environ.get("VINCE_REFERENCE_ROOT", _default_root())
_SHIM = os.path.join(_HERE, "dg_util_shim")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "vince_model.py"))


_cached = None


def load_reference():
    """Returns a namespace with the reference's hot-path classes/modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    import numpy as np

    if not hasattr(np, "bool"):
        np.bool = bool  # vince_model.py:54 uses the removed alias
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import constants as ref_constants  # noqa: F401  (reference top-level module)
        from models import vince_model as ref_vince_model
        from models.building_blocks import backbone_models as ref_backbones
        from utils import loss_util as ref_loss_util
        from utils import storage_queue as ref_storage_queue
    ns = types.SimpleNamespace(
        VinceModel=ref_vince_model.VinceModel,
        VinceQueueModel=ref_vince_model.VinceQueueModel,
        StorageQueue=ref_storage_queue.StorageQueue,
        loss_util=ref_loss_util,
        backbones=ref_backbones,
        vince_model=ref_vince_model,
    )
    _cached = ns
    return ns


_solver = None


def solver_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "solvers", "vince_solver.py"))


def load_reference_solver():
    """The reference's VinceSolver class (solvers/vince_solver.py, unmodified).  The module is loaded by file path with
    stub `solvers` / `datasets` packages in sys.modules, so the package __init__ files (which import every dataset and
    end-task solver, cv2, ...) are never executed; the only name taken from `datasets` is NPZDataset, which
    run_train_iteration does not touch."""
    global _solver
    if _solver is not None:
        return _solver
    import importlib.util
    load_reference()
    for name in ("solvers", "datasets"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, name)]
            sys.modules[name] = m
    if "datasets.npz_dataset" not in sys.modules:
        nd = types.ModuleType("datasets.npz_dataset")
        nd.NPZDataset = object
        sys.modules["datasets.npz_dataset"] = nd
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in ("base_solver", "vince_solver"):
            full = "solvers." + name
            spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, "solvers", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    _solver = mods["vince_solver"].VinceSolver
    return _solver


def make_args(backbone="ResNet18", num_frames=4, batch_size=8, queue_size=1024, embedding_size=128,
              temperature=0.07, self_temperature=0.03, momentum=0.999, inter_batch_comparison=True,
              self_batch_comparison=False, jigsaw=False):
    """argparse.Namespace-alike with the fields the hot path reads."""
    ref = load_reference()
    return types.SimpleNamespace(
        backbone=getattr(ref.backbones, backbone),
        num_frames=num_frames,
        use_attention=False,
        feature_extractor_gpu_ids=["cpu"],
        pytorch_gpu_ids=["cpu"],
        vince_embedding_size=embedding_size,
        vince_queue_size=queue_size,
        vince_temperature=temperature,
        vince_self_temperature=self_temperature,
        vince_momentum=momentum,
        jigsaw=jigsaw,
        inter_batch_comparison=inter_batch_comparison,
        self_batch_comparison=self_batch_comparison,
        batch_size=batch_size,
        use_imagenet=False,
        use_imagenet_weights=False,
        restore=False,
        save=False,
        checkpoint_dir="/tmp/_vince_ref_ckpt",
        long_save_checkpoint_dir="/tmp/_vince_ref_ckpt_long",
        long_save_frequency=10,
        saved_variable_prefix=None,
        new_variable_prefix=None,
    )
