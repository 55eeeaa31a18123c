"""Stage the UNMODIFIED reference's hot-path modules under oracle/_ref/ so they travel to the GPU box.
TEST / BASELINE INFRASTRUCTURE ONLY - never imported by vince_b200/.

    python oracle/build_ref.py        (also called by __graft_entry__.build() when /root/reference is mounted)

The reference is pure Python: "building" it means copying the 13 files its hot path imports
(constants.py, models/{__init__,base_model,vince_model}.py, models/building_blocks/{__init__,backbone_models,resnet}.py,
utils/{__init__,loss_util,storage_queue,util_functions}.py, solvers/{vince_solver,base_solver}.py, plus the four class-name lists util_functions reads at
import time; ~140 KB) byte for byte, with their directory layout, into
oracle/_ref/reference/.  oracle/_ref/ is git-ignored (reference SOURCES never enter the history) but not
gpurun-ignored, so `bench.py --impl reference` on the GPU box times the reference's own VinceModel / VinceQueueModel /
StorageQueue / loss_util (cpu_baseline.kind = "reference"), imported through oracle/ref_loader.py + the dg_util shim.
A manifest with the sha256 of every copied file is written next to them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VINCE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
FILES = [
    "constants.py",
    "models/__init__.py", "models/base_model.py", "models/vince_model.py",
    "models/building_blocks/__init__.py", "models/building_blocks/backbone_models.py", "models/building_blocks/resnet.py",
    "utils/__init__.py", "utils/loss_util.py", "utils/storage_queue.py", "utils/util_functions.py",
    # the training loop itself (tests/test_gpu_parity.py::test_reference_solver_runs_unchanged_on_vince_b200_classes runs
    # its run_train_iteration, unmodified, against the vince_b200 classes); loaded by file path, so the package
    # __init__ files of solvers/ and datasets/ (which import every dataset / end-task solver) are not needed
    "solvers/vince_solver.py", "solvers/base_solver.py",
    # class-name lists utils/util_functions.py:12-33 reads at import time (labels for visualisations; 80 KB)
    "datasets/info_files/imagenet_class_names.json", "datasets/info_files/sun_scene_class_names.txt",
    "datasets/info_files/kinetics_400_class_names.txt", "datasets/info_files/yt8m_class_names.txt",
]


def build(verbose=True):
    """Returns the staged root, or None when the reference is not mounted here (the GPU box: use what travelled)."""
    if not os.path.isfile(os.path.join(SRC, "models", "vince_model.py")):
        return DST if os.path.isfile(os.path.join(DST, "models", "vince_model.py")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    if verbose:
        print("oracle/_ref: staged %d reference files under %s" % (len(FILES), DST))
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
