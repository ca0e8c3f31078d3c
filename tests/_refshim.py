"""Imports the upstream reference (read-only at /root/reference) for oracle pinning.  Only available in the
build container; tests that need it are skipped elsewhere (the GPU box has no /root/reference)."""
import os
import sys
import types

REF_ROOT = "/root/reference"


def have_reference():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def cfg(num_classes=20, num_joints=16, layers=16, init_channels=64, refine_layers=1, search=False):
    ns = types.SimpleNamespace
    c = ns(DATASET=ns(NUM_CLASSES=num_classes, NUM_JOINTS=num_joints),
           TRAIN=ns(LAYERS=layers, INIT_CHANNELS=init_channels),
           SEARCH=ns(LAYERS=layers, INIT_CHANNELS=init_channels),
           MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=refine_layers))
    return c


def import_reference():
    """Returns the reference's modules namespace (models.operations, model_augment, criterion, ...)."""
    import numpy as np
    import torch
    if not hasattr(np, "int"):
        np.int = int  # utils/utils.py:202 uses the removed alias
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # our repo's `tests`/`oracle` packages do not shadow the reference's `models`/`core`/`utils`
    import importlib
    mods = types.SimpleNamespace()
    mods.operations = importlib.import_module("models.operations")
    mods.genotypes = importlib.import_module("models.genotypes")
    mods.model_augment = importlib.import_module("models.model_augment")
    mods.criterion = importlib.import_module("core.criterion")
    mods.evaluate = importlib.import_module("core.evaluate")
    return mods
