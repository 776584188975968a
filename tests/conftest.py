import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the unmodified reference, when it has been materialised (baseline/install_reference.py; it travels to the GPU box):
# with it importable the package's samplers / integrators / loss are subclasses of the reference's own classes
# (torchebm_b200/dropin.py).  EBM_B200_STANDALONE=1 runs the same suite against the standalone classes.
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(REF_DIR, "torchebm")) and REF_DIR not in sys.path:
    sys.path.insert(1, REF_DIR)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
