import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the upstream reference checkout at /root/reference")


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "postprocessing.py"))


def import_reference():
    """Import the unmodified reference modules (authoring container only)."""
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # preprocessing.py:61 uses the removed alias
    sys.dont_write_bytecode = True
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import KGnet, postprocessing, nms  # noqa: E401
    return KGnet, postprocessing, nms


@pytest.fixture(scope="session")
def reference():
    if not have_reference():
        pytest.skip("reference checkout not present")
    return import_reference()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
