import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the suites load libafmg.so (the product) and the oracle library (the checker); build them if a fresh checkout
    # has not been through `python __graft_entry__.py` yet (both are git-ignored build artefacts)
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "afivo_streamer_b200", "libafmg.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "afivo_streamer_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "libafmg_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
