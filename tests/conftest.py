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


# ---- the two execution paths of the 3D cycles (csrc/mega.cuh): the persistent kernel with grid barriers ("mega": every
# level inside it), the launch path ("launch": one kernel per operation, as in round 1) and their mixture ("hybrid":
# only levels of at most 8 boxes inside the persistent kernel, so segments open and close within a cycle) and the
# cluster variant ("cluster": the bottom levels in one thread-block cluster with the hardware cluster barrier; once with
# the default level limit, once with every level inside so that CTAs loop over several blocks per phase).  The
# bit-exact suites run under all of them; the handle reads the environment at afmg_create.
PATH_MODULES = {"test_gpu_kernels", "test_gpu_cycles", "test_golden"}
PATH_ENV = {"mega": {"AFMG_MEGA": "1", "AFMG_MEGA_MAX_BOXES": "1000000"},
            "launch": {"AFMG_MEGA": "0"},
            "hybrid": {"AFMG_MEGA": "1", "AFMG_MEGA_MAX_BOXES": "8"},
            "cluster": {"AFMG_MEGA": "1", "AFMG_MEGA_CLUSTER": "16"},
            "cluster_all": {"AFMG_MEGA": "1", "AFMG_MEGA_CLUSTER": "8", "AFMG_MEGA_MAX_BOXES": "1000000"}}


@pytest.fixture
def afmg_path(request):
    old = {k: os.environ.get(k) for k in ("AFMG_MEGA", "AFMG_MEGA_MAX_BOXES", "AFMG_MEGA_CLUSTER")}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(PATH_ENV[request.param])
    yield request.param
    for k, v in old.items():
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = v


def pytest_generate_tests(metafunc):
    if metafunc.module.__name__ in PATH_MODULES and "gpu" in [m.name for m in metafunc.definition.iter_markers()]:
        metafunc.fixturenames.insert(0, "afmg_path")
        metafunc.parametrize("afmg_path", sorted(PATH_ENV), indirect=True)
