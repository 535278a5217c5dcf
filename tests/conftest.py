import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sci-algorithms_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The shared library is a build artefact (git-ignored): build it on a fresh checkout so that the
# CPU suite (symbol/ABI checks) can load it.  nvcc cross-compiles without a GPU.
if not os.path.isfile(os.path.join(PKG, "scipnp", "libscipnp.so")):
    import subprocess
    subprocess.run(["bash", os.path.join(PKG, "build.sh")], check=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip silently;
    # plain runs here (no GPU) skip the gpu tests unless they were asked for.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    asked = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if asked:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
            return {k: d[k] for k in d.files}
    return _load
