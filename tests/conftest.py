import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(GOLDEN / name, allow_pickle=True)
        return cache[name]
    return load


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared object (built on demand: building the library is not using it)."""
    from wav2vec_heart_sounds_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib


def reference_modules():
    """The real reference, importable only in the build container (None on the GPU box)."""
    try:
        from oracle.make_golden import import_reference, REF_SRC
        if not REF_SRC.exists():
            return None
        return import_reference()
    except BaseException:
        return None
