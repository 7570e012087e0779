import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    """tests/golden/golden.npz: outputs of the unmodified reference (make_golden.py)."""
    data = np.load(ROOT / 'tests' / 'golden' / 'golden.npz')
    return {key: data[key] for key in data.files}


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as module
    module.build()
    return module
