import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_prims():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'prims_cv2.npz'))


@pytest.fixture(scope='session')
def golden_orb():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'orb_ref.npz'))


@pytest.fixture(scope='session')
def hvo():
    import hvo_b200
    return hvo_b200


@pytest.fixture(scope='session')
def synth():
    from hvo_b200 import synth as s
    return s
