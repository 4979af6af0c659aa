import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _y2_env_switches_reloaded():
    """The library caches its Y2_* environment switches once per process; tests that flip one call ops.reload_env().  This
    runs after monkeypatch has restored the environment and brings the cache back in line with it."""
    yield
    try:
        import torch
        if torch.cuda.is_available():
            from tensorflow_yolo2_b200 import ops
            ops.reload_env()
    except Exception:
        pass
