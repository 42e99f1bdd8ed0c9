import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
# synthetic-weight tests: a `pretrained=` / `teacher_ckpt=` that cannot be found is an error by default
# (ld_b200/detectors.py load_checkpoint_into); tests/test_config_dropin.py checks that default explicitly
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # a stalled kernel must not hold the whole tier: pytest-timeout ends the run with the stacks of every thread
        if config.pluginmanager.hasplugin('timeout'):
            for it in items:
                if 'gpu' in it.keywords and it.get_closest_marker('timeout') is None:
                    it.add_marker(pytest.mark.timeout(600))
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
