import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import sovits_b200  # noqa: E402,F401  (alias loader for the `so-vits-svc_b200/` package directory)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cfg():
    from sovits_b200.config import load_config
    return load_config()


@pytest.fixture(scope="session")
def sd(cfg):
    from sovits_b200 import synth
    return synth.synth_state_dict(cfg)
