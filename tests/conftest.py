import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def bunny():
    from icp_variants_b200 import synth
    return synth.load_bunny()


@pytest.fixture(scope="session")
def small_eth_pair():
    """A reduced ETH-shaped pair (~9.6k points per scan) the oracle handles in well under a second."""
    from icp_variants_b200 import synth
    return synth.eth_pair(seed=1234, n_sweeps=60, n_beams=160)
