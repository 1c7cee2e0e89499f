import os
import sys

import pytest

# fresh device scratch memory starts as garbage in the tests, so that a read of something no kernel wrote cannot pass
# by luck on zero-filled pages (audio-modem_b200/csrc/amb_api.cu, DevBuf::ensure)
os.environ.setdefault("AMB_POISON", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def amb():
    import audio_modem_b200

    audio_modem_b200.build()
    return audio_modem_b200


@pytest.fixture(scope="session")
def emu():
    import helpers

    return helpers.load_emu()
