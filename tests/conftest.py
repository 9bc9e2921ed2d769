import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def plasmids():
    from barcoder_b200.seqio import genbank_to_dict
    return genbank_to_dict(os.path.join(GOLDEN, "zmo_plasmids.gb"))


@pytest.fixture(scope="session")
def cn32_spacers():
    with open(os.path.join(GOLDEN, "cn32_spacers.txt")) as h:
        return [l.strip() for l in h if l.strip()]
