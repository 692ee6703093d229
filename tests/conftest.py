import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libpu_ref.so (the compiled unmodified reference)")


def pytest_collection_modifyitems(config, items):
    import refapi
    have_ref = refapi.available()
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="oracle/_ref/libpu_ref.so not built"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return {n: np.load(os.path.join(d, n + "_golden.npz")) for n in ("ldpc", "ofdm", "misc", "psk", "acquire", "dpsk_acquire", "chirp", "mcdpsk_chirp", "frame", "training_cfo")}
