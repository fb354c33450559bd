import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def luts():
    import cases
    return cases.load_luts()


@pytest.fixture(scope="session")
def golden_patches():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "patches.npz")))


@pytest.fixture(scope="session")
def golden_wav_patches():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "wav_patches.npz")))


@pytest.fixture(scope="session")
def golden_synth():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "synthetic.npz")))


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The CPU oracle libraries (gcc, seconds).  Building the checker is not using it."""
    import subprocess
    have_port = os.path.exists(os.path.join(ROOT, "oracle", "_build", "libskred_dropin_port_v64.so"))
    if os.path.exists("/root/reference/synth.c") or not have_port:
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_oracle.py"), "--voices", "64"],
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
