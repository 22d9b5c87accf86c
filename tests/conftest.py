import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cfg():
    import trajtrack_mpcndqn_rlboost_b200 as t
    return t.Configurator().to_ttmpc()


@pytest.fixture(scope="session")
def golden_problem():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "problem_default.npz"))


@pytest.fixture(scope="session")
def golden_qnet():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "qnet_ray.npz"))


@pytest.fixture(scope="session")
def golden_host():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "host_logic.npz"))
