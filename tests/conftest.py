import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

EXTRA = {'aar': 1200, 'r-o_ratio': 0.45, 'r-o_split': (0.10, 0.15, 0.15, 0.30, 0.30)}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


class Catchment(object):
    """The reference's test catchment (tests/data/in/Catchment) as processed by the reference."""

    def __init__(self):
        proc = load_golden("catchment_processed")
        self.rain = np.repeat(proc["rain_hourly_per_day"], 24)
        self.peva = np.repeat(proc["peva_hourly_per_day"], 24)
        self.obs = proc["nd_flow"]
        self.area = float(proc["area_m2"])
        self.n_steps = 87672
        self.gap = 24
        self.dt = 3600.0
        self.warm_up_days = 365
        self.extra = EXTRA


@pytest.fixture(scope="session")
def catchment():
    return Catchment()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle
