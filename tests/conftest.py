"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` (CPU, every round):  oracle vs the reference's golden vectors, host logic through
the numpy kernel double in tests/emu.py, C-ABI symbol export.  `-m gpu` (B200): the parity tests
proper, through the C-ABI library.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(path):
    from pycc_b200.synthetic import Synthetic, full_eri
    g = dict(np.load(path))
    syn = Synthetic(int(g["no"]), int(g["nv"]), g["B"], g["F"], float(g["scale"]), int(g["seed"]))
    return g, syn


@pytest.fixture(params=GOLDEN, ids=[os.path.basename(p)[4:-4] for p in GOLDEN])
def golden(request):
    return load_golden(request.param)
