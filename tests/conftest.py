import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("dtu_b2", "nerf_b4", "train_b2")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """npz of tensors dumped from the unmodified reference (oracle/make_golden.py)."""

    def __init__(self, name):
        self.name = name
        self._z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def __contains__(self, k):
        return k in self._z.files

    def np(self, k):
        return self._z[k]

    def t(self, k, dtype=None):
        x = torch.from_numpy(self._z[k])
        return x.to(dtype) if dtype is not None and x.is_floating_point() else x

    def mlp(self, prefix="mlp_", dtype=None):
        return {k[len(prefix):]: self.t(k, dtype) for k in self._z.files if k.startswith(prefix)}


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


# golden-case geometry (must match oracle/make_golden.py CASES)
CASE_SPECS = {
    "dtu_b2": dict(recipe="dtu_eval", B=1, V=3, H=32, W=32, near=425.0, far=905.0, focal=90.0, images="noise", tilt=0.0, train=False),
    "nerf_b4": dict(recipe="nerf_eval_4x4", B=2, V=3, H=32, W=32, near=2.5, far=5.5, focal=44.0, images="smooth", tilt=0.06, train=False),
    "train_b2": dict(recipe="dtu_pretrain", B=2, V=2, H=32, W=32, near=425.0, far=905.0, focal=90.0, images="smooth", tilt=0.04, train=True),
}
