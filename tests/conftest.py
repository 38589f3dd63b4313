import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    """One committed fixture produced by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(str(self.z["meta"]))
        self.shape = self.meta["shape"]
        self.kwargs = self.meta["kwargs"]
        self.gbar = self.meta["gbar"]

    @property
    def inputs(self):
        return tuple(self.z["in." + k] for k in ("masks", "fw", "bw", "rfw", "rbw"))

    @property
    def params(self):
        return {k[len("param."):]: v for k, v in self.z.items() if k.startswith("param.")}

    def head_kwargs(self):
        s = self.shape
        return dict(mask_layer=s["K"], mask_size=(s["H"], s["W"]), **self.kwargs)

    def ref(self, tag, key):
        return self.z[f"{tag}.{key}"]

    def has(self, tag, key):
        return f"{tag}.{key}" in self.z


@pytest.fixture(params=golden_names())
def golden(request):
    return Golden(request.param)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))
