"""Test configuration: `gpu` marker + session-scoped oracle cases.

Oracle cases are produced by running the reference itself (oracle/_ref/ref_harness, built
from /root/reference by oracle/Makefile and shipped as a prebuilt binary) on our input files.
When the prebuilt reference is unavailable the committed fixtures under tests/golden (made by
tests/golden/make_golden.py from the same harness) are used instead.
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import pyoracle as po  # noqa: E402  (tests are allowed to use the oracle)
from citcomcu_b200 import inputfile  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


CASES = {
    # name: (input text, nsteps, kat)
    "busse_l3": lambda: (inputfile.busse1a(levels=3, maxstep=2), 1, True),
    "busse_l4_tight": lambda: (inputfile.busse1a(levels=4, maxstep=1, accuracy=1e-8), 0, True),
    "tdepv_l3_tight": lambda: (inputfile.tdepv_box(16, 16, 8, 3, maxstep=1, accuracy=1e-8), 0, True),
    "tdepv_l3": lambda: (inputfile.tdepv_box(16, 16, 8, 3, maxstep=2), 1, True),
    # BASELINE config 2 (examples/input1 as Cartesian): non-uniform z spacing -> non-trivial interpolation weights / element sizes
    "input1_cart_l3": lambda: (inputfile.input1_cart(levels=3, maxstep=1), 0, True),
    # tall box: many z layers per column of the column-resident kernels (csrc/ccu_col.cuh), several ring turns
    "tdepv_tall": lambda: (inputfile.tdepv_box(8, 16, 64, 3, maxstep=1), 0, True),
}


class GoldenDump:
    """Same mapping interface as pyoracle.Dump, backed by a committed .npz fixture."""

    def __init__(self, path):
        self.z = np.load(path)

    def __contains__(self, k):
        return k in self.z.files

    def __getitem__(self, k):
        return self.z[k]

    levmin = property(lambda self: int(self["meta"][0]))
    levmax = property(lambda self: int(self["meta"][1]))
    dims = po.Dump.dims
    control = po.Dump.control


_cache = {}


def get_case(name):
    if name in _cache:
        return _cache[name]
    if po.have_ref():
        txt, nsteps, kat = CASES[name]()
        wd = Path(tempfile.mkdtemp(prefix=f"ccu_{name}_"))
        dumps, err = po.run_harness(txt, wd, nsteps=nsteps, kat=kat)
        _cache[name] = (dumps[0], err)
    else:
        f = GOLDEN / f"{name}.npz"
        if not f.exists():
            pytest.skip(f"no prebuilt reference and no golden fixture for {name}")
        _cache[name] = (GoldenDump(f), "")
    return _cache[name]


@pytest.fixture(scope="session")
def oracle_built():
    if not po.have_restate():
        po.build()
    return True


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
