import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_fixture_rows():
    """The reference's bundled data set (data/test.bed): packed rows [10000, 50]."""
    raw = open(os.path.join(GOLDEN, "fixture_n200_l10000.bed"), "rb").read()
    assert raw[:3] == b"\x6c\x1b\x01"
    return np.frombuffer(raw[3:], dtype=np.uint8).reshape(10000, 50).copy()


def load_case(name):
    """-> dict(rows=packed [L, bytes], y=[L, N], n, l, k, seed, rfreq, gold=npz)."""
    from terastructure_b200 import plink, synth
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    if name == "fixture":
        rows = load_fixture_rows()
        n, l, k, seed, rfreq = 200, 10000, 3, 1234, 1000
    else:
        n, l, k = (int(v) for v in gold["shape"])
        seed, rfreq = int(gold["seed"]), int(gold["rfreq"])
        y, _, _ = synth.psd_genotypes(n, l, k, seed=int(gold["data_seed"]),
                                      missing_rate=float(gold["missing_rate"]))
        rows = plink.pack(y)
    return dict(rows=rows, y=plink.unpack(rows, n), n=n, l=l, k=k, seed=seed, rfreq=rfreq, gold=gold)


@pytest.fixture(scope="session")
def fixture_case():
    return load_case("fixture")
