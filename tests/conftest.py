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


def mask_text(t):
    """Mask what legitimately differs between two runs of the same command: seconds in progress
    lines, and the date + pid prefix of infer.log lines (log.cc: "[%b %e %T] [pid] [ERR] ")."""
    import re
    t = re.sub(r"took \d+ secs", "took # secs", t)
    t = re.sub(r"@ \d+ secs", "@ # secs", t)
    t = re.sub(r"(?m)^\[[A-Z][a-z]{2} [ \d]\d \d\d:\d\d:\d\d\] \[\d+\]", "[DATE] [PID]", t)
    return t


def load_fixture_rows():
    """The reference's bundled data set (data/test.bed, N=200 x L=10000): PLINK-packed rows [10000, 50].
    Stored as a compressed array (tools/make_golden.py) rather than as a copy of the .bed file."""
    z = np.load(os.path.join(GOLDEN, "fixture_genotypes.npz"))
    rows = z["rows"]
    assert rows.shape == (10000, 50) and rows.dtype == np.uint8
    return rows.copy()


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
