"""CPU tests: the parity oracle (oracle/ts_oracle.c) against the reference's own fixture
(data/output_theta.txt) and against outputs of the reference binary (tests/golden/*.npz,
made by tools/make_golden.py from the unmodified sources at -nthreads 1)."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_case

# md5 of the reference's shipped data/output_theta.txt
SHIPPED_THETA_MD5 = "ae1136d8769318e9840b1f7dd1d1ac53"


def fmt_rows(a):
    return "".join("".join("%.8f\t" % v for v in row) + "\n" for row in a)


def test_golden_vectors_appendix_d(fixture_case):
    """validation loci, gamma_0, SNP sequence for seed 1234."""
    c = fixture_case
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    loc, off, ind = o.validation()
    assert loc[:8].tolist() == [124, 177, 584, 835, 1179, 1437, 1775, 1915]
    assert len(loc) == 50 and np.all(np.diff(off) == 20)
    np.testing.assert_allclose(o.gamma[:3], [[0.93831476, 0.91198268, 0.97582430],
                                             [0.97908901, 0.96681601, 0.92016829],
                                             [1.18469923, 1.04423320, 0.90116638]], atol=5e-9)
    stop, a, cnt = o.heldout(first=True)
    assert cnt == 1000 and abs(a - (-1.169339294)) < 5e-10 and not stop
    assert [o.sample_loc() for _ in range(12)] == [4512, 5810, 6508, 3177, 8093, 7291, 2829, 6486,
                                                   2733, 7778, 3857, 648]


def test_oracle_reproduces_shipped_theta(fixture_case):
    """data/run.sh line 1 end to end: same reports, same stop (9050), byte-identical theta.txt."""
    c = fixture_case
    g = c["gold"]
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    o.heldout(first=True)
    r = o.infer(c["rfreq"], 10 ** 9)
    assert r["stopped"]
    assert r["iters"].tolist() == g["val_iter"][1:].tolist()
    np.testing.assert_allclose(r["ll"], g["val_ll"][1:], atol=6e-10, rtol=0)
    assert r["count"].tolist() == g["val_count"][1:].tolist()
    assert hashlib.md5(fmt_rows(o.theta).encode()).hexdigest() == SHIPPED_THETA_MD5
    np.testing.assert_allclose(o.theta, g["shipped_theta"], atol=5.1e-9, rtol=0)
    np.testing.assert_allclose(o.gamma, g["gamma"], atol=5.1e-9, rtol=0)


@pytest.mark.parametrize("name", ["synthA", "synthB"])
def test_oracle_vs_reference_binary_synthetic(name):
    """Synthetic shapes (missing data; N<2000 and N>=2000 validation branches): validation LL
    and gamma of every report equal the reference binary's to print precision."""
    c = load_case(name)
    g = c["gold"]
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    stop, a, cnt = o.heldout(first=True)
    assert abs(a - g["val_ll"][0]) < 6e-10 and cnt == g["val_count"][0]
    np.testing.assert_allclose(o.gamma, g["gamma_0"], atol=5.1e-9, rtol=0)
    for i in range(1, len(g["val_iter"])):
        r = o.infer(c["rfreq"], int(g["val_iter"][i]))
        assert r["iters"].tolist() == [g["val_iter"][i]]
        assert abs(r["ll"][0] - g["val_ll"][i]) < 6e-10
        np.testing.assert_allclose(o.gamma, g[f"gamma_{g['val_iter'][i]}"], atol=5.1e-9, rtol=1e-12)


def test_oracle_compute_beta(fixture_case):
    """data/run.sh line 2 (-compute-beta, unseeded): beta.txt to print precision.  gamma is what
    load_gamma reads back from the 8-decimal gamma.txt (lossy on purpose)."""
    c = fixture_case
    g = c["gold"]
    o = ol.Oracle(c["y"], c["k"], 0, online_iterations=100, compute_beta=True)
    o.set_gamma(g["gamma"])
    o.compute_all_lambda()
    np.testing.assert_allclose(o.beta, g["beta"], atol=5.1e-9, rtol=0)
    np.testing.assert_allclose(o.beta[:2], [[0.74197435, 0.82574638, 0.86951150],
                                            [0.76966272, 0.37108201, 0.65817304]], atol=5.1e-9)


def test_recovery_vs_simulation_truth(fixture_case):
    """Statistical recovery: the fitted theta matches the simulation truth
    under the best column permutation to RMSE ~0.05 (not a tight-parity check)."""
    truth_path = "/root/reference/data/oracle_theta.txt"
    if not os.path.exists(truth_path):
        pytest.skip("simulation truth lives in /root/reference (absent here)")
    truth = np.loadtxt(truth_path)
    theta = fixture_case["gold"]["theta"]
    import itertools
    best = min(np.sqrt(np.mean((theta - truth[:, list(p)]) ** 2)) for p in itertools.permutations(range(3)))
    assert best < 0.06
