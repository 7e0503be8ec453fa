"""The persistent kernel's ALGORITHM on the host (tests/kernel_model.cpp: test infrastructure built from
the kernel's own host-compilable arithmetic, ts_expsi.cuh + ts_fixed.cuh) against the oracle: the
exp(psi) reformulation of the two softmaxes, b = f(lambda_t)/f(lambda_0+lambda_1), fixed-point totals,
convergence test and gamma step reproduce the reference's trajectory (snpsamplinge.cc:320-366,
:695-740) without a GPU.  The CUDA code itself is compared with the oracle in test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT, load_case


@pytest.fixture(scope="module")
def km(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("km") / "libkm.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "kernel_model.cpp")], check=True)
    return C.CDLL(so)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name,iters,ipt", [("fixture", 600, 3), ("synthA", 300, 4), ("synthB", 300, 1)])
def test_kernel_algorithm_model_vs_oracle(km, name, iters, ipt):
    import terastructure_b200 as ts
    c = load_case(name)
    n, l, k, seed = c["n"], c["l"], c["k"], c["seed"]
    y = np.ascontiguousarray(c["y"], dtype=np.uint8)
    r = ts.Rng(seed)                       # the host side of the product: same draws as the reference
    vl, vo, vi = r.sample_validation(n, l, c["rows"])
    g = np.ascontiguousarray(r.init_gamma(n, k))
    ym = y.copy()                          # what k_build_vcol does: held-out genotypes become "missing"
    for i, loc in enumerate(vl):
        ym[loc, vi[vo[i]:vo[i + 1]]] = 3
    locs = r.sample_locs(l, iters).astype(np.uint32)
    cnt = np.zeros(n, np.uint32)
    lam = np.ones((l, k, 2))
    rounds = np.zeros(iters, np.uint32)
    km.km_train(n, l, k, _ptr(ym), _ptr(g), _ptr(cnt), _ptr(lam), _ptr(locs), iters, 10, C.c_double(1e-3),
                C.c_double(1.0 / k), C.c_double(1.0), C.c_double(2.0), ipt, _ptr(rounds))
    o = ol.Oracle(y, k, seed)
    ro = []
    for i in range(iters):
        loc = o.sample_loc()
        assert loc == locs[i]
        ro.append(o.train_loc(loc))
    o.flush()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()  # noqa: E731
    assert np.array_equal(rounds, np.array(ro, dtype=np.uint32))
    assert np.array_equal(cnt, o.counts)
    assert rel(g, o.gamma) < 1e-12 and rel(lam, o.lam) < 1e-12
