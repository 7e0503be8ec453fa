"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (include/tsgpu.h),
against the oracle on the same seeded inputs, against the committed golden vectors of the
reference binary, and -- at sizes the oracle cannot reach -- through size-independent
invariants.  Tolerances (BASELINE.json north_star): SNP sequence bit-exact, theta/beta within
1e-6 relative, held-out log-likelihood within 1e-5."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import load_case

pytestmark = pytest.mark.gpu

THETA_RTOL = 1e-6
LL_ATOL = 1e-5
# what we actually expect: fp64 kernels differ from the sequential reference only by rounding
TIGHT = 1e-9


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def make_driver(c, **kw):
    import terastructure_b200 as ts
    env = ts.Env(c["n"], c["k"], c["l"], seed=c["seed"], rfreq=c["rfreq"])
    return ts.SNPSamplingE(env, c["rows"], **kw)


def test_device_digamma_matches_scipy():
    """ts_get_elogtheta exposes the device digamma: psi(g) - psi(rowsum) vs SciPy."""
    import terastructure_b200 as ts
    from scipy.special import digamma
    rs = np.random.RandomState(0)
    n, k = 4096, 5
    g = np.exp(rs.uniform(np.log(0.02), np.log(5e6), size=(n, k)))
    e = ts.Engine(n, 8, k)
    e.load_bed(np.zeros((8, n // 4), np.uint8))
    e.set_gamma(g)
    ref = digamma(g) - digamma(g.sum(1, keepdims=True))
    got = e.elogtheta
    assert np.max(np.abs(got - ref)) < 2e-14 * np.max(np.abs(ref))
    np.testing.assert_allclose(e.theta, g / g.sum(1, keepdims=True), rtol=1e-15)
    np.testing.assert_array_equal(e.gamma, g)


def test_first_snp_step(fixture_case):
    """Minimum slice: SNP 4512, the first one sampled under seed 1234."""
    c = fixture_case
    s = make_driver(c)
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    np.testing.assert_array_equal(s.val_loc, o.validation()[0])
    np.testing.assert_array_equal(s.val_indiv, o.validation()[2])
    np.testing.assert_array_equal(s.engine.gamma, o.gamma)
    _, a0, cnt = o.heldout(first=True)
    assert s.validation_rows[0][3] == cnt and abs(s.validation_rows[0][2] - a0) < 1e-12
    loc = o.sample_loc()
    assert loc == 4512
    r_gpu = s.engine.step(loc)
    r_cpu = o.train_loc(loc)
    o.flush()
    assert r_gpu == r_cpu
    # 1e-12: the E-step's reciprocals take one Newton step (relative error < 1e-11 in a weight of the sums)
    assert rel_err(s.engine.get_lambda(loc, 1)[0], o.lam[loc]) < 1e-12
    assert rel_err(s.engine.gamma, o.gamma) < 1e-12
    assert rel_err(s.engine.elogtheta, o.elogtheta) < 1e-11
    np.testing.assert_array_equal(s.engine.counts, o.counts)


def test_fixture_trajectory_to_stop(fixture_case):
    """data/run.sh line 1 on the GPU: same reports and stop iteration as the reference binary,
    LL within 1e-5, theta within 1e-6 relative of the reference's shipped output_theta.txt."""
    c = fixture_case
    g = c["gold"]
    s = make_driver(c)
    s.infer()
    assert s.stopped and s._iter == 9050
    it = [r[0] for r in s.validation_rows]
    ll = np.array([r[2] for r in s.validation_rows])
    assert it == g["val_iter"].tolist()
    assert [r[3] for r in s.validation_rows] == g["val_count"].tolist()
    assert np.max(np.abs(ll - g["val_ll"])) < LL_ATOL
    assert np.max(np.abs(ll - g["val_ll"])) < 6e-10          # print precision of validation.txt
    theta = s.engine.theta
    # golden file holds 8 decimals: allow its half-ulp on top of the relative tolerance
    assert np.all(np.abs(theta - g["shipped_theta"]) <= THETA_RTOL * np.abs(g["shipped_theta"]) + 5.1e-9)
    assert np.max(np.abs(theta - g["shipped_theta"])) < 5.1e-9
    assert np.max(np.abs(s.engine.gamma - g["gamma"])) < 5.1e-9 + 1e-9 * np.max(g["gamma"])


@pytest.mark.parametrize("name", ["synthA", "synthB"])
def test_synthetic_trajectory_vs_oracle_and_reference(name):
    """Missing data, N<2000 and N>=2000 validation branches: every report of the reference
    binary (golden) and the oracle's full-precision state at the last one."""
    c = load_case(name)
    g = c["gold"]
    s = make_driver(c)
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    o.heldout(first=True)
    last = int(g["val_iter"][-1])
    s.infer(max_iter=last)
    o.infer(c["rfreq"], last)
    it = [r[0] for r in s.validation_rows]
    ll = np.array([r[2] for r in s.validation_rows])
    assert it == g["val_iter"].tolist()
    assert np.max(np.abs(ll - g["val_ll"])) < 6e-10
    assert rel_err(s.engine.gamma, o.gamma) < TIGHT
    assert rel_err(s.engine.theta, o.theta) < TIGHT
    assert rel_err(s.engine.get_lambda(), o.lam) < TIGHT
    np.testing.assert_array_equal(s.engine.counts, o.counts)
    assert np.max(np.abs(s.engine.gamma - g[f"gamma_{last}"])) < 5.1e-9 + 1e-9 * np.max(o.gamma)


def test_training_step_on_validation_locus(fixture_case):
    """A training visit to a validation locus must skip the held-out individuals (kv_ok)."""
    c = fixture_case
    s = make_driver(c)
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    locs = [int(s.val_loc[0]), 17, int(s.val_loc[3]), int(s.val_loc[0])]
    rounds = s.engine.steps(np.array(locs, np.uint32), want_rounds=True)
    ro = [o.train_loc(l) for l in locs]
    o.flush()
    assert rounds.tolist() == ro
    assert rel_err(s.engine.gamma, o.gamma) < 1e-11
    np.testing.assert_array_equal(s.engine.counts, o.counts)
    held = s.val_indiv[s.val_off[0]:s.val_off[1]]
    assert np.all(s.engine.counts[held] <= 2)  # never stepped on the two visits to val_loc[0]


def test_heldout_per_locus(fixture_case):
    c = fixture_case
    s = make_driver(c)
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    locs = [o.sample_loc() for _ in range(25)]
    s.engine.steps(np.array(locs, np.uint32))
    for l in locs:
        o.train_loc(l)
    ssum, cnt, per = s.engine.heldout_ll(False)
    _, a, cnt_o, per_o = o.heldout(first=False, per_locus=True)
    assert cnt == cnt_o == 1000
    np.testing.assert_allclose(per, per_o, rtol=1e-11)
    assert abs(ssum / cnt - a) < 1e-12
    # hol-mode passes leave gamma alone but do overwrite lambda of the validation loci
    assert rel_err(s.engine.gamma, o.gamma) < 1e-11
    assert rel_err(s.engine.get_lambda(), o.lam) < 1e-11


def test_compute_beta_pass(fixture_case):
    """data/run.sh line 2 (-compute-beta): beta.txt of the reference binary (golden)."""
    import terastructure_b200 as ts
    c = fixture_case
    g = c["gold"]
    env = ts.Env(c["n"], c["k"], c["l"], seed=0, compute_beta=True)
    s = ts.SNPSamplingE(env, c["rows"], gamma0=g["gamma"])
    beta = s.compute_all_lambda()
    assert np.all(np.abs(beta - g["beta"]) <= 1e-6 * np.abs(g["beta"]) + 5.1e-9)
    assert np.max(np.abs(beta - g["beta"])) < 5.1e-9


@pytest.mark.parametrize("k", [1, 2, 7, 13, 20, 32])
def test_k_sweep_vs_oracle(k):
    """K sweep (BASELINE configs[4]) at a size the oracle does in seconds."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink, synth
    n, l = 1501, 400   # ragged: N not a multiple of 4
    y, _, _ = synth.psd_genotypes(n, l, max(k, 2), seed=3, missing_rate=0.03)
    rows = plink.pack(y)
    env = ts.Env(n, k, l, seed=21, rfreq=10 ** 6)
    s = ts.SNPSamplingE(env, rows)
    o = ol.Oracle(y, k, 21)
    locs = [o.sample_loc() for _ in range(40)]
    r = s.engine.steps(np.array(locs, np.uint32), want_rounds=True)
    ro = [o.train_loc(x) for x in locs]
    o.flush()
    assert r.tolist() == ro
    assert rel_err(s.engine.gamma, o.gamma) < TIGHT
    assert rel_err(s.engine.get_lambda(), o.lam) < TIGHT


def test_all_missing_column_and_tiny_n():
    """Edge cases: a locus where every genotype is missing (lambda falls back to eta, no gamma
    step), and N smaller than one warp."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink
    n, l, k = 13, 300, 3
    rs = np.random.RandomState(5)
    y = rs.randint(0, 3, size=(l, n)).astype(np.uint8)
    y[7, :] = 3
    rows = plink.pack(y)
    env = ts.Env(n, k, l, seed=4, rfreq=10 ** 6)
    s = ts.SNPSamplingE(env, rows)
    o = ol.Oracle(y, k, 4)
    locs = [5, 7, 7, 9, 7]
    r = s.engine.steps(np.array(locs, np.uint32), want_rounds=True)
    ro = [o.train_loc(x) for x in locs]
    o.flush()
    assert r.tolist() == ro
    np.testing.assert_array_equal(s.engine.get_lambda(7, 1)[0], np.ones((k, 2)))
    assert rel_err(s.engine.gamma, o.gamma) < TIGHT


def test_invariants_at_scale():
    """N=100K x L=64, K=10 generated on the device -- far beyond what the sequential oracle
    does in seconds.  Size-independent checks:
      sum_k lambda[loc][k][0] - K*eta0 = sum_n y[n]       over non-missing n (phi sums to one)
      sum_k lambda[loc][k][1] - K*eta1 = sum_n (2 - y[n])
      rowsum(gamma) follows  s <- (1-rho) s + rho (K*alpha + 2L)  per visit,  rho = (2+c)^-0.5
      counts[n] = number of visits with a non-missing genotype."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink, synth
    n, l, k = 100_000, 64, 10
    theta, beta = synth.psd_params(n, l, k, seed=1)
    e = ts.Engine(n, l, k)
    e.synth_bed(99, theta, beta, missing_rate=0.01)
    rs = np.random.RandomState(2)
    g0 = rs.gamma(100.0, 0.01, size=(n, k))
    e.set_gamma(g0)
    locs = rs.randint(0, l, size=30).astype(np.uint32)
    rounds = e.steps(locs, want_rounds=True)
    assert np.all(rounds == 10)  # at N>=10K every visit runs all 10 rounds
    rowsum = g0.sum(1)
    cnt = np.zeros(n, np.int64)
    last_visit = {}
    for i, loc in enumerate(locs):
        y = plink.unpack(e.get_bed_row(int(loc))[None, :], n)[0]
        ok = y != 3
        rho = (2.0 + cnt) ** -0.5
        rowsum = np.where(ok, (1 - rho) * rowsum + rho * (k * (1.0 / k) + 2.0 * l), rowsum)
        cnt += ok
        last_visit[int(loc)] = y
    np.testing.assert_array_equal(e.counts, cnt)
    np.testing.assert_allclose(e.gamma.sum(1), rowsum, rtol=1e-12)
    for loc, y in last_visit.items():
        lam = e.get_lambda(loc, 1)[0]
        ok = y != 3
        assert abs(lam[:, 0].sum() - k - y[ok].sum()) < 1e-7 * n
        assert abs(lam[:, 1].sum() - k - (2 - y[ok].astype(np.int64)).sum()) < 1e-7 * n
    # generator sanity: allele frequency tracks theta.beta
    y0 = plink.unpack(e.get_bed_row(0)[None, :], n)[0]
    q = theta @ beta[0]
    assert abs(y0[y0 != 3].mean() / 2 - q.mean()) < 0.01
    assert abs((y0 == 3).mean() - 0.01) < 0.003


def test_two_shards_match_one(fixture_case):
    """Individuals sharded over two engines (here: both on cuda:0, exchange through the same
    peer-store protocol used across NVLink) give the single-engine result."""
    import terastructure_b200 as ts
    from terastructure_b200 import capi
    c = load_case("synthA")
    n, l, k = c["n"], c["l"], c["k"]
    r = ts.Rng(c["seed"])
    vl, vo, vi = r.sample_validation(n, l, c["rows"])
    g0 = r.init_gamma(n, k)
    locs = r.sample_locs(l, 50)
    locs[10] = vl[0]
    one = ts.Engine(n, l, k)
    one.load_bed(c["rows"]); one.set_validation(vl, vo, vi); one.set_gamma(g0)
    half = 300
    sh = [ts.Engine(n, l, k, rank=i, nranks=2, n_begin=i * half, n_local=half) for i in range(2)]
    for i, e in enumerate(sh):
        e.load_bed(c["rows"]); e.set_validation(vl, vo, vi); e.set_gamma(g0[i * half:(i + 1) * half])
    capi.connect_local(sh)
    one.steps(locs)
    # interleave the two shards SNP by SNP so neither stream waits long for its peer
    for x in locs:
        for e in sh:
            e.steps(np.array([x], np.uint32))
    s1, c1, p1 = one.heldout_ll(False)
    per = np.zeros(len(vl)); cnt = 0
    import threading
    res = [None, None]
    th = [threading.Thread(target=lambda i=i: res.__setitem__(i, sh[i].heldout_ll(False))) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for sres in res:
        per += sres[2]; cnt += sres[1]
    assert cnt == c1
    np.testing.assert_allclose(per, p1, rtol=1e-10)
    gs = np.concatenate([e.gamma for e in sh])
    assert rel_err(gs, one.gamma) < 1e-11
    np.testing.assert_array_equal(sh[0].get_lambda(), sh[1].get_lambda())  # bit-identical on all ranks
    assert rel_err(sh[0].get_lambda(), one.get_lambda()) < 1e-11


def test_cli_drop_in(tmp_path):
    """The drop-in CLI (terastructure_b200/bin/terastructure) run exactly as data/run.sh runs the
    reference: same output directory and files; theta.txt/validation.txt/beta.txt match the
    reference binary's (golden) within the north-star tolerances."""
    import os
    import shutil
    import subprocess
    from conftest import GOLDEN, ROOT
    from terastructure_b200 import plink
    exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
    assert os.path.exists(exe), "CLI not built"
    g = np.load(os.path.join(GOLDEN, "fixture.npz"))
    from conftest import load_fixture_rows
    rows = load_fixture_rows()
    plink.write_bed(str(tmp_path / "test"), rows, 200)
    r = subprocess.run([exe, "-file", "test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic", "-nthreads", "1",
                        "-rfreq", "1000", "-seed", "1234", "-label", "test"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = tmp_path / "n200-k3-l10000-test-seed1234"
    for f in ("param.txt", "infer.log", "validation.txt", "gamma.txt", "theta.txt", "network.dat"):
        assert (d / f).exists() or (d / f).is_symlink(), f
    val = np.loadtxt(d / "validation.txt")
    assert val[:, 0].astype(int).tolist() == g["val_iter"].tolist()
    assert np.max(np.abs(val[:, 2] - g["val_ll"])) < LL_ATOL
    assert val[:, 3].astype(int).tolist() == g["val_count"].tolist()
    theta = np.loadtxt(d / "theta.txt")
    assert np.all(np.abs(theta - g["shipped_theta"]) <= THETA_RTOL * np.abs(g["shipped_theta"]) + 1.01e-8)
    # line 2 of data/run.sh: -compute-beta from inside the output directory
    r = subprocess.run([exe, "-file", "../test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic", "-nthreads", "1",
                        "-compute-beta"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    bdir = [p for p in d.iterdir() if p.is_dir()][0]
    beta = np.loadtxt(bdir / "beta.txt")
    assert beta[:, 0].astype(int).tolist() == list(range(10000))
    assert np.all(np.abs(beta[:, 1:] - g["beta"]) <= 1e-6 * np.abs(g["beta"]) + 1.01e-8)
    assert (bdir / "gammasave.txt").exists()
    # the reference refuses an existing output directory unless -force (log.cc:97-118)
    r = subprocess.run([exe, "-file", "test.bed", "-n", "200", "-l", "10000", "-k", "3", "-seed", "1234", "-label", "test"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "already exists" in r.stderr


def _staged(*args):
    """Run tests/staged_helper.py (the staged path lives in the test-only library) and load its dump."""
    import os
    import subprocess
    import sys
    import tempfile
    from conftest import ROOT
    out = os.path.join(tempfile.mkdtemp(prefix="tsstaged."), "out.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "staged_helper.py"), *[str(a) for a in args], out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def test_staged_path_matches_persistent():
    """The staged path (one launch per round, test-only library) is an independent implementation of the
    same mathematics: the persistent kernel must agree with it to rounding on a trajectory with reports."""
    c = load_case("synthA")
    last = int(c["gold"]["val_iter"][2])
    s1 = make_driver(c)
    s1.infer(max_iter=last)
    s2 = _staged("trajectory", "synthA", last)
    assert [r[0] for r in s1.validation_rows] == s2["val_iter"].tolist()
    assert np.max(np.abs(np.array([r[2] for r in s1.validation_rows]) - s2["val_ll"])) < 1e-12
    assert rel_err(s1.engine.gamma, s2["gamma"]) < 1e-11
    assert rel_err(s1.engine.get_lambda(), s2["lam"]) < 1e-11
    assert int(s2["launches"]) > 5 * s1.engine.launch_count


def test_product_library_has_no_staged_path(monkeypatch):
    """TSGPU_PATH=staged is refused by the product library (the staged kernels are test infrastructure)."""
    import terastructure_b200 as ts
    monkeypatch.setenv("TSGPU_PATH", "staged")
    with pytest.raises(ts.TsError, match="test-only library"):
        ts.Engine(64, 8, 2)


def test_exp_digamma_table_on_device():
    """The persistent kernel's exp(digamma(x)) (ts_expsi.cuh: series for x >= 8, recurrence + exp
    below) on the device: E is not exposed, so check its effect -- SVI steps on a problem whose
    gamma spans 0.04 .. 3e6 must match the oracle (which uses log/exp/series).  The host model of
    the same source is checked against mpmath in test_host.py."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink
    n, l, k = 2048, 201, 4
    rs = np.random.RandomState(3)
    y = rs.randint(0, 3, size=(l, n)).astype(np.uint8)
    g0 = np.exp(rs.uniform(np.log(0.04), np.log(3e6), size=(n, k)))
    e = ts.Engine(n, l, k)
    e.load_bed(plink.pack(y))
    e.set_gamma(g0)
    o = ol.Oracle(y, k, 1)
    o.set_gamma(g0)
    for loc in (3, 77, 3, 150):
        assert e.step(loc) == o.train_loc(loc)
    o.flush()
    assert rel_err(e.gamma, o.gamma) < 1e-12
    assert rel_err(e.get_lambda(), o.lam) < 1e-12


@pytest.mark.parametrize("eta", [0.25, 1.0])
def test_control_path_table_and_fallback_on_device(eta, tmp_path):
    """The control warp's b = f(lambda_t) / f(lambda_0 + lambda_1) on the device: table-driven (ts_ftab.cuh) with the
    reference's prior eta = 1, and through the analytic fallback with eta = 0.25 (lambda below 1 lies outside the
    table's domain) -- against the host model of the same sources (tests/kernel_model.cpp), which test_kernel_model.py
    ties to the oracle.  Same rounds, gamma and lambda to rounding (device FMA contraction and MUFU seed differ)."""
    import ctypes as C
    import os
    import subprocess
    import terastructure_b200 as ts
    from conftest import ROOT
    from terastructure_b200 import plink
    so = str(tmp_path / "libkm.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "kernel_model.cpp")], check=True)
    km = C.CDLL(so)
    n, l, k = 3001, 120, 5
    rs = np.random.RandomState(11)
    y = rs.randint(0, 3, size=(l, n)).astype(np.uint8)
    y[rs.rand(l, n) < 0.02] = 3
    g0 = rs.gamma(100.0, 0.01, size=(n, k))
    g0[:, k - 1] *= 1e-3   # a population nobody belongs to yet: its statistics stay near zero, lambda near eta, in every round
    locs = np.array([5, 9, 5, 33, 9, 100, 5, 61, 61, 7], np.uint32)
    e = ts.Engine(n, l, k, eta=eta)
    e.load_bed(plink.pack(y))
    e.set_gamma(g0)
    rounds = e.steps(locs, want_rounds=True)
    g, cnt, lam, ro = np.ascontiguousarray(g0.copy()), np.zeros(n, np.uint32), np.full((l, k, 2), eta), np.zeros(len(locs), np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    km.km_train(n, l, k, p(y), p(g), p(cnt), p(lam), p(locs), len(locs), 10, C.c_double(1e-3), C.c_double(1.0 / k),
                C.c_double(eta), C.c_double(2.0), 3, p(ro))
    assert rounds.tolist() == ro.tolist()
    np.testing.assert_array_equal(e.counts, cnt)
    assert rel_err(e.gamma, g) < 1e-11
    assert rel_err(e.get_lambda(), lam) < 1e-11
    assert (lam[locs].min() < 1.0) == (eta < 1.0)   # with eta = 0.25 the rows of the visited loci leave the table's domain


def test_cli_012_input_sigterm_and_gpus(tmp_path):
    """CLI on a .012 text file (snp.cc:6-93) with missing genotypes, stopped by SIGTERM after a few
    reports like the reference (main.cc:28-39: save the model, exit 0); `-gpus 2` when two GPUs are
    visible.  validation.txt rows and gamma_<iter>.txt (-file-suffix) against the reference binary."""
    import os
    import signal
    import subprocess
    import time
    import terastructure_b200 as ts
    from conftest import ROOT
    c = load_case("synthA")
    g = c["gold"]
    y = c["y"]  # [l, n]
    with open(tmp_path / "d.012", "w") as f:
        for row in y:
            f.write("".join("-" if v == 3 else str(int(v)) for v in row) + "\n")
    exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
    ng = 2 if ts.lib().ts_device_count() >= 2 else 1
    cmd = [exe, "-file", "d.012", "-n", str(c["n"]), "-l", str(c["l"]), "-k", str(c["k"]), "-rfreq", str(c["rfreq"]),
           "-seed", str(c["seed"]), "-label", "g", "-file-suffix", "-gpus", str(ng)]
    p = subprocess.Popen(cmd, cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    d = tmp_path / f"n{c['n']}-k{c['k']}-l{c['l']}-g-seed{c['seed']}"
    vf = d / "validation.txt"
    t0 = time.time()
    nrep = len(g["val_iter"])
    while p.poll() is None and time.time() - t0 < 240:
        time.sleep(0.2)
        if vf.exists() and sum(1 for _ in open(vf)) >= nrep:
            p.send_signal(signal.SIGTERM)
            break
    rc = p.wait(timeout=60)
    assert rc == 0, p.stderr.read()[-2000:]
    val = np.loadtxt(vf)[:nrep]
    assert val[:, 0].astype(int).tolist() == g["val_iter"].tolist()
    assert np.max(np.abs(val[:, 2] - g["val_ll"])) < 6e-10
    for it in g["val_iter"]:
        gam = np.loadtxt(d / f"gamma_{it}.txt")
        ref = g[f"gamma_{it}"]
        assert np.all(np.abs(gam - ref) <= 1e-6 * np.abs(ref) + 1.01e-8), it


def test_large_shard_tiered_kernel():
    """Shards beyond the register-resident capacity (148 CTAs x 256 threads x 4 individuals at
    K <= 12) run the tiered persistent kernel (registers + shared memory, here all on chip):
    same invariants, and agreement with the staged path on the same inputs."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink, synth
    n, l, k = 160_000, 16, 4
    theta, beta = synth.psd_params(n, l, k, seed=2)
    g0 = np.random.RandomState(1).gamma(100.0, 0.01, size=(n, k))
    locs = np.array([3, 9, 3, 5], np.uint32)
    e = ts.Engine(n, l, k)
    e.synth_bed(7, theta, beta, 0.01)
    e.set_gamma(g0)
    before = e.launch_count
    rounds = e.steps(locs, want_rounds=True)
    assert np.all(rounds == 10) and e.launch_count - before == 1    # one persistent launch for the batch
    assert e.tiers[0] >= 1 and e.tiers[1] == 0                       # shared-memory tier in use, nothing streams
    for loc in (3, 9, 5):
        y = plink.unpack(e.get_bed_row(loc)[None, :], n)[0]
        ok = y != 3
        lam = e.get_lambda(loc, 1)[0]
        assert abs(lam[:, 0].sum() - k - y[ok].sum()) < 1e-7 * n
        assert abs(lam[:, 1].sum() - k - (2 - y[ok].astype(np.int64)).sum()) < 1e-7 * n
    s2 = _staged("steps", n, l, k, 2, 7, 0.01, 1, ",".join(str(x) for x in locs))
    assert int(s2["launches"]) >= len(locs) * 11                     # one launch per round
    assert rel_err(e.gamma, s2["gamma"]) < 1e-11 and rel_err(e.get_lambda(), s2["lam"]) < 1e-11
    np.testing.assert_array_equal(e.counts, s2["counts"])


def test_run_to_run_determinism():
    """The grid-wide sums are integer (fixed-point) additions, so their result does not depend on
    the order in which CTAs arrive: two runs from the same state are bit-identical (the reference
    itself is only reproducible at -nthreads 1)."""
    import terastructure_b200 as ts
    from terastructure_b200 import synth
    n, l, k = 30_000, 64, 6
    theta, beta = synth.psd_params(n, l, k, seed=4)
    g0 = np.random.RandomState(3).gamma(100.0, 0.01, size=(n, k))
    locs = np.random.RandomState(5).randint(0, l, size=200).astype(np.uint32)
    outs = []
    for _ in range(2):
        e = ts.Engine(n, l, k)
        e.synth_bed(11, theta, beta, 0.02)
        e.set_gamma(g0)
        e.steps(locs)
        outs.append((e.gamma, e.get_lambda(), e.counts))
        e.close()
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)


def test_invariants_at_full_bench_size():
    """BASELINE configs[2] at full size: 100 000 individuals x 1 000 000 SNPs, K = 10 (25 GB packed,
    generated on the device, validation set of 5 000 loci x 1 000 individuals sampled with the exact
    RNG stream).  The oracle cannot run this; check size-independent properties after 300 SVI
    iterations that include visits to validation loci:
      sum_k lambda[loc][k][0] - K*eta0 = sum over non-held-out n of y[n], same for (2 - y);
      step counts = visits minus held-out visits;  gamma row sums follow the rho recurrence."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink, synth
    n, l, k = 100_000, 1_000_000, 10
    _, beta = synth.psd_params(1, l, k, seed=1)
    theta = np.random.RandomState(1001).dirichlet(np.full(k, 0.1), size=n)
    e = ts.Engine(n, l, k)
    e.synth_bed(1, theta, beta, 0.0)
    rng = ts.Rng(1234)
    vl, vo, vi = rng.sample_validation(n, l, None)
    assert len(vl) == 5000 and np.all(np.diff(vo) == 1000)
    e.set_validation(vl, vo, vi)
    g0 = rng.init_gamma(n, k)
    e.set_gamma(g0)
    locs = rng.sample_locs(l, 300)
    locs[[7, 150]] = vl[[11, 4000]]          # force two training visits to validation loci
    rounds = e.steps(locs, want_rounds=True)
    assert np.all(rounds == 10)
    cnt = np.zeros(n, np.int64)
    rowsum = g0.sum(1)
    slot = {int(v): j for j, v in enumerate(vl)}   # ~1.5 of 300 random loci are validation loci anyway
    held = {int(x): vi[vo[slot[int(x)]]:vo[slot[int(x)] + 1]] for x in locs if int(x) in slot}
    for loc in locs:
        ok = np.ones(n, bool)
        if int(loc) in held:
            ok[held[int(loc)]] = False
        rho = (2.0 + cnt) ** -0.5
        rowsum = np.where(ok, (1 - rho) * rowsum + rho * (1.0 + 2.0 * l), rowsum)
        cnt += ok
    np.testing.assert_array_equal(e.counts, cnt)
    np.testing.assert_allclose(e.gamma.sum(1), rowsum, rtol=1e-12)
    for loc in (int(locs[0]), int(vl[11]), int(locs[-1])):
        y = plink.unpack(e.get_bed_row(loc)[None, :], n)[0].astype(np.int64)
        ok = np.ones(n, bool)
        if loc in held:
            ok[held[loc]] = False
        lam = e.get_lambda(loc, 1)[0]
        assert abs(lam[:, 0].sum() - k - y[ok].sum()) < 1e-7 * n
        assert abs(lam[:, 1].sum() - k - (2 - y[ok]).sum()) < 1e-7 * n
    s, c, per = e.heldout_ll(first=True)      # 5 * 10^6 held-out genotypes at Ebeta = 1/2
    assert c == 5_000_000 and np.isfinite(s) and -2.0 < s / c < -0.5


# ------------------------------------------------------------------------------------------------
# Every instantiation of the persistent kernel against the oracle (the kernels that
# BENCH/SCALE time -- <10,3> and <10,4> -- and the streaming variant had only been checked through
# sum-over-k invariants).  TSGPU_IPT pins the individuals per thread at oracle-sized inputs.
# ------------------------------------------------------------------------------------------------
def _oracle_vs_engine(y, k, seed, nsteps, extra_locs=(), tol=TIGHT, online_iterations=10):
    """SNPSamplingE-style init (validation set, gamma) + a batch of SVI iterations that includes
    training visits to validation loci, an immediate repeat and an A,B,A pattern; returns the engine."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink
    l, n = y.shape
    rows = plink.pack(y)
    r = ts.Rng(seed)
    vl, vo, vi = r.sample_validation(n, l, rows)
    g0 = r.init_gamma(n, k)
    e = ts.Engine(n, l, k, online_iterations=online_iterations)
    e.load_bed(rows)
    e.set_validation(vl, vo, vi)
    e.set_gamma(g0)
    o = ol.Oracle(y, k, seed, online_iterations=online_iterations)
    np.testing.assert_array_equal(vl, o.validation()[0])
    np.testing.assert_array_equal(g0, o.gamma)
    locs = [o.sample_loc() for _ in range(nsteps)]
    if len(vl):
        locs[1] = int(vl[0])                      # training visit to a validation locus (kv_ok)
    locs[3] = locs[2]                             # same locus twice in a row
    locs[6] = locs[4]                             # A, B, A
    locs += [int(x) for x in extra_locs]
    rounds = e.steps(np.array(locs, np.uint32), want_rounds=True)
    ro = [o.train_loc(x) for x in locs]
    o.flush()
    assert rounds.tolist() == ro
    assert rel_err(e.gamma, o.gamma) < tol
    assert rel_err(e.get_lambda(), o.lam) < tol
    assert rel_err(e.theta, o.theta) < tol
    np.testing.assert_array_equal(e.counts, o.counts)
    # held-out pass through the same kernel in hol mode
    _, cnt, per = e.heldout_ll(False)
    _, a, cnt_o, per_o = o.heldout(first=False, per_locus=True)
    assert cnt == cnt_o
    np.testing.assert_allclose(per, per_o, rtol=1e-10, atol=1e-12)
    assert rel_err(e.get_lambda(), o.lam) < tol
    return e


@pytest.mark.parametrize("k", [10, 20])
@pytest.mark.parametrize("ipt", [1, 2, 3, 4])
def test_forced_instantiation_vs_oracle(ipt, k, monkeypatch):
    """The register-only kernels k_persist<K, I, false> for I = 1..4, ragged N, missing data."""
    from terastructure_b200 import synth
    if ipt == 4 and k == 20:
        pytest.skip("K > 12 keeps at most 3 individuals per thread in registers")
    monkeypatch.setenv("TSGPU_IPT", str(ipt))
    n, l = 1501, 400
    y, _, _ = synth.psd_genotypes(n, l, k, seed=5, missing_rate=0.03)
    e = _oracle_vs_engine(y, k, 31 + ipt, 40)
    assert e.plan[0] == ipt and e.tiers == (-1, 0)


@pytest.mark.parametrize("k,j,grid,block", [(10, 2, 2, 64), (10, 0, 3, 64), (10, 5, 1, 256), (20, 2, 2, 64), (3, 16, 1, 32), (32, 1, 2, 64)])
def test_tiered_kernel_vs_oracle(k, j, grid, block, monkeypatch):
    """k_persist<K, I, true>: register tier + shared-memory tier (j individuals per thread) + streaming tier,
    shaped by the test knobs so that all three tiers hold individuals at N = 1501."""
    from terastructure_b200 import synth
    for name, v in (("TSGPU_IPT", 0), ("TSGPU_TIER_J", j), ("TSGPU_TIER_GRID", grid), ("TSGPU_TIER_BLOCK", block)):
        monkeypatch.setenv(name, str(v))
    n, l = 1501, 400
    y, _, _ = synth.psd_genotypes(n, l, max(k, 2), seed=6, missing_rate=0.03)
    e = _oracle_vs_engine(y, k, 41 + j, 40)
    ipt, g, b = e.plan
    assert (g, b) == (grid, block) and e.tiers == (j, max(0, n - (ipt + j) * grid * block))


def test_config1_shape_vs_oracle():
    """BASELINE configs[1] shape: 10 000 individuals, K = 6 (L cut to 400), 200 SVI iterations."""
    from terastructure_b200 import synth
    y, _, _ = synth.psd_genotypes(10_000, 400, 6, seed=3, missing_rate=0.01)
    e = _oracle_vs_engine(y, 6, 1234, 200)
    assert e.plan[1] == 148


@pytest.mark.parametrize("n,ipt,block", [(100_000, 3, 256), (125_000, 4, 224)])
def test_bench_geometry_vs_oracle(n, ipt, block):
    """The exact kernels BENCH and SCALE time -- k_persist<10,3> at 148 x 256 (100 000 individuals,
    BASELINE configs[2]) and k_persist<10,4> at 148 x 224 (125 000 per GPU, configs[3] / 8) --
    against the sequential oracle: 20 SVI iterations incl. a validation locus, then a held-out pass."""
    from terastructure_b200 import synth
    y, _, _ = synth.psd_genotypes(n, 400, 10, seed=3, missing_rate=0.005)
    e = _oracle_vs_engine(y, 10, 1234, 20)
    assert e.plan == (ipt, 148, block)


def test_tiered_kernel_large_shard_vs_oracle(monkeypatch):
    """The tiered kernel at a size that selects it by itself (beyond 148 x 256 x 4 individuals), with the
    shared-memory tier capped so that part of the shard streams."""
    from terastructure_b200 import synth
    monkeypatch.setenv("TSGPU_TIER_J", "1")
    y, _, _ = synth.psd_genotypes(160_000, 200, 4, seed=9, missing_rate=0.01)
    e = _oracle_vs_engine(y, 4, 7, 12)
    assert e.plan == (2, 148, 256) and e.tiers == (1, 160_000 - 3 * 148 * 256)


@pytest.mark.parametrize("ipt", [1, 3])
def test_single_round_revisit_pattern(ipt, monkeypatch):
    """online_iterations = 1: no grid barrier separates CTA 0's store of a locus' row from another
    CTA's read of it when the locus sequence is A, B, A: rows must not be read early."""
    from terastructure_b200 import synth
    monkeypatch.setenv("TSGPU_IPT", str(ipt))
    y, _, _ = synth.psd_genotypes(30_011, 64, 5, seed=8, missing_rate=0.02)
    _oracle_vs_engine(y, 5, 3, 10, extra_locs=[3, 9, 3, 9, 3, 9, 9, 3] * 6, online_iterations=1)


def test_two_devices_match_one():
    """Real NVLink exchange: one engine per device (ts_comm_connect_local), batches of SNPs issued
    from one host thread per engine, against the single-engine result and the oracle."""
    import threading
    import terastructure_b200 as ts
    from terastructure_b200 import capi
    if ts.lib().ts_device_count() < 2:
        pytest.skip("needs two GPUs")
    c = load_case("synthB")
    n, l, k = c["n"], c["l"], c["k"]
    r = ts.Rng(c["seed"])
    vl, vo, vi = r.sample_validation(n, l, c["rows"])
    g0 = r.init_gamma(n, k)
    o = ol.Oracle(c["y"], k, c["seed"])
    locs = np.array([o.sample_loc() for _ in range(300)], np.uint32)
    locs[10] = vl[0]
    locs[12] = locs[11]
    one = ts.Engine(n, l, k)
    one.load_bed(c["rows"]); one.set_validation(vl, vo, vi); one.set_gamma(g0)
    per = ((n + 1) // 2 + 3) // 4 * 4
    sh = [ts.Engine(n, l, k, device=i, rank=i, nranks=2, n_begin=i * per, n_local=min(per, n - i * per)) for i in range(2)]
    for i, e in enumerate(sh):
        e.load_bed(c["rows"]); e.set_validation(vl, vo, vi); e.set_gamma(g0[i * per:i * per + e.n_local])
    capi.connect_local(sh)
    one.steps(locs)
    res = [None, None]

    def run(i):
        for lo in range(0, len(locs), 100):       # three batches of 100 SNPs per engine
            sh[i].steps(locs[lo:lo + 100])
        sh[i].sync()
        res[i] = sh[i].heldout_ll(False)
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for x in locs:
        o.train_loc(int(x))
    s1, c1, p1 = one.heldout_ll(False)
    _, a, cnt_o, per_o = o.heldout(first=False, per_locus=True)
    np.testing.assert_array_equal(sh[0].get_lambda(), sh[1].get_lambda())   # bit-identical on all ranks
    gs = np.concatenate([e.gamma for e in sh])
    assert rel_err(gs, one.gamma) < 1e-11 and rel_err(gs, o.gamma) < TIGHT
    assert rel_err(sh[0].get_lambda(), o.lam) < TIGHT
    assert res[0][1] + res[1][1] == c1 == cnt_o
    np.testing.assert_allclose(res[0][2] + res[1][2], per_o, rtol=1e-10, atol=1e-12)


def test_cli_side_outputs_match_reference(tmp_path):
    """f4 of SURVEY 8(f): what downstream scripts read besides theta/beta.  The CLI's stdout progress
    lines, param.txt and infer.log of data/run.sh line 1, gammasave.txt (-idfile labels, snp.cc:256-276,
    snpsamplinge.cc:839-861) and beta.txt under -locations-file (snpsamplinge.cc:385-413, :781-798)
    against the reference binary's (tests/golden/fixture_text.json, tools/make_golden.py), byte for
    byte after masking seconds, dates and pids.  Numbers inside gammasave/beta are compared at print
    precision."""
    import json
    import os
    import subprocess
    from conftest import GOLDEN, ROOT, load_fixture_rows, mask_text
    from terastructure_b200 import plink
    gold = json.load(open(os.path.join(GOLDEN, "fixture_text.json")))
    exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
    plink.write_bed(str(tmp_path / "test"), load_fixture_rows(), 200)
    r = subprocess.run([exe, "-file", "test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic", "-nthreads", "1",
                        "-rfreq", "1000", "-seed", "1234", "-label", "test"], cwd=tmp_path, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = tmp_path / "n200-k3-l10000-test-seed1234"
    assert mask_text(r.stdout.decode()) == gold["train_stdout"]
    assert open(d / "param.txt").read() == gold["train_param"]
    assert mask_text(open(d / "infer.log").read()) == gold["train_infer_log"]

    (tmp_path / "ids.txt").write_text("".join("ind%03d\n" % i for i in range(200)))
    (tmp_path / "locs_tab.txt").write_text("".join("%d\trs%d A G\n" % (x, x) for x in (5, 17, 4512, 9999, 17, 300)))
    (tmp_path / "locs_bare.txt").write_text("".join("%d\n" % x for x in (5, 17, 4512, 9999, 17)))

    def numeric_rows(text, skip):
        return [[float(v) for v in line.split("\t")[skip:] if v.strip()] for line in text.splitlines()]

    r = subprocess.run([exe, "-file", "../test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic", "-nthreads", "1",
                        "-compute-beta", "-idfile", "../ids.txt", "-label", "idrun"], cwd=d, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    b = d / "n200-k3-l10000-idrun"
    assert mask_text(r.stdout.decode()) == gold["idrun_stdout"]
    assert open(b / "param.txt").read() == gold["idrun_param"]
    assert mask_text(open(b / "infer.log").read()) == gold["idrun_infer_log"]
    ours, ref = open(b / "gammasave.txt").read(), gold["idrun_gammasave"]
    assert [l.split("\t")[:2] for l in ours.splitlines()] == [l.split("\t")[:2] for l in ref.splitlines()]   # index, label
    assert [l.split("\t")[-1] for l in ours.splitlines()] == [l.split("\t")[-1] for l in ref.splitlines()]   # argmax
    # gamma.txt holds 8 decimals of the GPU run's gamma (within 1e-6 relative of the reference's): same here
    np.testing.assert_allclose(np.array(numeric_rows(ours, 2))[:, :3], np.array(numeric_rows(ref, 2))[:, :3], rtol=1e-6, atol=2e-8)

    for name in ("tab", "bare"):
        r = subprocess.run([exe, "-file", "../test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic", "-nthreads", "1",
                            "-compute-beta", "-locations-file", f"../locs_{name}.txt", "-label", f"loc{name}"],
                           cwd=d, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        b = d / f"n200-k3-l10000-loc{name}"
        assert mask_text(open(b / "infer.log").read()) == gold[f"loc{name}_infer_log"]
        assert mask_text(r.stdout.decode()) == gold[f"loc{name}_stdout"]
        ours, ref = open(b / "beta.txt").read(), gold[f"loc{name}_beta"]
        assert [l.split("\t")[0] for l in ours.splitlines()] == [l.split("\t")[0] for l in ref.splitlines()]
        np.testing.assert_allclose(np.array(numeric_rows(ours, 1)), np.array(numeric_rows(ref, 1)), rtol=1e-6, atol=2e-8)


def test_fanout_loader_matches_rows(tmp_path):
    """ts_load_bed_fanout (one pass over a memory-mapped .bed for all engines of a process, snp.cc:186-229):
    every engine ends up with exactly its byte range of every row -- ragged N, shards that start inside
    the file's rows, more loci than one staging chunk would hold at this pitch."""
    import terastructure_b200 as ts
    from terastructure_b200 import capi, plink
    n, l, k = 1501, 3000, 3
    rs = np.random.RandomState(4)
    y = rs.randint(0, 4, size=(l, n)).astype(np.uint8)
    bed = plink.write_bed(str(tmp_path / "f"), plink.pack(y), n)
    mm = np.memmap(bed, dtype=np.uint8, mode="r", offset=3).reshape(l, (n + 3) // 4)
    per = 752
    eng = [ts.Engine(n, l, k, rank=i, nranks=2, n_begin=i * per, n_local=min(per, n - i * per)) for i in range(2)]
    capi.load_bed_fanout(eng, mm)
    for i, e in enumerate(eng):
        for loc in (0, 1, 777, l - 1):
            got = plink.unpack(e.get_bed_row(loc)[None, :], e.n_local)[0]
            np.testing.assert_array_equal(got, y[loc, i * per:i * per + e.n_local])
    one = ts.Engine(n, l, k)
    capi.load_bed_fanout([one], mm[:1000], loc_begin=0)
    capi.load_bed_fanout([one], mm[1000:], loc_begin=1000)
    for loc in (0, 999, 1000, l - 1):
        np.testing.assert_array_equal(plink.unpack(one.get_bed_row(loc)[None, :], n)[0], y[loc])


@pytest.mark.parametrize("gpus", [1, 2])
def test_cli_synthetic_trajectory(tmp_path, gpus):
    """The C++ driver's own infer()/compute_likelihood()/save_model() (ts_driver.hpp) on the N >= 2000
    validation branch: every report of the reference binary (golden synthB: validation.txt rows and
    gamma_<iter>.txt), through the CLI on a .bed file, on one GPU and sharded over two."""
    import os
    import signal
    import subprocess
    import time
    import terastructure_b200 as ts
    from conftest import ROOT
    from terastructure_b200 import plink
    if ts.lib().ts_device_count() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    c = load_case("synthB")
    g = c["gold"]
    plink.write_bed(str(tmp_path / "d"), c["rows"], c["n"])
    exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
    cmd = [exe, "-file", "d.bed", "-n", str(c["n"]), "-l", str(c["l"]), "-k", str(c["k"]), "-stochastic", "-rfreq", str(c["rfreq"]),
           "-seed", str(c["seed"]), "-label", "g", "-file-suffix", "-gpus", str(gpus)]
    p = subprocess.Popen(cmd, cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    d = tmp_path / f"n{c['n']}-k{c['k']}-l{c['l']}-g-seed{c['seed']}"
    vf = d / "validation.txt"
    nrep = len(g["val_iter"])
    t0 = time.time()
    while p.poll() is None and time.time() - t0 < 240:
        time.sleep(0.2)
        if vf.exists() and sum(1 for _ in open(vf)) >= nrep:
            p.send_signal(signal.SIGTERM)
            break
    assert p.wait(timeout=60) == 0, p.stderr.read()[-2000:]
    val = np.loadtxt(vf)[:nrep]
    assert val[:, 0].astype(int).tolist() == g["val_iter"].tolist()
    assert val[:, 3].astype(int).tolist() == g["val_count"].tolist()
    assert np.max(np.abs(val[:, 2] - g["val_ll"])) < 6e-10
    for it in g["val_iter"]:
        gam = np.loadtxt(d / f"gamma_{it}.txt")
        ref = g[f"gamma_{it}"]
        assert np.all(np.abs(gam - ref) <= 1e-6 * np.abs(ref) + 1.01e-8), it
        theta = np.loadtxt(d / f"theta_{it}.txt")
        np.testing.assert_allclose(theta, ref / ref.sum(1, keepdims=True), rtol=1e-6, atol=1.01e-8)
