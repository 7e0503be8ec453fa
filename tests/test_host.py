"""CPU tests of the product's host side: the C ABI library loads and exports every symbol
include/tsgpu.h declares, the GSL-exact host RNG / validation sampler / init_gamma agree with
the oracle and the golden vectors, the packed-genotype helpers round-trip, and compute entry
points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT, load_case


def test_cabi_exports_match_header():
    import terastructure_b200 as ts
    from terastructure_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "tsgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ts_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(ts.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in tsgpu.h but not exported"
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    assert ts.lib().ts_abi_version() == 1


def test_mt19937_known_answer():
    """MT19937 seeded with 4357 (GSL's default seed): 10000th output, and seed 5489's first."""
    import terastructure_b200 as ts
    r = ts.Rng(0)
    rs = np.random.RandomState(4357)
    ref = rs.randint(0, 2 ** 32, size=1000, dtype=np.uint64)
    got = [r.get() for _ in range(1000)]
    assert got == ref.tolist()
    r = ts.Rng(5489)
    assert r.get() == 3499211612  # reference value of the MT19937 paper's init_genrand(5489)


def test_host_init_matches_oracle_and_golden(fixture_case):
    import terastructure_b200 as ts
    c = fixture_case
    r = ts.Rng(c["seed"])
    vl, vo, vi = r.sample_validation(c["n"], c["l"], c["rows"])
    g0 = r.init_gamma(c["n"], c["k"])
    locs = r.sample_locs(c["l"], 2000)
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    ovl, ovo, ovi = o.validation()
    np.testing.assert_array_equal(vl, ovl)
    np.testing.assert_array_equal(vo, ovo)
    np.testing.assert_array_equal(vi, ovi)
    np.testing.assert_array_equal(g0, o.gamma)
    assert locs.tolist() == [o.sample_loc() for _ in range(2000)]
    assert locs[:12].tolist() == [4512, 5810, 6508, 3177, 8093, 7291, 2829, 6486, 2733, 7778, 3857, 648]


@pytest.mark.parametrize("name", ["synthA", "synthB"])
def test_host_init_synthetic(name):
    """Missing genotypes are rejected by the sampler (kv_ok); N>=2000 uses N/100 per locus."""
    import terastructure_b200 as ts
    c = load_case(name)
    r = ts.Rng(c["seed"])
    vl, vo, vi = r.sample_validation(c["n"], c["l"], c["rows"])
    g0 = r.init_gamma(c["n"], c["k"])
    o = ol.Oracle(c["y"], c["k"], c["seed"])
    ovl, ovo, ovi = o.validation()
    np.testing.assert_array_equal(vl, ovl)
    np.testing.assert_array_equal(vi, ovi)
    np.testing.assert_array_equal(g0, o.gamma)
    np.testing.assert_allclose(g0, c["gold"]["gamma_0"], atol=5.1e-9)
    per = c["n"] // 10 if c["n"] < 2000 else c["n"] // 100
    assert np.all(np.diff(vo) == per)
    for j in range(len(vl)):
        assert np.all(c["y"][vl[j], vi[vo[j]:vo[j + 1]]] != 3)


def test_plink_roundtrip_and_codes():
    from terastructure_b200 import plink
    rs = np.random.RandomState(1)
    for n in (1, 3, 4, 5, 203):
        y = rs.randint(0, 4, size=(17, n)).astype(np.uint8)
        rows = plink.pack(y)
        assert rows.shape == (17, (n + 3) // 4)
        np.testing.assert_array_equal(plink.unpack(rows, n), y)
        np.testing.assert_array_equal(ol.decode_bed(rows.tobytes(), n, 17), y)
    # snp.cc:203-216: 00->0, 01->missing, 10->1, 11->2 (low bits first)
    assert plink.unpack(np.array([[0b11100100]], np.uint8), 4).tolist() == [[0, 3, 1, 2]]


def test_read_bed_checks(tmp_path):
    from terastructure_b200 import plink
    y = np.random.RandomState(0).randint(0, 3, size=(6, 9)).astype(np.uint8)
    bed = plink.write_bed(str(tmp_path / "t"), plink.pack(y), 9)
    np.testing.assert_array_equal(plink.unpack(plink.read_bed(bed, 9, 6), 9), y)
    with pytest.raises(ValueError, match="-l input"):
        plink.read_bed(bed, 9, 7)
    with pytest.raises(ValueError, match="-n input"):
        plink.read_bed(bed, 8, 6)
    raw = bytearray(open(bed, "rb").read()); raw[2] = 0
    open(bed, "wb").write(raw)
    with pytest.raises(ValueError, match="individual major"):
        plink.read_bed(bed, 9, 6)


def test_compute_fails_loudly_without_gpu():
    import terastructure_b200 as ts
    if ts.lib().ts_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ts.TsError, match="no CUDA device"):
        ts.Engine(16, 4, 2)


def test_argument_validation():
    import terastructure_b200 as ts
    from terastructure_b200 import capi
    cfg = capi.TsConfig()
    ts.lib().ts_config_defaults(C.byref(cfg), 100, 10, 3)
    assert (cfg.alpha, cfg.eta0, cfg.nodetau0, cfg.nodekappa, cfg.online_iterations) == (1 / 3, 1.0, 2.0, 0.5, 10)
    h = C.c_void_p()
    cfg.k = 33
    assert ts.lib().ts_create(C.byref(cfg), C.byref(h)) == -1 and b"K=33" in ts.lib().ts_last_error()
    cfg.k = 3; cfg.n_begin = 2
    assert ts.lib().ts_create(C.byref(cfg), C.byref(h)) == -1
    assert ts.lib().ts_steps(None, None, 0, 0, None) == -1


def _cli():
    exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
    assert os.path.exists(exe), "CLI not built (make -C terastructure_b200/csrc)"
    return exe


def test_cli_flags_without_gpu(tmp_path):
    """The front end keeps the reference's flag handling (src/main.cc:84-187): usage with no
    arguments (exit -1), -help (exit 0), unknown option -> abort (the reference: assert(0)),
    -batch prints and exits 0; without a GPU it refuses to run rather than falling back."""
    import subprocess
    exe = _cli()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and "terastructure [OPTIONS]" in r.stdout
    r = subprocess.run([exe, "-help"], capture_output=True, text=True)
    assert r.returncode == 0 and "-rfreq" in r.stdout
    r = subprocess.run([exe, "-n", "5", "-bogus"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "unknown option -bogus" in r.stdout
    r = subprocess.run([exe, "-batch"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "batch option currently not available" in r.stdout
    r = subprocess.run([exe, "-file", "x.bed", "-n", "10", "-l", "5", "-k", "2", "-use-test-set"],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "-use-test-set" in r.stderr
    import terastructure_b200 as ts
    if ts.lib().ts_device_count() == 0:
        r = subprocess.run([exe, "-file", "x.bed", "-n", "10", "-l", "5", "-k", "2"], capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode != 0 and "no CPU path" in r.stderr
        assert not any(p.is_dir() for p in tmp_path.iterdir())  # no output directory was created


def test_ziggurat_tables_are_reproducible(tmp_path):
    """The GSL-exact Gaussian ziggurat tables (product copy and oracle-shim copy) are regenerated from
    first principles by tools/gen_zig_tables.py: the committed files must equal a fresh run."""
    import subprocess
    import sys
    pytest.importorskip("mpmath")
    gen = os.path.join(ROOT, "tools", "gen_zig_tables.py")
    for committed, prefix in ((os.path.join(ROOT, "terastructure_b200", "csrc", "zig_tables.inc"), "tszig"),
                              (os.path.join(ROOT, "oracle", "gsl_shim", "zig_tables.h"), "zig")):
        out = tmp_path / os.path.basename(committed)
        subprocess.run([sys.executable, gen, str(out), prefix], check=True, stdout=subprocess.DEVNULL)
        assert open(out).read() == open(committed).read()


class _OracleEngine:
    """Stand-in for capi.Engine backed by the CPU oracle: lets the host-side driver logic of
    SNPSamplingE (report cadence, _iter bookkeeping, stopping rule, RNG pre-draws) run without a
    GPU.  Test infrastructure only."""

    def __init__(self, case):
        self.o = ol.Oracle(case["y"], case["k"], case["seed"])
        self.n_local, self.nval = case["n"], 0

        class _Cfg:
            n_begin = 0
        self.cfg = _Cfg()

    def set_validation(self, vl, vo, vi):
        ovl, _, ovi = self.o.validation()
        np.testing.assert_array_equal(vl, ovl)
        np.testing.assert_array_equal(vi, ovi)
        self.nval = len(vl)

    def set_gamma(self, g):
        np.testing.assert_array_equal(g, self.o.gamma)

    def steps(self, locs, hol_mode=False, want_rounds=False):
        for loc in locs:
            self.o.train_loc(int(loc))

    def heldout_ll(self, first=False):
        _, a, cnt, per = self.o.heldout(first=first, per_locus=True)
        return float(per.sum()), cnt, per

    def sync(self):
        pass

    gamma = property(lambda s: (s.o.flush(), s.o.gamma)[1])
    theta = property(lambda s: (s.o.flush(), s.o.theta)[1])


def test_driver_host_logic_against_golden(fixture_case, tmp_path):
    """SNPSamplingE's host loop (Python mirror of snpsamplinge.cc:417-544) on top of the oracle:
    same report iterations, counts and stop (9050) as the reference binary, validation.txt and
    theta.txt written in the reference's formats."""
    import terastructure_b200 as ts
    c = fixture_case
    g = c["gold"]
    env = ts.Env(c["n"], c["k"], c["l"], seed=c["seed"], rfreq=c["rfreq"], outdir=str(tmp_path))
    s = ts.SNPSamplingE(env, None, engine=_OracleEngine(c))
    s.infer()
    assert s.stopped and s._iter == 9050
    assert [r[0] for r in s.validation_rows] == g["val_iter"].tolist()
    assert [r[3] for r in s.validation_rows] == g["val_count"].tolist()
    np.testing.assert_allclose([r[2] for r in s.validation_rows], g["val_ll"], atol=6e-10, rtol=0)
    val = np.loadtxt(tmp_path / "validation.txt")
    assert val.shape == (10, 5) and val[:, 0].astype(int).tolist() == g["val_iter"].tolist()
    import hashlib
    assert hashlib.md5(open(tmp_path / "theta.txt", "rb").read()).hexdigest() == "ae1136d8769318e9840b1f7dd1d1ac53"


def test_shard_plan_geometry():
    """ts_plan_shard (host arithmetic behind ts_create): the persistent kernel's geometry covers the
    shard, respects the per-(K, I) thread caps, and reproduces the choices the round-1 measurements
    were made with (profiles/r1_summary.md)."""
    import terastructure_b200 as ts
    sms = 148
    assert ts.plan_shard(100000, 10, sms) == (3, 148, 256)      # BASELINE configs[2] on one B200
    assert ts.plan_shard(125000, 10, sms) == (4, 148, 224)      # configs[3] over 8 GPUs
    assert ts.plan_shard(60000, 10, sms)[0] == 2 and ts.plan_shard(80000, 10, sms)[0] == 3
    assert ts.plan_shard(200, 3, sms) == (1, 4, 64)             # the reference's bundled data set
    # beyond the register-resident capacity: 2 individuals per thread in registers, more in shared memory
    assert ts.plan_shard(400000, 10, sms) == (2, 148, 256) and ts.plan_tiers(400000, 10, sms) == (9, 0)
    j, streamed = ts.plan_tiers(1_000_000, 10, sms)
    assert j == 10 and streamed == 1_000_000 - 12 * 148 * 256   # 227 KB of shared memory per CTA: the control path's 20 KB table + 10 more
    assert ts.plan_tiers(100000, 10, sms) == (0, 0)
    for k in (1, 2, 6, 10, 12, 13, 16, 20, 21, 32):
        for n in (1, 3, 31, 33, 200, 4097, 9999, 37888, 37889, 56000, 75776, 99999, 113664, 113665, 151552, 151553, 10 ** 6):
            ipt, grid, block = ts.plan_shard(n, k, sms)
            j, streamed = ts.plan_tiers(n, k, sms)
            assert 1 <= grid <= sms and block % 32 == 0 and 32 <= block <= 512
            assert 1 <= ipt <= 4 and 0 <= j <= 16
            assert grid * block * (ipt + j) + streamed >= n, (k, n, ipt, grid, block, j, streamed)
            if streamed:
                assert grid * block * (ipt + j) + streamed == n and grid == sms
    with pytest.raises(ts.TsError):
        ts.plan_shard(100, 33, sms)


def test_fixed_point_words_host_check(tmp_path):
    """ts_fixed.cuh compiled for the host (the same source the persistent kernel inlines): split /
    normalize / arrival counts / fold over ranks / to_double are exact, order-independent, survive the
    wrap-around of the monotonic words, and the >= 2^52 low-word sum over 4+ GPUs is carried."""
    exe = str(tmp_path / "fx_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "fx_check.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def test_exp_digamma_host_model_vs_mpmath(tmp_path):
    """ts_expsi.cuh compiled for the host (the kernel's own source; the MUFU reciprocal seed modelled
    by a 20-bit reciprocal): f = exp(digamma), fast_rcp and exp_neg against mpmath."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    exe = str(tmp_path / "expsi_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "expsi_check.cpp")], check=True)
    out = subprocess.run([exe, "1200"], capture_output=True, text=True, check=True).stdout
    worst = {"large": 0.0, "mid": 0.0, "small": 0.0, "rcp": 0.0, "exp": 0.0}
    for line in out.splitlines():
        x, f, r, e = (float.fromhex(t) for t in line.split())
        X = mp.mpf(x)
        ref = mp.exp(mp.digamma(X))
        err = abs(float((mp.mpf(f) - ref) / ref)) if ref > mp.mpf("1e-290") else abs(f)
        key = "large" if x >= 8 else ("mid" if x >= 0.03 else "small")
        worst[key] = max(worst[key], err)
        worst["rcp"] = max(worst["rcp"], abs(float(mp.mpf(r) * X - 1)))
        y = x if x < 750.0 else 750.0 + 1e-7 * x
        if y <= 700.0:
            worst["exp"] = max(worst["exp"], abs(float(mp.mpf(e) / mp.exp(-mp.mpf(y)) - 1)))
        else:
            assert e == 0.0
    assert worst["large"] < 5e-16 and worst["mid"] < 1e-14 and worst["small"] < 3e-13, worst
    assert worst["rcp"] < 2.3e-16 and worst["exp"] < 1e-14, worst


def test_control_path_table_is_reproducible(tmp_path):
    """The committed coefficient table (terastructure_b200/csrc/ts_ftab.inc) is exactly what tools/gen_ftab.py writes."""
    pytest.importorskip("mpmath")
    out = str(tmp_path / "ts_ftab.inc")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_ftab.py"), out], check=True, capture_output=True)
    assert open(out).read() == open(os.path.join(ROOT, "terastructure_b200", "csrc", "ts_ftab.inc")).read()


def test_control_path_table_vs_mpmath(tmp_path):
    """ts_ftab.cuh compiled for the host (the kernel's own source and coefficient table, tools/gen_ftab.py): the
    piecewise polynomials for f = exp(digamma) and 1/f that the control warp turns a lambda row into b with
    (estimate_beta, snpsamplinge.cc:279-296) against mpmath over the whole domain [1, 2^24), interval boundaries
    included; arguments outside are refused (the kernel falls back to ts_expsi.cuh there); b = f(x) / f(s) through
    either path agrees with mpmath."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    exe = str(tmp_path / "ftab_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "ftab_check.cpp")], check=True)
    out = subprocess.run([exe, "2400"], capture_output=True, text=True, check=True).stdout
    wf = wg = wb = 0.0
    n_in = 0
    for line in out.splitlines():
        x, inside, f, g, t, b = line.split()
        x, inside = float.fromhex(x), int(inside)
        assert inside == (1.0 <= x < 2.0 ** 24), x
        X = mp.mpf(x)
        F = mp.exp(mp.digamma(X))
        B = F / mp.exp(mp.digamma(2 * X + mp.mpf("0.25")))
        wb = max(wb, abs(float(mp.mpf(float.fromhex(b)) / B - 1)))
        if not inside:
            continue
        n_in += 1
        assert -1.0 <= float.fromhex(t) < 1.0
        wf = max(wf, abs(float(mp.mpf(float.fromhex(f)) / F - 1)))
        wg = max(wg, abs(float(mp.mpf(float.fromhex(g)) * F - 1)))
    assert n_in > 2000 and wf < 6e-16 and wg < 8e-16 and wb < 2e-14, (n_in, wf, wg, wb)


@pytest.mark.parametrize("n,l,seed", [(2500, 1800, 99), (150, 1200, 3)])
def test_validation_sampler_with_missing_genotypes(n, l, seed):
    """set_validation_sample (cc:196-224) with 3 % missing genotypes (kv_ok rejects them, which shifts
    the RNG stream): the host sampler (one reusable bitmap, blocks permuted in place) against the
    oracle's restatement, for both N >= 2000 (N/100 per locus) and N < 2000 (N/10)."""
    import terastructure_b200 as ts
    from terastructure_b200 import plink
    rs = np.random.RandomState(seed)
    y = rs.randint(0, 3, size=(l, n)).astype(np.uint8)
    y[rs.rand(l, n) < 0.03] = 3
    o = ol.Oracle(y, 3, seed)
    ovl, ovo, ovi = o.validation()
    vl, vo, vi = ts.Rng(seed).sample_validation(n, l, plink.pack(y))
    assert np.array_equal(vl, ovl) and np.array_equal(vo, ovo) and np.array_equal(vi, ovi)


def test_validation_sampler_matches_stream_emulation():
    """300 validation loci (long cycles in the in-place block permutation): every locus keeps its own
    held-out individuals, and the RNG stream ends where a draw-by-draw emulation of
    set_validation_sample (cc:196-224) ends."""
    import terastructure_b200 as ts
    n, l, seed = 3000, 60000, 5
    r = ts.Rng(seed)
    vl, vo, vi = r.sample_validation(n, l, None)
    q = ts.Rng(seed)
    h, nlocs = n // 100, int(l * 0.005)
    taken, out = set(), {}
    while len(out) < nlocs:
        loc = int(q.sample_locs(l, 1)[0])
        if loc in taken:
            continue
        taken.add(loc)
        s = set()
        while len(s) < h:
            s.add(int(q.sample_locs(n, 1)[0]))
        out[loc] = sorted(s)
    assert list(vl) == sorted(out)
    for i, loc in enumerate(sorted(out)):
        assert list(vi[vo[i]:vo[i + 1]]) == out[loc]
    assert int(r.sample_locs(1000, 1)[0]) == int(q.sample_locs(1000, 1)[0])


def test_transposed_reduction_slot_mapping(tmp_path):
    """tsfx::tr_slot (which statistics a lane holds after the kernel's transposed warp reduction)
    against a lock-step host emulation of the shuffle network, for every V = 2K, K = 1..32 (the
    mapping uses nominal, not live, counts: V = 10, 20, 26, 40 were wrong once)."""
    exe = str(tmp_path / "tr_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "terastructure_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "tr_check.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout[-2000:]
