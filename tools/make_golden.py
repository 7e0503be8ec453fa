#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/terastructure_ref,
built by `make -C oracle ref` from /root/reference/src) at -nthreads 1.

Runs only in the build container (needs /root/reference).  The .npz files and the data fixture
are committed so that the tests can run where /root/reference does not exist (the GPU box).

Cases
  fixture   the reference's bundled data/test.{bed,bim,fam} (N=200, L=10000, K=3), data/run.sh
            line 1 (-seed 1234 -rfreq 1000) to its own stop, then line 2 (-compute-beta).
            Also stores the reference's shipped data/output_theta.txt.
  synthA    N=600  L=3000 K=4, 2% missing, data seed 11, -seed 77  -rfreq 500  (N<2000 branch)
  synthB    N=2400 L=1600 K=5, no missing, data seed 5,  -seed 9   -rfreq 400  (N>=2000 branch)
            for the synthetic cases the reference is stopped (SIGTERM) after a few reports;
            -file-suffix keeps gamma_<iter>.txt of every report.
"""
import glob
import os
import shutil
import signal
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from terastructure_b200 import plink, synth  # noqa: E402

REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "terastructure_ref")
OUT = os.path.join(ROOT, "tests", "golden")


def read_validation(path):
    rows = [l.split() for l in open(path) if l.strip()]
    return (np.array([int(r[0]) for r in rows], np.int64), np.array([float(r[2]) for r in rows]),
            np.array([int(r[3]) for r in rows], np.int64))


def run_fixture():
    tmp = tempfile.mkdtemp(prefix="tsgold.")
    for e in ("bed", "bim", "fam"):
        shutil.copy(f"{REF}/data/test.{e}", tmp)
    subprocess.run([BIN, "-file", "test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic",
                    "-nthreads", "1", "-rfreq", "1000", "-seed", "1234", "-label", "test"],
                   cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = os.path.join(tmp, "n200-k3-l10000-test-seed1234")
    it, ll, cnt = read_validation(os.path.join(d, "validation.txt"))
    theta = np.loadtxt(os.path.join(d, "theta.txt"))
    gamma = np.loadtxt(os.path.join(d, "gamma.txt"))
    shipped = np.loadtxt(f"{REF}/data/output_theta.txt")
    assert open(os.path.join(d, "theta.txt")).read() == open(f"{REF}/data/output_theta.txt").read()
    subprocess.run([BIN, "-file", "../test.bed", "-n", "200", "-l", "10000", "-k", "3", "-stochastic",
                    "-nthreads", "1", "-compute-beta"],
                   cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    bdir = [p for p in glob.glob(os.path.join(d, "n200-k3-l10000*")) if os.path.isdir(p)][0]
    beta = np.loadtxt(os.path.join(bdir, "beta.txt"))[:, 1:]
    np.savez_compressed(os.path.join(OUT, "fixture.npz"), val_iter=it, val_ll=ll, val_count=cnt,
                        theta=theta, gamma=gamma, shipped_theta=shipped, beta=beta)
    # the data fixture itself (N=200 x L=10000, 500 KB packed); .bim/.fam are only line-counted
    shutil.copy(f"{REF}/data/test.bed", os.path.join(OUT, "fixture_n200_l10000.bed"))
    shutil.rmtree(tmp)
    print("fixture: reports", it.tolist(), "final ll", ll[-1])


def run_synth(name, n, l, k, data_seed, miss, seed, rfreq, nreports):
    y, _, _ = synth.psd_genotypes(n, l, k, seed=data_seed, missing_rate=miss)
    tmp = tempfile.mkdtemp(prefix="tsgold.")
    plink.write_bed(os.path.join(tmp, "d"), plink.pack(y), n)
    p = subprocess.Popen([BIN, "-file", "d.bed", "-n", str(n), "-l", str(l), "-k", str(k), "-stochastic",
                          "-nthreads", "1", "-rfreq", str(rfreq), "-seed", str(seed), "-label", "g",
                          "-file-suffix"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = os.path.join(tmp, f"n{n}-k{k}-l{l}-g-seed{seed}")
    vf = os.path.join(d, "validation.txt")
    t0 = time.time()
    while p.poll() is None and time.time() - t0 < 1800:
        time.sleep(0.5)
        if os.path.exists(vf) and sum(1 for _ in open(vf)) >= nreports + 1:
            p.send_signal(signal.SIGTERM)
            p.wait()
            break
    it, ll, cnt = read_validation(vf)
    it, ll, cnt = it[:nreports + 1], ll[:nreports + 1], cnt[:nreports + 1]
    gam = {f"gamma_{i}": np.loadtxt(os.path.join(d, f"gamma_{i}.txt")) for i in it}
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), val_iter=it, val_ll=ll, val_count=cnt,
                        shape=np.array([n, l, k]), data_seed=data_seed, missing_rate=miss, seed=seed,
                        rfreq=rfreq, **gam)
    shutil.rmtree(tmp)
    print(name, "reports", it.tolist(), ll.tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["fixture", "synthA", "synthB"]
    if "fixture" in which:
        run_fixture()
    if "synthA" in which:
        run_synth("synthA", 600, 3000, 4, 11, 0.02, 77, 500, 3)
    if "synthB" in which:
        run_synth("synthB", 2400, 1600, 5, 5, 0.0, 9, 400, 2)
