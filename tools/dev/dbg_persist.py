import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import terastructure_b200 as ts
from terastructure_b200 import plink, synth
def run(k, n, staged, nsteps):
    l = 400
    y, _, _ = synth.psd_genotypes(n, l, max(k, 2), seed=3, missing_rate=0.03)
    rows = plink.pack(y)
    if staged: os.environ["TSGPU_PATH"] = "staged"
    else: os.environ.pop("TSGPU_PATH", None)
    e = ts.Engine(n, l, k)
    e.load_bed(rows)
    r = ts.Rng(21)
    vl, vo, vi = r.sample_validation(n, l, rows)
    e.set_validation(vl, vo, vi)
    e.set_gamma(r.init_gamma(n, k))
    locs = r.sample_locs(l, nsteps)
    rounds = e.steps(locs, want_rounds=True)
    return locs, rounds, e.get_lambda(), e.gamma
for k, n in ((10, 300),):
    for nsteps in (1, 3):
        la, ra, lama, ga = run(k, n, True, nsteps)
        lb, rb, lamb, gb = run(k, n, False, nsteps)
        dl = np.abs(lama - lamb) / np.abs(lama)
        dg = np.abs(ga - gb) / np.abs(ga)
        print(f"K={k} N={n} steps={nsteps} rounds {ra.tolist()} vs {rb.tolist()} lam relerr max {dl.max():.3e} gamma relerr max {dg.max():.3e}")
        if dl.max() > 1e-9:
            loc = la[0]
            print("  lam staged ", lama[loc].ravel())
            print("  lam persist", lamb[loc].ravel())
