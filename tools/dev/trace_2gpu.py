import os, sys, threading
import numpy as np
os.environ["TSGPU_TRACE"] = "1"
# the phase stamps are compiled into the developer library only (make -C terastructure_b200/csrc trace; K = 10)
os.environ.setdefault("TSGPU_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "terastructure_b200", "lib", "libtsgpu_trace.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import terastructure_b200 as ts
from terastructure_b200 import synth, capi
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per, l, k = 125000, 20000, 10
n = per * ng
_, beta = synth.psd_params(1, l, k, seed=1)
engs = []
for r in range(ng):
    theta = np.random.RandomState(5 + r).dirichlet(np.full(k, 0.1), size=per)
    e = ts.Engine(n, l, k, device=r, rank=r, nranks=ng, n_begin=r * per, n_local=per)
    e.synth_bed(1, theta, beta, 0.0)
    engs.append(e)
rng = ts.Rng(1234)
vl, vo, vi = rng.sample_validation(n, l, None)
g0 = rng.init_gamma(n, k)
for r, e in enumerate(engs):
    e.set_validation(vl, vo, vi); e.set_gamma(g0[r * per:(r + 1) * per])
capi.connect_local(engs)
def run(locs, out, r):
    engs[r].timer_start(); engs[r].steps(locs); out[r] = engs[r].timer_stop()
def both(locs):
    out = [0] * ng
    th = [threading.Thread(target=run, args=(locs, out, r)) for r in range(ng)]
    [t.start() for t in th]; [t.join() for t in th]
    return out
both(rng.sample_locs(l, 200))
ms = both(rng.sample_locs(l, 64))
print("64 SNPs on %d GPUs: %s us per SNP" % (ng, ["%.1f" % (1e3 * m / 64) for m in ms]))
t = engs[0].debug_trace().astype(np.float64)
it = t[8:60]
def d(a, b): return np.mean(it[:, b] - it[:, a])
names = ["E-step", "tr_reduce+st", "sync1", "2nd-level+publish", "poll(+exchange)", "lambda/b", "sync2"]
tot = np.zeros(7)
for x in range(10):
    base = 2 + 8 * x
    prev = (2 + 8 * (x - 1) + 6) if x else 1
    tot += np.array([d(prev, base)] + [d(base + j, base + j + 1) for j in range(6)])
print("mean/round: " + "  ".join("%s %.0f" % (nm, v / 10) for nm, v in zip(names, tot)) + "  total %.0f" % (tot.sum() / 10))
print("whole item %.0f cycles" % np.mean(it[1:, 0] - it[:-1, 0]))
