import os, sys
import numpy as np
os.environ["TSGPU_TRACE"] = "1"
# the phase stamps are compiled into the developer library only (make -C terastructure_b200/csrc trace; K = 10)
os.environ.setdefault("TSGPU_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "terastructure_b200", "lib", "libtsgpu_trace.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import terastructure_b200 as ts
from terastructure_b200 import synth
n, l, k = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 20000, 10
_, beta = synth.psd_params(1, l, k, seed=1)
theta = np.random.RandomState(5).dirichlet(np.full(k, 0.1), size=n)
e = ts.Engine(n, l, k)
e.synth_bed(1, theta, beta, 0.0)
r = ts.Rng(1234)
vl, vo, vi = r.sample_validation(n, l, None)
e.set_validation(vl, vo, vi)
e.set_gamma(r.init_gamma(n, k))
e.steps(r.sample_locs(l, 200)); e.sync()
e.timer_start(); e.steps(r.sample_locs(l, 64)); ms = e.timer_stop()
print("64 SNPs: %.1f us per SNP" % (1e3 * ms / 64))
t = e.debug_trace().astype(np.float64)
it = t[8:60]
def d(a, b): return np.mean(it[:, b] - it[:, a])
print("cycles: item-start b-compute+sync %.0f" % d(0, 1))
names = ["E-step", "tr_reduce+st", "sync1", "2nd-level+publish", "poll", "lambda/b", "sync2"]
tot = np.zeros(7)
for x in range(10):
    base = 2 + 8 * x
    prev = (2 + 8 * (x - 1) + 6) if x else 1
    seg = [d(prev, base)] + [d(base + j, base + j + 1) for j in range(6)]
    tot += np.array(seg)
    print("round %d: " % x + "  ".join("%s %.0f" % (nm, v) for nm, v in zip(names, seg)))
print("mean/round: " + "  ".join("%s %.0f" % (nm, v / 10) for nm, v in zip(names, tot)) + "  total %.0f" % (tot.sum() / 10))
print("gamma phase %.0f  end-sync %.0f  whole item %.0f" % (d(2 + 8 * 9 + 6, 100), d(100, 101), np.mean(it[1:, 0] - it[:-1, 0])))
