# soak test: a long SVI run at bench scale through the Python mirror (reports, hol passes, stop rule)
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import terastructure_b200 as ts
from terastructure_b200 import synth
n, l, k = 100_000, 20_000, 10
theta, beta = synth.psd_params(n, l, k, seed=1)
e = ts.Engine(n, l, k)
e.synth_bed(1, theta, beta, 0.005)
rows = np.stack([e.get_bed_row(i) for i in range(l)])          # the data set as the host would hold it
e.close()
env = ts.Env(n, k, l, seed=1234, rfreq=20000)
t0 = time.time()
s = ts.SNPSamplingE(env, rows)
print("init %.1fs, initial LL %.6f" % (time.time() - t0, s.validation_rows[0][2]))
t0 = time.time()
s.infer(max_iter=200_000)
dt = time.time() - t0
for r in s.validation_rows: print("iter %7d  LL %.6f  count %d" % (r[0], r[2], r[3]))
print("%d iterations in %.1fs (%.1f us each incl. reports); stopped=%s" % (s._iter, dt, 1e6 * dt / s._iter, s.stopped))
th = s.engine.theta
# recovery of the simulated ancestry: best column matching by correlation
import itertools
c = np.corrcoef(th.T, theta.T)[:k, k:]
print("mean best-match correlation theta vs truth: %.3f" % np.mean(np.max(c, axis=1)))
