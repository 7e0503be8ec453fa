"""TEST INFRASTRUCTURE: run a computation on the STAGED path (one launch per round; an independent
implementation of the same mathematics, compiled only into lib/libtsgpu_staged.so) in its own process
and dump the state, so that tests can cross-check the product's persistent kernel against it.

  python tests/staged_helper.py trajectory <case> <last_iter> <out.npz>
  python tests/staged_helper.py steps <n> <l> <k> <synth_seed> <bed_seed> <missing> <gamma_seed> <loc,loc,...> <out.npz>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TSGPU_LIB"] = os.path.join(ROOT, "terastructure_b200", "lib", "libtsgpu_staged.so")
os.environ["TSGPU_PATH"] = "staged"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import terastructure_b200 as ts  # noqa: E402


def main():
    mode = sys.argv[1]
    if mode == "trajectory":
        from conftest import load_case
        c = load_case(sys.argv[2])
        last, out = int(sys.argv[3]), sys.argv[4]
        env = ts.Env(c["n"], c["k"], c["l"], seed=c["seed"], rfreq=c["rfreq"])
        s = ts.SNPSamplingE(env, c["rows"])
        s.infer(max_iter=last)
        np.savez(out, val_iter=[r[0] for r in s.validation_rows], val_ll=[r[2] for r in s.validation_rows],
                 gamma=s.engine.gamma, lam=s.engine.get_lambda(), launches=s.engine.launch_count)
    elif mode == "steps":
        from terastructure_b200 import synth
        n, l, k, synth_seed, bed_seed = (int(v) for v in sys.argv[2:7])
        missing, gamma_seed = float(sys.argv[7]), int(sys.argv[8])
        locs = np.array([int(v) for v in sys.argv[9].split(",")], np.uint32)
        out = sys.argv[10]
        theta, beta = synth.psd_params(n, l, k, seed=synth_seed)
        g0 = np.random.RandomState(gamma_seed).gamma(100.0, 0.01, size=(n, k))
        e = ts.Engine(n, l, k)
        e.synth_bed(bed_seed, theta, beta, missing)
        e.set_gamma(g0)
        before = e.launch_count
        rounds = e.steps(locs, want_rounds=True)
        np.savez(out, gamma=e.gamma, lam=e.get_lambda(), counts=e.counts, rounds=rounds, launches=e.launch_count - before)
    else:
        raise SystemExit("unknown mode " + mode)


if __name__ == "__main__":
    main()
