"""Host-side mirror of the reference's SNPSamplingE driver (src/snpsamplinge.cc) on top of
the C ABI.  Same names, same argument meaning, same RNG stream and report/stop rules; the
per-SNP work goes to the B200 through libtsgpu.so (no fallback)."""
import math
import os
import time

import numpy as np

from . import capi


class Env:
    """Hyper-parameters and run options (src/env.hh:50-313, defaults :200-249; flag names
    from src/main.cc:84-187)."""

    def __init__(self, n, k, l, *, seed=0.0, rfreq=100000, nthreads=6, stop_threshold=1e-5,
                 label="", compute_beta=False, file_suffix=False, force=False, outdir=None):
        self.n, self.k, self.l = n, k, l
        self.t = 2
        self.seed = float(seed)
        self.reportfreq = rfreq
        self.nthreads = nthreads  # accepted for CLI compatibility; the GPU replaces the pool
        self.stop_threshold = stop_threshold
        self.label = label
        self.compute_beta = compute_beta
        self.file_suffix = file_suffix
        self.force = force
        self.meanchangethresh = 0.001
        self.alpha = 1.0 / k
        self.validation_ratio = 0.005
        self.eta0 = self.eta1 = 1.0
        self.nodetau0 = 1.0
        self.nodekappa = 0.5
        self.online_iterations = 100 if compute_beta else 10  # snpsamplinge.cc:75
        self.use_validation_stop = True
        self.terminate = False
        self.outdir = outdir  # None: keep results in memory only


def _fmt_rows(a):
    return "".join("".join("%.8f\t" % v for v in row) + "\n" for row in a)


class SNPSamplingE:
    """SNPSamplingE(env, snp) + infer() (snpsamplinge.cc:6-120, :417-459).

    `bed_rows`: SNP-major PLINK-packed genotype rows [L, ceil(N/4)] (plink.read_bed).
    Multi-GPU: pass this rank's shard (rank, nranks, device) and an `allgather(obj)->list`
    callable; every rank runs the same host loop on the same RNG stream."""

    def __init__(self, env, bed_rows, *, device=0, rank=0, nranks=1, allgather=None, gamma0=None, engine=None,
                 connect=True):
        self.env = env
        self._n, self._k, self._l = env.n, env.k, env.l
        self._iter = 0
        self._start_time = time.time()
        self._prev_h = -2147483647.0
        self._max_h = -2147483647.0
        self._nh = 0
        self.rank, self.nranks = rank, nranks
        self._allgather = allgather
        self.validation_rows = []  # (iter, secs, mean LL, count, exp(mean LL)) as validation.txt
        self.stopped = False
        # engine != None: genotypes are already resident (device-generated synthetic data without
        # missing entries); bed_rows is then None
        if bed_rows is not None:
            bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8)

        # random number generation (cc:58-63)
        self._r = capi.Rng(env.seed)
        # init_heldout_sets -> set_validation_sample (cc:97, :196-224)       [RNG draws #1]
        self.val_loc, self.val_off, self.val_indiv = self._r.sample_validation(self._n, self._l, bed_rows)

        # shard geometry: contiguous blocks whose boundaries are multiples of 4
        per = (self._n + nranks - 1) // nranks
        per = (per + 3) // 4 * 4
        n_begin = min(rank * per, self._n)
        n_local = min(per, self._n - n_begin)
        self.n_begin, self.n_local = n_begin, n_local
        if engine is not None:
            assert engine.n_local == n_local and int(engine.cfg.n_begin) == n_begin
            self.engine = engine
        else:
            self.engine = capi.Engine(self._n, self._l, self._k, device=device, rank=rank, nranks=nranks,
                                      n_begin=n_begin, n_local=n_local,
                                      online_iterations=env.online_iterations)
            self.engine.load_bed(bed_rows)
        self.engine.set_validation(self.val_loc, self.val_off, self.val_indiv)
        if nranks > 1 and connect:  # connect=False: the caller wired the exchange already (dist.connect)
            handles = allgather(self.engine.comm_export())
            self.engine.comm_connect(handles)

        if env.compute_beta:
            # cc:74-95: gamma comes from gamma.txt (load_gamma); lambda is never initialised in
            # the reference (zero pages; equivalent to eta because a k-constant cancels)
            if gamma0 is None:
                raise ValueError("compute_beta needs gamma0 (the reference reads ./gamma.txt)")
            self.engine.set_gamma(np.asarray(gamma0)[n_begin:n_begin + n_local])
            return

        # init_gamma (cc:98, :226-237)                                        [RNG draws #2]
        g0 = self._r.init_gamma(self._n, self._k) if gamma0 is None else np.asarray(gamma0, dtype=np.float64)
        self.gamma0 = g0
        self.engine.set_gamma(g0[n_begin:n_begin + n_local])  # + estimate_all_theta (cc:101)
        # init_lambda (cc:99) happened in ts_create.
        # "+ computing initial heldout likelihood" (cc:103-104)
        self.compute_likelihood(True, True)
        self.save_model()

    # ------------------------------------------------------------------------------------
    def duration(self):
        return int(time.time() - self._start_time)

    def _reduce_ll(self, per_locus, count):
        """Add the shards' per-locus sums in rank order, then loci in ascending order."""
        if self.nranks > 1:
            parts = self._allgather((per_locus, count))
            per_locus = np.zeros_like(per_locus)
            count = 0
            for p, c in parts:
                per_locus = per_locus + p
                count += c
        s = 0.0
        for v in per_locus:
            s += float(v)
        return s, count

    def compute_likelihood(self, first, validation=True):
        """cc:461-544.  Returns True when the stopping rule fired."""
        assert validation, "-use-test-set is broken in the reference (SURVEY.md section 2)"
        _, cnt, per = self.engine.heldout_ll(first)
        if not first:
            self._iter += len(self.val_loc)  # snp_likelihood: _iter++ per locus (hh:333)
        s, k = self._reduce_ll(per, cnt)
        a = s / k if k else float("nan")  # the reference divides 0.0/0 here (N < 10): NaN, no stop
        self.validation_rows.append((self._iter, self.duration(), a, k, math.exp(a) if a == a else a))
        if self.env.outdir and self.rank == 0:
            with open(os.path.join(self.env.outdir, "validation.txt"), "a") as f:
                f.write("%d\t%d\t%.9f\t%d\t%f\n" % self.validation_rows[-1])
        stop = False
        if self._iter > 2000:
            if a > self._prev_h and self._prev_h != 0 and abs((a - self._prev_h) / self._prev_h) < self.env.stop_threshold:
                stop = True
            elif a < self._prev_h:
                self._nh += 1
            elif a > self._prev_h:
                self._nh = 0
            if a > self._max_h:
                self._max_h = a
            if self._nh > 3:
                stop = True
        self._prev_h = a
        if stop and self.env.use_validation_stop:
            self.save_model()
            self.stopped = True
            return True
        return False

    def save_model(self):
        """save_gamma (cc:546-576): gamma.txt / theta.txt, K x '%.8f\\t' per row."""
        if not self.env.outdir:
            return
        gamma, theta = self.gather_gamma(), self.gather_theta()
        if self.rank != 0:
            return
        suffix = "_%d.txt" % self._iter if self.env.file_suffix else ".txt"
        with open(os.path.join(self.env.outdir, "gamma" + suffix), "w") as f:
            f.write(_fmt_rows(gamma))
        with open(os.path.join(self.env.outdir, "theta" + suffix), "w") as f:
            f.write(_fmt_rows(theta))

    def _gather(self, local):
        if self.nranks == 1:
            return local
        return np.concatenate(self._allgather(local), axis=0)

    def gather_gamma(self):
        return self._gather(self.engine.gamma)

    def gather_theta(self):
        return self._gather(self.engine.theta)

    def infer(self, max_iter=None, max_seconds=None):
        """cc:417-459.  The reference never returns (exit(0) from compute_likelihood); here the
        loop returns when the stopping rule fires, on env.terminate, after max_iter
        iterations (tests) or at the first report after max_seconds (benchmarks).  SNP indices are pre-drawn up to the next report: in steady state
        the RNG is consumed by SNP sampling only, so the stream is unchanged."""
        rf = self.env.reportfreq
        t_begin = time.time()
        while not self.stopped:
            m = rf - self._iter % rf
            if max_iter is not None:
                if self._iter >= max_iter:
                    break
                m = min(m, max_iter - self._iter)
            locs = self._r.sample_locs(self._l, m)       # _loc = gsl_rng_uniform_int(_r, _l)
            self.engine.steps(locs)                      # optimize_lambda + gamma step, x m
            self._iter += m
            if self._iter % rf == 0:
                if self.compute_likelihood(False, True):
                    break
                self.save_model()
                if max_seconds is not None:  # benchmarks: give up at a report boundary, all ranks together
                    over = time.time() - t_begin > max_seconds
                    if self.nranks > 1:
                        over = self._allgather(over)[0]
                    if over:
                        break
            if self.env.terminate:
                self.save_model()
                break
        self.engine.sync()
        return self

    def compute_all_lambda(self):
        """-compute-beta sweep (cc:368-381): every locus in order; gamma keeps stepping."""
        locs = np.arange(self._l, dtype=np.uint32)
        self.engine.steps(locs)
        self._iter += self._l
        return self.engine.get_beta()
