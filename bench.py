#!/usr/bin/env python3
"""bench.py -- genotypes/s of SVI updates (BASELINE.json metric) on synthetic PSD genotypes.

  python bench.py --gpus N --steps K --warmup W                 our arm (one process per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   the reference's CPU build

One "step" = one batch of BATCH consecutive SVI iterations (sample SNP -> <=10 rounds of
{E-step over all individuals, 2K-sum, lambda update} -> gamma step) handed to the engine in a
single ts_steps call.  genotypes/s = N_individuals * SVI iterations / seconds.

Workload: N=1 -> BASELINE configs[2] (100K individuals x 1M SNPs, K=10, 25 GB packed), the
largest configuration that fits one B200; the metric's own shape (configs[3], 1M x 1M, 250 GB
packed) needs 8 GPUs.  N>1 -> 125K individuals per GPU x 1M SNPs (N=8 is exactly configs[3]);
individuals are sharded, every round exchanges 2K doubles between the GPUs ("weak" scaling).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 10
L_FULL = 1_000_000
N_ONE_GPU = 100_000
N_PER_GPU_MULTI = 125_000
BATCH = 1000           # SVI iterations per step
DATA_SEED = 1
INFER_SEED = 1234
METRIC = "genotypes/sec of SVI updates"
UNIT = "genotypes/s"


def algorithmic_bytes_per_genotype(k):
    """0.25 (packed y) + 8K x {read Elogtheta, read gamma, write gamma, write Elogtheta}."""
    return 32.0 * k + 0.25


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_kernel_metrics(k, ipt, n_per):
    """Per-instantiation numbers that only a profiler can give, from the committed `ncu --set full`
    captures (profiles/r2_ncu_metrics.json, made by tools/ncu_extract.py from the .ncu-rep files):
    DRAM bytes per SVI iteration and FP64 pipe utilisation of tsp::k_persist<K,I> at this shard size.
    None when no capture of this instantiation/shard size is committed."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_metrics.json")
    try:
        m = json.load(open(p)).get("%d,%d,%d" % (k, ipt, n_per))
        return m
    except Exception:
        return None


def kernel_label(ts, n_per, k):
    """Name and launch geometry of the dominant kernel for a shard of n_per individuals."""
    label = "tsp::k_persist<%d" % k
    try:
        ipt, grid, block = ts.plan_shard(n_per, k)
        label += ",%d>, %d CTAs x %d threads" % (ipt, grid, block)
    except Exception:
        label += ">"
    return label + " (one cooperative launch per step = BATCH SVI iterations)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "terastructure_ref")
REF_SAMPLE_L = 2000   # the reference holds N x L BYTES unpacked: cut L so it fits host RAM
REF_MAX_N = 250_000   # largest sample the reference arm runs (N x L bytes in host RAM)
REF_STEP_ITERS = 100  # the reference prints a progress line every 100 iterations (snpsamplinge.cc:436-439)


def write_sample_bed(n, l, k, path_prefix):
    """A small-L sample of the bench's synthetic distribution for the CPU arm (same N and K)."""
    from terastructure_b200 import plink, synth
    theta, beta = synth.psd_params(n, l, k, seed=DATA_SEED)
    rs = np.random.RandomState(DATA_SEED + 7919)
    rows = np.empty((l, (n + 3) // 4), np.uint8)
    for lo in range(0, l, 100):  # bounded memory
        q = np.clip(beta[lo:lo + 100] @ theta.T, 0, 1)
        rows[lo:lo + 100] = plink.pack(rs.binomial(2, q).astype(np.uint8))
    return plink.write_bed(path_prefix, rows, n)


def run_reference_cpu(n, k, steps, warmup, budget_s=150.0, threads=None, tmp=None):
    """Time the reference binary (unmodified sources, oracle/_ref) with `threads` worker threads
    (default: all host threads).  A step = 100 SVI iterations (its progress-line granularity).
    Returns a dict or raises."""
    if not os.path.exists(REF_BIN):
        raise RuntimeError(f"{REF_BIN} not built (make -C oracle ref)")
    cores = threads or len(os.sched_getaffinity(0))
    own_tmp = tmp is None
    if own_tmp:
        tmp = tempfile.mkdtemp(prefix="tsbench.")
        write_sample_bed(n, REF_SAMPLE_L, k, os.path.join(tmp, "s"))
    subprocess.run(["rm", "-rf", os.path.join(tmp, f"n{n}-k{k}-l{REF_SAMPLE_L}-b-seed{INFER_SEED}")])
    cmd = [REF_BIN, "-file", "s.bed", "-n", str(n), "-l", str(REF_SAMPLE_L), "-k", str(k), "-stochastic",
           "-nthreads", str(cores), "-rfreq", "100000000", "-seed", str(INFER_SEED), "-label", "b"]
    p = subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    stamps = {}  # iteration -> time
    buf = b""
    t_start = time.time()
    want_last = (warmup + steps) * REF_STEP_ITERS
    try:
        os.set_blocking(p.stdout.fileno(), False)
        while p.poll() is None:
            chunk = p.stdout.read(65536)
            now = time.time()
            if chunk:
                buf += chunk
                if b"initialization end" in buf:      # ctor done (snpsamplinge.cc:112): infer() starts
                    stamps.setdefault(0, now)
                for m in re.finditer(rb"iteration = (\d+) took", buf):
                    stamps.setdefault(int(m.group(1)), now)
                buf = buf[-64:]
            else:
                time.sleep(0.01)
            done = [i for i in stamps if i >= want_last]
            first_timed = warmup * REF_STEP_ITERS
            have = sorted(i for i in stamps if i >= first_timed)
            if done or (now - t_start > budget_s and len(have) >= 2) or now - t_start > 2 * budget_s:
                break
    finally:
        p.kill()
        p.wait()
        if own_tmp:
            subprocess.run(["rm", "-rf", tmp])
    its = sorted(stamps)
    if len(its) < 2:
        raise RuntimeError(f"reference produced {len(its)} progress lines in {time.time() - t_start:.0f}s")
    i0 = max(warmup * REF_STEP_ITERS, its[0])
    i0 = min((i for i in its if i >= i0), default=its[0])
    i1 = its[-1] if its[-1] <= want_last else want_last
    if i1 <= i0:
        i0, i1 = its[0], its[-1]
    dt = stamps[i1] - stamps[i0]
    nsteps = (i1 - i0) // REF_STEP_ITERS
    return {"value": n * (i1 - i0) / dt, "cores": cores, "iters": i1 - i0, "seconds": dt, "steps": nsteps,
            "ms_per_step": 1e3 * dt / max(nsteps, 1),
            "sample": f"{i1 - i0} SVI iterations at N={n}, K={k}, L cut to {REF_SAMPLE_L} (the reference stores "
                      f"N x L bytes unpacked; cost per iteration does not depend on L), -nthreads {cores}, "
                      f"timed between its own progress lines {i0} and {i1}"}


def run_reference_best(n, k, steps, warmup, budget_s):
    """the reference with -nthreads in {1, nproc/2, nproc}; the best one is reported
    (its thread pool is sync-bound at some sizes, so more threads are not always faster)."""
    nproc = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(prefix="tsbench.")
    tried = {}
    best = None
    try:
        write_sample_bed(n, REF_SAMPLE_L, k, os.path.join(tmp, "s"))
        for t in sorted({1, max(1, nproc // 2), nproc}, reverse=True):
            try:
                r = run_reference_cpu(n, k, steps, warmup, budget_s=budget_s, threads=t, tmp=tmp)
            except Exception as ex:
                tried[str(t)] = f"failed: {ex}"
                continue
            tried[str(t)] = r["value"]
            if best is None or r["value"] > best["value"]:
                best = r
    finally:
        subprocess.run(["rm", "-rf", tmp])
    if best is None:
        raise RuntimeError(f"reference failed at every thread count: {tried}")
    best["tried"] = tried
    best["sample"] += f"; best of -nthreads {sorted(int(t) for t in tried)} (genotypes/s: {tried})"
    return best


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = N_ONE_GPU if args.gpus == 1 else N_PER_GPU_MULTI * args.gpus
    n_cpu = min(n, REF_MAX_N)  # N x L bytes in host RAM; per-genotype cost is N-independent at this size
    r = run_reference_best(n_cpu, K, max(1, min(args.steps, 3)), min(args.warmup, 1), budget_s=60.0)
    cfg = workload_config(args.gpus)
    cfg["reference_ran_individuals"] = n_cpu
    cfg["reference_ran_snps"] = REF_SAMPLE_L
    cfg["reference_note"] = ("the reference holds N x L bytes unpacked in host RAM: it is timed on a sample of "
                             f"{n_cpu} individuals x {REF_SAMPLE_L} SNPs of the same distribution and K; its cost per "
                             "genotype does not depend on L and is flat in N at this size")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                         "sample": r["sample"], "threads_tried": r["tried"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(gpus):
    n = N_ONE_GPU if gpus == 1 else N_PER_GPU_MULTI * gpus
    return {"workload": f"synthetic PSD genotypes, {n} individuals x {L_FULL} SNPs, K={K} "
                        + ("(BASELINE configs[2], largest single-GPU config)" if gpus == 1 else
                           "(125K individuals per GPU; 8 GPUs = configs[3], 1M x 1M)"),
            "individuals": n, "snps": L_FULL, "K": K, "svi_iterations_per_step": BATCH,
            "sharding": f"individuals over {gpus} GPU(s)",
            "l2": "state (gamma + exp(psi(gamma)), 16 MB/100K individuals) is L2-resident by design between "
                  "consecutive SNPs; genotype columns come from a 25+ GB array (>> L2); L2 flushed between steps"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def parity_check(eng, locs, vl, n_per, world, rank, dist, torch):
    """After the timed region: (1) the lambda rows of the loci visited last are BIT-IDENTICAL on all
    ranks (every rank forms lambda from the same integer totals); (2) sum_k lambda[loc][k][0] - K*eta0 =
    sum over all individuals of y and the same for 2 - y (phi sums to one over k), from the genotype
    rows each rank holds.  Validation loci are skipped (their held-out individuals do not count)."""
    from terastructure_b200 import plink
    val = set(int(v) for v in vl)
    pick = []
    for x in locs[::-1]:
        x = int(x)
        if x not in val and x not in pick:
            pick.append(x)
        if len(pick) == 16:
            break
    lam = np.stack([eng.get_lambda(x, 1)[0] for x in pick])                       # [16, K, 2]
    ysum = np.zeros((len(pick), 2))
    for i, x in enumerate(pick):
        y = plink.unpack(eng.get_bed_row(x)[None, :], n_per)[0].astype(np.int64)
        ok = y != 3
        ysum[i] = (y[ok].sum(), (2 - y[ok]).sum())
    identical = True
    if world > 1:
        t = torch.from_numpy(lam.view(np.int64).copy()).cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        identical = all(bool(torch.equal(parts[0], q)) for q in parts[1:])
        ys = torch.from_numpy(ysum).cuda()
        dist.all_reduce(ys)
        ysum = ys.cpu().numpy()
    k = lam.shape[1]
    got = np.stack([lam[:, :, 0].sum(1) - k, lam[:, :, 1].sum(1) - k], axis=1)
    err = float(np.max(np.abs(got - ysum) / np.maximum(ysum, 1.0)))
    return {"ok": bool(identical and err < 1e-9), "lambda_bit_identical_across_ranks": bool(identical),
            "sum_k_identity_max_rel_err": err, "loci_checked": len(pick)}


def nccl_allreduce_latency_us(dist, torch, k, iters=200):
    """The alternative the north star names: one ncclAllReduce of the 2K fp64 statistics, timed on the
    device back to back (latency per call, the collective a per-round host-launched exchange would pay)."""
    t = torch.ones(2 * k, dtype=torch.float64, device="cuda")
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def solo_shard_us(ts, synth, n_per, k, local_rank, iters=1000):
    """The same shard of individuals ALONE on this GPU (no exchange), L cut to 20 000 (cost per SVI
    iteration does not depend on L): microseconds per SVI iteration."""
    l = 20_000
    _, beta = synth.psd_params(1, l, k, seed=DATA_SEED)
    theta = np.random.RandomState(DATA_SEED + 1000).dirichlet(np.full(k, 0.1), size=n_per)
    e = ts.Engine(n_per, l, k, device=local_rank)
    e.synth_bed(DATA_SEED, theta, beta, 0.0)
    r = ts.Rng(INFER_SEED)
    vl, vo, vi = r.sample_validation(n_per, l, None)
    e.set_validation(vl, vo, vi)
    e.set_gamma(r.init_gamma(n_per, k))
    for _ in range(2):
        e.steps(r.sample_locs(l, iters))
    e.sync()
    best = 1e30
    for _ in range(3):
        locs = r.sample_locs(l, iters)
        e.timer_start()
        e.steps(locs)
        best = min(best, e.timer_stop())
    e.close()
    return 1e3 * best / iters


def convergence_run(ts, eng, n_total, l, k, world, rank, local_rank, dist, max_seconds):
    """BASELINE metric, second half: wall-clock to the reference's convergence criterion
    (compute_likelihood's stop rule, snpsamplinge.cc:510-541, -rfreq 100000 = the reference default),
    from a fresh initialisation (same seed as the reference would use) on the resident genotypes."""
    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    eng.reset_lambda()
    eng.reset_counts()
    env = ts.Env(n_total, k, l, seed=INFER_SEED, rfreq=100_000)
    t0 = time.perf_counter()
    s = ts.SNPSamplingE(env, None, device=local_rank, rank=rank, nranks=world,
                        allgather=allgather if world > 1 else None, engine=eng, connect=False)
    t_init = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # every rank must stop at the same report: the time limit is checked on rank 0's clock
    s.infer(max_seconds=max_seconds)
    eng.sync()
    if world > 1:
        dist.barrier()
    t_run = time.perf_counter() - t0
    return {"converged": bool(s.stopped), "converged_at_iter": int(s._iter), "time_to_convergence_s": t_run,
            "init_s": t_init, "rfreq": 100_000, "stop_threshold": 1e-5,
            "heldout_ll": [(int(r[0]), float(r[2])) for r in s.validation_rows],
            "note": "includes the held-out passes of every report; stop rule of snpsamplinge.cc:510-541"
                    + ("" if s.stopped else f"; NOT converged within the {max_seconds:.0f} s limit of this run")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--individuals", type=int, default=0, help="override individuals per GPU (debug)")
    ap.add_argument("--snps", type=int, default=0, help="override L (debug)")
    ap.add_argument("--batch", type=int, default=0, help="override SVI iterations per step (debug/profiling)")
    ap.add_argument("--k", type=int, default=0, help="override K (BASELINE configs[4] sweep)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the measurements outside the timed region (solo shard, NCCL latency, convergence run)")
    ap.add_argument("--converge-seconds", type=float, default=150.0, help="time limit of the convergence run (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3
    global BATCH, K
    if args.batch:
        BATCH = args.batch
    if args.k:
        K = args.k
    override = bool(args.individuals or args.snps or args.batch or args.k)

    import torch
    import torch.distributed as dist
    import terastructure_b200 as ts
    from terastructure_b200 import synth
    from terastructure_b200 import dist as tsdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_per = args.individuals or (N_ONE_GPU if world == 1 else N_PER_GPU_MULTI)
    n_total = n_per * world
    l = args.snps or L_FULL
    n_begin = rank * n_per

    # ---- synthetic data, generated on the device (25 GB at the full shape) ----
    _, beta = synth.psd_params(1, l, K, seed=DATA_SEED)
    theta = np.random.RandomState(DATA_SEED + 1000 + rank).dirichlet(np.full(K, 0.1), size=n_per)
    eng = ts.Engine(n_total, l, K, device=local_rank, rank=rank, nranks=world, n_begin=n_begin, n_local=n_per)
    eng.synth_bed(DATA_SEED, theta, beta, 0.0)
    rng = ts.Rng(INFER_SEED)
    vl, vo, vi = rng.sample_validation(n_total, l, None)       # set_validation_sample, exact draw order
    eng.set_validation(vl, vo, vi)
    g0 = rng.init_gamma(n_total, K)                             # init_gamma, N*K sequential gamma draws
    eng.set_gamma(g0[n_begin:n_begin + n_per])
    xinfo = tsdist.connect(eng) if world > 1 else {"exchange": "none (one GPU)"}
    ipt, grid, block = eng.plan

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    total_steps = args.warmup + args.steps
    locs = rng.sample_locs(l, BATCH * total_steps).reshape(total_steps, BATCH)
    pinned = torch.from_numpy(locs.copy()).pin_memory().numpy()

    # ---- device-timed region: K steps, CUDA events on the engine's stream, inputs resident ----
    for s in range(args.warmup):
        eng.steps(pinned[s])
    eng.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for s in range(args.warmup, total_steps):
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()   # ranks enter the step together (the flush above is not part of the step)
        eng.timer_start()
        eng.steps(pinned[s])
        dev_ms += eng.timer_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    clocks = sampler.stop()

    # ---- end to end through the public API: host RNG draws the SNPs, host buffers in, rounds out ----
    rng2 = ts.Rng(INFER_SEED + 1)
    barrier()
    t0 = time.perf_counter()
    rounds_total = 0
    last_batch = None
    for s in range(args.steps):
        batch = rng2.sample_locs(l, BATCH)                    # host-side SNP sampling (bit-exact GSL stream)
        rounds = eng.steps(batch, want_rounds=True)           # H2D work items, kernels, D2H rounds, sync
        rounds_total += int(rounds.sum())
        last_batch = batch
    barrier()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])

    # ---- outside the timed regions: parity of what was just computed, and the side measurements ----
    parity = parity_check(eng, last_batch, vl, n_per, world, rank, dist, torch)
    us_iter = 1e3 * dev_ms / (BATCH * args.steps)
    mean_rounds = rounds_total / (BATCH * args.steps)
    extras = {}
    if not args.no_extras:
        if world > 1:
            extras["nccl_allreduce_2k_f64_us"] = nccl_allreduce_latency_us(dist, torch, K)
            solo = torch.tensor([solo_shard_us(ts, synth, n_per, K, local_rank)], dtype=torch.float64, device="cuda")
            dist.all_reduce(solo, op=dist.ReduceOp.MAX)
            extras["same_shard_1gpu_us"] = float(solo[0])
            extras["exchange_us_per_round"] = (us_iter - float(solo[0])) / max(mean_rounds, 1.0)
            extras["exchange_us_per_snp"] = us_iter - float(solo[0])
            extras["efficiency_vs_same_shard_alone"] = float(solo[0]) / us_iter
        if args.converge_seconds > 0 and not override:
            try:
                extras["convergence"] = convergence_run(ts, eng, n_total, l, K, world, rank, local_rank, dist,
                                                        args.converge_seconds)
            except Exception as ex:
                extras["convergence"] = {"converged": False, "note": f"failed: {type(ex).__name__}: {ex}"}

    if rank == 0:
        genos = float(n_total) * BATCH * args.steps
        value = genos / (dev_ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        bpg = algorithmic_bytes_per_genotype(K)
        # per GPU: each GPU's kernels process n_per individuals per SVI iteration
        achieved = (float(n_per) * BATCH * args.steps * bpg) / (dev_ms * 1e-3) / 1e9
        prof = profiled_kernel_metrics(K, ipt, n_per)
        cfg = workload_config(world) if not override else {
            "workload": f"override: {n_total} individuals x {l} SNPs, K={K}", "individuals": n_total, "snps": l, "K": K,
            "svi_iterations_per_step": BATCH}
        cfg["exchange"] = xinfo
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "clocks": clocks,
            "e2e": {"value": genos / e2e_s, "unit": UNIT, "h2d_bytes_per_step": BATCH * 24,
                    "d2h_bytes_per_step": BATCH * 4,
                    "note": "host draws SNP indices with the GSL-exact RNG, ts_steps(host buffer), rounds read back"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (prof["dram_bytes_per_svi_iteration"] * BATCH) if prof else None,
                         "traffic_note": "dram__bytes_read+write of one launch (= one step) from the committed ncu --set full "
                                         "capture of this instantiation and shard size (profiles/r2_ncu_metrics.json); "
                                         "null = no capture of this geometry",
                         "peak_source": peak_src,
                         "kernel": "tsp::k_persist<%d,%d,%s,%s> (K, individuals per thread in registers, tiered, several GPUs), "
                                   "%d CTAs x %d threads (one cooperative launch per step = %d SVI iterations)"
                                   % (K, ipt, "true" if eng.tiers[0] >= 0 else "false", "true" if world > 1 else "false", grid, block, BATCH),
                         "algorithmic_bytes_per_genotype": bpg,
                         "fp64_pipe_pct": prof["fp64_pipe_pct"] if prof else None,
                         "fp64_note": "sm__pipe_fp64_cycles_active (% of peak) from the same capture; the path is bound by the "
                                      "latency of 10 dependent grid-wide reductions per SVI iteration, not by HBM or FP64 "
                                      "(DESIGN.md section 4.1); FP64 ceiling of this algorithm = 0.67 of the HBM roofline"},
            "wall_s_timed_region": t_wall, "mean_rounds_per_snp": mean_rounds,
            "us_per_svi_iteration": us_iter,
            "parity_check": parity,
        }
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:
            try:
                r = run_reference_cpu(min(n_total, REF_MAX_N), K, 1, 0, budget_s=40.0)
                line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                                        "kind": "reference", "sample": r["sample"]}
            except Exception as ex:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)),
                                        "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
