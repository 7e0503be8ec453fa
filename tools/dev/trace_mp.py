# multi-process (torchrun) phase trace of rank 0's CTA 0: where a multi-GPU round spends its time
import os, sys
import numpy as np
os.environ["TSGPU_TRACE"] = "1"
# the phase stamps are compiled into the developer library only (make -C terastructure_b200/csrc trace; K = 10)
os.environ.setdefault("TSGPU_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "terastructure_b200", "lib", "libtsgpu_trace.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import terastructure_b200 as ts
from terastructure_b200 import synth
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
per, l, k = 125000, 20000, 10
n = per * world
_, beta = synth.psd_params(1, l, k, seed=1)
theta = np.random.RandomState(5 + rank).dirichlet(np.full(k, 0.1), size=per)
e = ts.Engine(n, l, k, device=lr, rank=rank, nranks=world, n_begin=rank * per, n_local=per)
e.synth_bed(1, theta, beta, 0.0)
rng = ts.Rng(1234)
vl, vo, vi = rng.sample_validation(n, l, None)
e.set_validation(vl, vo, vi)
e.set_gamma(rng.init_gamma(n, k)[rank * per:(rank + 1) * per])
from terastructure_b200 import dist as tsdist
xinfo = tsdist.connect(e)
e.steps(rng.sample_locs(l, 200)); e.sync(); dist.barrier()
e.timer_start(); e.steps(rng.sample_locs(l, 64)); ms = e.timer_stop()
if rank == 0:
    sys.stdout = open("gpurun_out/trace_mp_%d_%s.txt" % (world, os.environ.get("TSGPU_XCHG", "default")), "w")
    print("%d GPUs: %.1f us per SNP; exchange: %s" % (world, 1e3 * ms / 64, xinfo))
    t = e.debug_trace().astype(np.float64)
    it = t[8:60]
    def d(a, b): return np.mean(it[:, b] - it[:, a])
    acc = np.zeros(7)
    for x in range(9):  # rounds 0..8 (the last one hides the gamma step)
        base = 2 + 8 * x
        prev = (2 + 8 * (x - 1) + 6) if x else 1
        acc += np.array([d(prev, base), d(base, base + 1), d(base + 1, base + 2), d(base + 2, base + 3),
                         d(base + 3, base + 4), d(base + 4, base + 5), d(base + 5, base + 6)])
    names = ["E-step", "tr_reduce", "sync1", "CTA-sum + arrival (atomic with return, forward if last)", "wait for the ranks", "lambda/b", "sync2"]
    print("mean/round (rounds 0-8): " + "  ".join("%s %.0f" % (nm, v / 9) for nm, v in zip(names, acc)) + "  total %.0f" % (acc.sum() / 9))
    print("whole item %.0f cycles" % np.mean(it[1:, 0] - it[:-1, 0]))
    sys.stdout.flush()
e.close()
dist.destroy_process_group()
