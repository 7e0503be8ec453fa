#!/usr/bin/env python3
"""Wall-clock to the reference's convergence criterion (BASELINE.json metric, second half) on
synthetic PSD genotypes generated on the device.  One process per GPU (torchrun) or a single GPU.

  python tools/converge.py [--per-gpu 125000] [--snps 1000000] [--k 10] [--rfreq 100000]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import terastructure_b200 as ts
from terastructure_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--per-gpu", type=int, default=125000)
ap.add_argument("--snps", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--rfreq", type=int, default=100000)
ap.add_argument("--max-iter", type=int, default=3_000_000)
ap.add_argument("--out", default="gpurun_out/converge.json")
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def allgather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out

n, l, k = a.per_gpu * world, a.snps, a.k
t0 = time.time()
_, beta = synth.psd_params(1, l, k, seed=1)
theta = np.random.RandomState(1000 + rank).dirichlet(np.full(k, 0.1), size=a.per_gpu)
eng = ts.Engine(n, l, k, device=lr, rank=rank, nranks=world, n_begin=rank * a.per_gpu, n_local=a.per_gpu)
eng.synth_bed(1, theta, beta, 0.0)
t_data = time.time() - t0
env = ts.Env(n, k, l, seed=1234, rfreq=a.rfreq)
t0 = time.time()
s = ts.SNPSamplingE(env, None, device=lr, rank=rank, nranks=world, allgather=allgather if world > 1 else None, engine=eng)
t_init = time.time() - t0
if world > 1: dist.barrier()
t0 = time.time()
s.infer(max_iter=a.max_iter)
eng.sync()
if world > 1: dist.barrier()
t_run = time.time() - t0
if rank == 0:
    res = {"individuals": n, "snps": l, "K": k, "gpus": world, "rfreq": a.rfreq, "iterations": s._iter, "stopped_by_rule": s.stopped,
           "seconds_data_generation": t_data, "seconds_init": t_init, "seconds_to_convergence": t_run,
           "genotypes_per_second_incl_reports": n * float(s._iter) / t_run,
           "validation": [(r[0], r[2], r[3]) for r in s.validation_rows]}
    print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
