#!/usr/bin/env python3
"""f2 at scale: write a synthetic .bed of BASELINE configs[2] size (100 000 individuals x
1 000 000 SNPs = 25 GB packed), run the drop-in CLI on it exactly as a user would, and report the load
rate (the CLI's own "+ genotypes resident" line), the time to the end of the RNG-exact initialisation
and the time to the first 100 SVI iterations.  The process is then stopped with SIGTERM like the
reference (main.cc:28-39: save the model, exit 0).

  python tools/dev/big_bed_test.py [--n 100000] [--l 1000000] [--dir /dev/shm] [--gpus 1]"""
import argparse, json, os, signal, subprocess, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--l", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/big_bed.json")
a = ap.parse_args()
d = os.path.join(a.dir, "tsbig")
os.makedirs(d, exist_ok=True)
bps = (a.n + 3) // 4
t0 = time.time()
rs = np.random.RandomState(1)
# genotype codes 00/10/11 (no missing) with allele frequency ~0.3: built from a table of valid bytes
codes = np.array([0, 2, 3], np.uint8)
tab = np.array([codes[i % 3] | (codes[(i // 3) % 3] << 2) | (codes[(i // 9) % 3] << 4) | (codes[(i // 27) % 3] << 6) for i in range(81)], np.uint8)
chunk_rows = max(1, (256 << 20) // bps)
base = tab[rs.randint(0, 81, size=(chunk_rows, bps))]
with open(os.path.join(d, "big.bed"), "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01]))
    done = 0
    while done < a.l:
        m = min(chunk_rows, a.l - done)
        blk = np.roll(base[:m], done % 977, axis=1)      # cheap variation between chunks
        f.write(blk.tobytes())
        done += m
with open(os.path.join(d, "big.bim"), "w") as f:
    f.write("1\n" * a.l)
with open(os.path.join(d, "big.fam"), "w") as f:
    f.write("1\n" * a.n)
t_write = time.time() - t0
size = os.path.getsize(os.path.join(d, "big.bed"))
exe = os.path.join(ROOT, "terastructure_b200", "bin", "terastructure")
cmd = [exe, "-file", "big.bed", "-n", str(a.n), "-l", str(a.l), "-k", str(a.k), "-stochastic", "-rfreq", "100000000",
       "-seed", "1234", "-label", "big", "-gpus", str(a.gpus), "-force"]
t0 = time.time()
p = subprocess.Popen(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
os.set_blocking(p.stdout.fileno(), False)
os.set_blocking(p.stderr.fileno(), False)
so, se, marks = b"", b"", {}
while p.poll() is None and time.time() - t0 < 900:
    so += p.stdout.read() or b""
    se += p.stderr.read() or b""
    now = time.time() - t0
    for key, pat in (("init_begin", b"initialization begin"), ("init_end", b"initialization end"), ("iter_100", b"iteration = 100 took")):
        if key not in marks and pat in so:
            marks[key] = now
    if b"genotypes resident" in se and "resident" not in marks:
        marks["resident"] = now
    if "iter_100" in marks:
        p.send_signal(signal.SIGTERM)
        break
    time.sleep(0.05)
rc = p.wait(timeout=300)
se += p.stderr.read() or b""
line = [l for l in se.decode().splitlines() if "genotypes resident" in l]
out = os.path.join(d, f"n{a.n}-k{a.k}-l{a.l}-big-seed1234")
res = {"bed_bytes": size, "individuals": a.n, "snps": a.l, "K": a.k, "gpus": a.gpus, "write_seconds": t_write, "dir": a.dir,
       "cli_exit_code": rc, "seconds_since_start": marks, "cli_load_line": line[0] if line else None,
       "gamma_txt_rows": sum(1 for _ in open(os.path.join(out, "gamma.txt"))) if os.path.exists(os.path.join(out, "gamma.txt")) else None,
       "param_txt_tail": open(os.path.join(out, "param.txt")).read().splitlines()[-14:-9] if os.path.exists(os.path.join(out, "param.txt")) else None}
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
subprocess.run(["rm", "-rf", d])
